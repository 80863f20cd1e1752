// TMA-fed streaming kernels (sm_100a): the multicolour smoother of the big levels and the SpMV-class operators
// (residual / restriction / prolongation) read the immutable operator through a shared-memory ring that one elected
// producer thread fills with 1-D bulk copies (cp.async.bulk, completion on an mbarrier with complete_tx); eight
// consumer warps take the row chunks out of shared memory, gather the vector from L2, fold and store.
//
// Why (profiles/r01_sor_mc_packed_4M_ncu.txt): the register-fed sweep was latency bound -- long_scoreboard 16.3 and
// barrier 5.2 stalls per issue, DRAM 48 % busy with clean traffic (1.01x algorithmic): every colour phase drained the
// memory pipeline into grid.sync() and refilled it afterwards, static tiles quantised 7.04 tiles per CTA to 8, and the
// loads in flight were bounded by registers.  Here
//   * the bytes in flight are bounded by shared memory (stages x tile bytes per CTA), not by registers;
//   * the operator is immutable, only x is phase ordered: the producer runs ahead ACROSS the colour barrier, so the
//     first tiles of colour c+1 are already resident when the barrier opens;
//   * tiles are handed out by a per-phase ticket counter (atomicAdd, next ticket prefetched), so no CTA is left with a
//     whole extra tile at the end of a phase;
//   * the colour barrier is one arrival counter per phase (red.release / ld.acquire), no cooperative-groups object.
// Arithmetic per row is the one of k_sor_mc_packed / k_spmv2 (same lane mapping, same reduction tree): same bits.
#include <algorithm>
#include <cstdlib>

#include "mmg_device.cuh"
#include "mmg_internal.hpp"

namespace mmg {

namespace {

constexpr int kConsumerWarps = 8;
constexpr int kConsumers = kConsumerWarps * 32;
constexpr int kStreamThreads = kConsumers + 32;   // + the producer warp (one elected lane issues the copies)
constexpr int kMaxStages = 12;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {   // try_wait suspends in hardware between probes
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MMG_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MMG_DONE;\n"
      "bra MMG_WAIT;\n"
      "MMG_DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine, SASS UBLKCP), bytes a multiple of 16, completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar, unsigned long long policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void red_release_add(int* p, int v) { asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory"); }

struct RingCtl {
  unsigned long long full[kMaxStages], empty[kMaxStages];
  int phase[kMaxStages], row0[kMaxStages], nrows[kMaxStages];
};

__device__ __forceinline__ RingCtl* ring_setup(unsigned char* smem, int stages, unsigned tile_bytes, unsigned full_arrivals = 1) {
  RingCtl* C = reinterpret_cast<RingCtl*>(smem + (size_t)stages * tile_bytes);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; s++) { mbar_init(&C->full[s], full_arrivals); mbar_init(&C->empty[s], kConsumerWarps); }

    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  return C;
}

// chunk of local row `lr` of a staged tile -> registers (lane gl of an LPR-lane group takes slots gl, gl+LPR, ...)
template <int LPR, int ITER>
__device__ __forceinline__ void tile_row_fetch(const unsigned char* tile, unsigned chunk_bytes, int W, int lr, bool valid, int gl, double (&v)[ITER], int (&c)[ITER]) {
  const double* pv = reinterpret_cast<const double*>(tile + (size_t)(valid ? lr : 0) * chunk_bytes);
  const int* pc = reinterpret_cast<const int*>(pv + W);
#pragma unroll
  for (int t = 0; t < ITER; t++) {
    const int k = gl + t * LPR;
    const bool ok = valid && k < W;
    v[t] = ok ? pv[k] : 0.0;
    c[t] = ok ? pc[k] : -1;
  }
}

// ------------------------------------------------------------------------------------------------
// Multicolour SOR, every phase of all sweeps of one smoothing call over the colour-major packed copy of the operator
// (DESIGN.md section 3).  A sweep is `pps` phases: the interior colours and, on a grid with a Neumann boundary, one more
// phase that (a) evaluates the Neumann boundary rows -- Grid::bound_eval_neumann, grid.cpp:73-103: they read interior values
// and themselves only, so they are one more set of independent rows, stored as the last range of the packed copy -- and
// (b) forms the dot product of the dense regularisation row (grid.cpp:566-576), the last "colour" of the sweep: every CTA
// reduces a fixed slice into reg_partial[sweep][cta]; after that phase's barrier every CTA folds the partials in the same
// fixed order and stores the same new value of the regularisation unknown.  No atomics on doubles: the result is
// reproducible for a given grid size.  Rows longer than the chunk width (implicit-Neumann fill-in, grid.cpp:607-657) carry
// bit 30 in their diagonal column and fetch their tail from the overflow CSR of the natural-order operator.
// ctl[0 .. nphases) = tile tickets, ctl[nphases .. 2 nphases) = barrier arrivals, zeroed by the host before the launch.
// Cooperative launch (the colour barrier spins), one producer warp + eight consumer warps per CTA.
// ------------------------------------------------------------------------------------------------
struct RegRow {            // regularisation row of a Neumann-type grid (reg_row < 0: none)
  const int* col;
  const double* val;
  int len, row;
  double diag;
  double* partial;         // iters x gridDim.x
};
constexpr int kColMask = 0x3fffffff;   // bit 31: neighbour of a lower colour (k_sor_mc_flow), bit 30: the row has an overflow tail

__device__ __forceinline__ void grid_arrive(int* arrivals, int p) {
  consumer_sync();
  if (threadIdx.x == 0) { __threadfence(); red_release_add(&arrivals[p], 1); }
}
__device__ __forceinline__ void grid_wait(const int* arrivals, int p, int* abort_flag, long long timeout_cycles) {
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while (ld_acquire(&arrivals[p]) < (int)gridDim.x) {
      // watchdog: never hang the device; the sweep is abandoned (results invalid) and the host raises MMG_ERR_TIMEOUT
      if (*(volatile int*)abort_flag || clock64() - t0 > timeout_cycles) { atomicExch(abort_flag, 1); break; }
    }
  }
  consumer_sync();
}

template <int LPR, int ITER, int ROWS>
__global__ void __launch_bounds__(kStreamThreads) k_sor_mc_tma(const unsigned char* __restrict__ chunks, unsigned chunk_bytes, int W,
                                                               const int* __restrict__ phase_ptr, int pps, int bnd_phase, int iters,
                                                               const double* __restrict__ b, double* x, double omega, int* ctl, int stages,
                                                               int dynamic, int* abort_flag, long long timeout_cycles, int debug_flags, HybView A,
                                                               RegRow reg, int static_8ths) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ double red_scratch[kConsumerWarps];
  constexpr int GPW = 32 / LPR;
  constexpr int TR = kConsumerWarps * GPW * ROWS;      // rows per tile
  const unsigned tile_bytes = TR * chunk_bytes;
  RingCtl* C = ring_setup(smem, stages, tile_bytes);
  const int nphases = iters * pps;
  int* tickets = ctl;
  int* arrivals = ctl + nphases;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == kConsumerWarps) {                          // ---- producer
    if (lane != 0) return;
    const unsigned long long pol = policy_evict_first();
    int s = 0;
    unsigned par = 0;
    int phase = 0;
    // Tiles of a phase: every CTA first takes a contiguous share of `share` tiles (neighbours along the Z-curve, so consecutive
    // tiles of a CTA re-use each other's x lines in L1), the rest is dealt by tickets to even out the end of the phase.
    const int G = (int)gridDim.x;
    int j = 0, ticket = -1;
    for (;;) {
      int t = 0, first = 0, count = 0, share = 0;
      while (phase < nphases) {
        const int c = phase % pps;
        first = phase_ptr[c];
        count = phase_ptr[c + 1] - first;
        const int tiles = (count + TR - 1) / TR;
        share = dynamic ? (int)(((long long)tiles * static_8ths) / (8ll * G)) : 0;
        if (j < share) { t = (int)blockIdx.x * share + j; break; }
        if (ticket < 0) ticket = dynamic ? atomicAdd(&tickets[phase], 1) : (int)blockIdx.x;
        t = G * share + ticket;
        if (t < tiles) break;                            // otherwise the ticket is stale: the phase has no tiles left
        ++phase; j = 0; ticket = -1;
      }
      mbar_wait(&C->empty[s], par ^ 1u);
      if (phase >= nphases) {                            // terminator
        C->phase[s] = nphases;
        mbar_arrive(&C->full[s]);
        return;
      }
      if (j < share) { if (++j == share) ticket = dynamic ? atomicAdd(&tickets[phase], 1) : (int)blockIdx.x; }
      else ticket = dynamic ? atomicAdd(&tickets[phase], 1) : ticket + G;       // in flight while this tile is issued
      const int r0 = first + t * TR;
      const int n = min(TR, count - t * TR);
      C->phase[s] = phase; C->row0[s] = r0; C->nrows[s] = n;
      const unsigned bytes = (unsigned)n * chunk_bytes;
      mbar_arrive_expect_tx(&C->full[s], bytes);
      bulk_g2s(smem + (size_t)s * tile_bytes, chunks + (size_t)r0 * chunk_bytes, bytes, &C->full[s], pol);
      if (++s == stages) { s = 0; par ^= 1u; }
    }
  }

  // ---- consumers
  const int gl = lane % LPR, q = lane / LPR;
  const unsigned gmask = group_mask<LPR>(lane);
  const unsigned long long keep = policy_evict_last();
  const double om1 = 1 - omega;
  int cur_phase = 0, s = 0;
  unsigned par = 0;
  double reg_old = 0.0;
  for (;;) {
    mbar_wait(&C->full[s], par);
    const int ph = C->phase[s];
    if (ph != cur_phase) {
      // Colour barrier.  Walk the phases this CTA leaves behind (it may have had no tile in some of them): arrive on each; a
      // boundary / regularisation phase additionally needs the interior colours of its sweep complete before the dot product
      // and every partial written before the new value of the regularisation unknown is formed.
      int waited = cur_phase - 1;                        // barrier cur_phase-1 was passed when this CTA entered cur_phase
      for (int p = cur_phase; p < ph; p++) {
        const bool is_b = bnd_phase >= 0 && (p % pps) == bnd_phase;
        if (is_b && reg.row >= 0) {
          if (waited < p - 1) { grid_wait(arrivals, p - 1, abort_flag, timeout_cycles); waited = p - 1; }
          const int per = (reg.len + (int)gridDim.x - 1) / (int)gridDim.x;
          const int lo = min(reg.len, (int)blockIdx.x * per), hi = min(reg.len, lo + per);
          double sum = 0.0;
          for (int k = lo + (int)threadIdx.x; k < hi; k += kConsumers) sum = __dadd_rn(sum, __dmul_rn(reg.val[k], x[reg.col[k]]));
          for (int o = 16; o > 0; o >>= 1) sum = __dadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o));
          if (lane == 0) red_scratch[warp] = sum;
          consumer_sync();
          if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < kConsumerWarps; w++) t = __dadd_rn(t, red_scratch[w]);
            reg.partial[(size_t)(p / pps) * gridDim.x + blockIdx.x] = t;
            reg_old = x[reg.row];                        // read before anybody can store the new value (that happens after barrier p)
          }
        }
        grid_arrive(arrivals, p);
        if (is_b && reg.row >= 0) {
          grid_wait(arrivals, p, abort_flag, timeout_cycles); waited = p;
          if (warp == 0) {                               // same fold order in every CTA: the value stored is identical
            double dot = 0.0;
            const double* part = reg.partial + (size_t)(p / pps) * gridDim.x;
            for (int i = lane; i < (int)gridDim.x; i += 32) dot = __dadd_rn(dot, ld_relaxed(part + i));
            for (int o = 16; o > 0; o >>= 1) dot = __dadd_rn(dot, __shfl_xor_sync(0xffffffffu, dot, o));
            if (lane == 0) {                             // SOR update of the regularisation row, the last row of the sweep (grid.cpp:117-143)
              double xi = -dot;
              xi = __dadd_rn(xi, b[reg.row]);
              xi = __dmul_rn(xi, omega / reg.diag);
              xi = __dadd_rn(xi, __dmul_rn(om1, reg_old));
              x[reg.row] = xi;
            }
          }
          consumer_sync();
        }
      }
      if (ph >= nphases) return;
      if (waited < ph - 1 && !(debug_flags & 1)) grid_wait(arrivals, ph - 1, abort_flag, timeout_cycles);
      cur_phase = ph;
    }
    const bool bnd = bnd_phase >= 0 && (ph % pps) == bnd_phase;
    const int n = C->nrows[s];
    const unsigned char* tile = smem + (size_t)s * tile_bytes;
    double v[ROWS][ITER], xx[ROWS][ITER], bi[ROWS], acc[ROWS];
    int c[ROWS][ITER];
#pragma unroll
    for (int h = 0; h < ROWS; h++) {
      const int lr = h * (kConsumerWarps * GPW) + warp * GPW + q;
      tile_row_fetch<LPR, ITER>(tile, chunk_bytes, W, lr, lr < n, gl, v[h], c[h]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&C->empty[s]);            // the chunks are in registers: the stage can be refilled
    if (++s == stages) { s = 0; par ^= 1u; }
    int tail[ROWS];
#pragma unroll
    for (int h = 0; h < ROWS; h++) {
      tail[h] = (gl == 0 && c[h][0] != -1) ? (c[h][0] >> 30) & 1 : 0;
#pragma unroll
      for (int t = 0; t < ITER; t++) {
        if (c[h][t] != -1) c[h][t] &= kColMask;
        xx[h][t] = c[h][t] >= 0 ? ldg_keep(x + ((debug_flags & 2) ? c[h][0] : c[h][t]), keep) : 0.0;
      }
    }
#pragma unroll
    for (int h = 0; h < ROWS; h++) bi[h] = (gl == 0 && c[h][0] >= 0) ? b[c[h][0]] : 0.0;   // slot 0 is the diagonal: its column is the row
#pragma unroll
    for (int h = 0; h < ROWS; h++) {
      double a = 0.0;
#pragma unroll
      for (int t = 0; t < ITER; t++) {
        if (t == 0 && gl == 0) continue;
        a = __dsub_rn(a, __dmul_rn(v[h][t], xx[h][t]));
      }
      acc[h] = a;
    }
    if (A.n_ovf) {                                       // Neumann-type grids only: tails of the few rows longer than the chunk
#pragma unroll
      for (int h = 0; h < ROWS; h++) {
        const int has = __shfl_sync(gmask, tail[h], (lane / LPR) * LPR);
        if (has) {
          const int row = __shfl_sync(gmask, c[h][0], (lane / LPR) * LPR);
          const int o = ovf_find(A, row);
          for (int k = A.ovf_ptr[o] + gl; k < A.ovf_ptr[o + 1]; k += LPR) acc[h] = __dsub_rn(acc[h], __dmul_rn(A.ovf_val[k], x[A.ovf_col[k]]));
        }
      }
    }
#pragma unroll
    for (int h = 0; h < ROWS; h++) acc[h] = group_sum<LPR>(acc[h], gmask);
    if (gl == 0) {
#pragma unroll
      for (int h = 0; h < ROWS; h++) {
        if (c[h][0] >= 0) {
          double xi = __dadd_rn(acc[h], bi[h]);
          if (bnd) {
            xi = xi / v[h][0];                           // x_c = (b_c - sum_{k != c} a_ck x_k) / a_cc, grid.cpp:96-100
          } else {
            xi = __dmul_rn(xi, omega / v[h][0]);
            xi = __dadd_rn(xi, __dmul_rn(om1, xx[h][0])); // xx[h][0] on lane 0 is x[row] before the update
          }
          x[c[h][0]] = xi;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// TMA-fed barrier-free sweep: the stream of k_sor_mc_tma with the synchronisation of k_sor_mc_flow (mmg_kernels.cu).
// Every sweep writes its own version xs[s+1] of the vector, whose swept rows start as the sentinel NaN; a row of colour c
// reads neighbours of a lower colour from xs[s+1] (bit 31 of the packed column) and everything else from xs[s], polling with
// ld.relaxed while it sees the sentinel -- the values are the ready flags, there is no barrier and no fence.  Tiles are taken
// in (sweep, colour, tile) order and every CTA works through its tiles in that order, so the lowest unfinished tile always has a
// resident owner whose operands are complete (cooperative launch, clock64 watchdog).  For the levels that are too small to
// amortise a colour barrier (<= 1.5M rows) and, with PEER, for a rank's row block of a partitioned level: a row next to a cut
// is also stored into the neighbour rank's copy of the same version over NVLink (st.relaxed.sys), where that rank's rows poll it.
// Per-row arithmetic and lane mapping of k_sor_mc_flow / k_sor_mc_packed: same bits.
// ------------------------------------------------------------------------------------------------
template <int LPR, int ITER, int ROWS, bool PEER>
__global__ void __launch_bounds__(kStreamThreads, 3) k_sor_mc_tma_flow(const unsigned char* __restrict__ chunks, unsigned chunk_bytes, int W,
                                                                    const int* __restrict__ phase_ptr, int pps, int iters, const double* __restrict__ b,
                                                                    double* xs, size_t stride, double omega, int* ctl, int stages, int dynamic,
                                                                    int* abort_flag, long long timeout_cycles, PeerSends peers, int l1_first) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int GPW = 32 / LPR;
  constexpr int TR = kConsumerWarps * GPW * ROWS;
  const unsigned tile_bytes = TR * chunk_bytes;
  RingCtl* C = ring_setup(smem, stages, tile_bytes);
  const int nphases = iters * pps;
  int* tickets = ctl;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == kConsumerWarps) {                          // ---- producer (as in k_sor_mc_tma)
    if (lane != 0) return;
    const unsigned long long pol = policy_evict_first();
    int s = 0;
    unsigned par = 0;
    int phase = 0;
    int ticket = dynamic ? atomicAdd(&tickets[0], 1) : (int)blockIdx.x;
    for (;;) {
      int t = ticket, first = 0, count = 0;
      while (phase < nphases) {
        const int c = phase % pps;
        first = phase_ptr[c];
        count = phase_ptr[c + 1] - first;
        if (t < (count + TR - 1) / TR) break;
        if (++phase < nphases) t = dynamic ? atomicAdd(&tickets[phase], 1) : (int)blockIdx.x;
      }
      mbar_wait(&C->empty[s], par ^ 1u);
      if (phase >= nphases) {
        C->phase[s] = nphases;
        mbar_arrive(&C->full[s]);
        return;
      }
      ticket = dynamic ? atomicAdd(&tickets[phase], 1) : t + (int)gridDim.x;
      const int r0 = first + t * TR;
      const int n = min(TR, count - t * TR);
      C->phase[s] = phase; C->row0[s] = r0; C->nrows[s] = n;
      const unsigned bytes = (unsigned)n * chunk_bytes;
      mbar_arrive_expect_tx(&C->full[s], bytes);
      bulk_g2s(smem + (size_t)s * tile_bytes, chunks + (size_t)r0 * chunk_bytes, bytes, &C->full[s], pol);
      if (++s == stages) { s = 0; par ^= 1u; }
    }
  }

  // ---- consumers
  const int gl = lane % LPR, q = lane / LPR;
  const unsigned gmask = group_mask<LPR>(lane);
  const double om1 = 1 - omega;
  const unsigned long long keep = policy_evict_last();
  const long long t_start = clock64();
  int s = 0;
  unsigned par = 0;
  for (;;) {
    mbar_wait(&C->full[s], par);
    const int ph = C->phase[s];
    if (ph >= nphases) return;
    const int it = ph / pps;
    const double* xold = xs + (size_t)it * stride;
    double* xnew = xs + (size_t)(it + 1) * stride;
    const int n = C->nrows[s];
    const unsigned char* tile = smem + (size_t)s * tile_bytes;
    double v[ROWS][ITER], xx[ROWS][ITER], bi[ROWS], acc[ROWS];
    int c[ROWS][ITER];
#pragma unroll
    for (int h = 0; h < ROWS; h++) {
      const int lr = h * (kConsumerWarps * GPW) + warp * GPW + q;
      tile_row_fetch<LPR, ITER>(tile, chunk_bytes, W, lr, lr < n, gl, v[h], c[h]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&C->empty[s]);
    if (++s == stages) { s = 0; par ^= 1u; }
    unsigned pend = 0, newer = 0;
#pragma unroll
    for (int h = 0; h < ROWS; h++)
#pragma unroll
      for (int t = 0; t < ITER; t++) {
        const int raw = c[h][t];
        if (raw != -1) {
          c[h][t] = raw & kColMask;
          if (raw < 0) newer |= 1u << (h * ITER + t);
          const double* src = (raw < 0 ? xnew : xold) + c[h][t];
          // First probe through L1: within one launch every entry of a version goes sentinel -> value exactly once, so a cached
          // line can hold a stale SENTINEL (then the strong poll below fetches the value from L2) but never a stale value.
          xx[h][t] = l1_first ? ldg_keep(src, keep) : PEER ? ld_relaxed_sys(src) : ld_relaxed(src);
          if (is_sentinel(xx[h][t])) pend |= 1u << (h * ITER + t);
        } else xx[h][t] = 0.0;
      }
#pragma unroll
    for (int h = 0; h < ROWS; h++) bi[h] = (gl == 0 && c[h][0] != -1) ? b[c[h][0]] : 0.0;
    bool aborted = false;
    unsigned spins = 0;
    while (__any_sync(0xffffffffu, pend != 0)) {         // rare: an operand of this tile is still being computed
#pragma unroll
      for (int h = 0; h < ROWS; h++)
#pragma unroll
        for (int t = 0; t < ITER; t++)
          if (pend & (1u << (h * ITER + t))) {
            const double* src = ((newer >> (h * ITER + t)) & 1u ? xnew : xold) + c[h][t];
            xx[h][t] = PEER ? ld_relaxed_sys(src) : ld_relaxed(src);
            if (!is_sentinel(xx[h][t])) pend &= ~(1u << (h * ITER + t));
          }
      if ((++spins & 0x3f) == 0 && (*(volatile int*)abort_flag || clock64() - t_start > timeout_cycles)) { atomicExch(abort_flag, 1); aborted = true; break; }
    }
#pragma unroll
    for (int h = 0; h < ROWS; h++) {
      double a = 0.0;
#pragma unroll
      for (int t = 0; t < ITER; t++) {
        if (t == 0 && gl == 0) continue;
        a = __dsub_rn(a, __dmul_rn(v[h][t], xx[h][t]));
      }
      acc[h] = a;
    }
#pragma unroll
    for (int h = 0; h < ROWS; h++) acc[h] = group_sum<LPR>(acc[h], gmask);
    if (gl == 0) {
#pragma unroll
      for (int h = 0; h < ROWS; h++) {
        if (c[h][0] != -1) {
          double xi = __dadd_rn(acc[h], bi[h]);
          xi = __dmul_rn(xi, omega / v[h][0]);
          xi = __dadd_rn(xi, __dmul_rn(om1, xx[h][0]));
          const int row = c[h][0];
          if (aborted) xi = 0.0;                          // on abort: unblock everyone behind us
          st_relaxed(xnew + row, xi);
          if (PEER) {
            if (peers.n > 0 && row >= peers.lo[0] && row < peers.hi[0]) st_relaxed_sys(peers.base[0] + (size_t)(it + 1) * stride + row, xi);
            if (peers.n > 1 && row >= peers.lo[1] && row < peers.hi[1]) st_relaxed_sys(peers.base[1] + (size_t)(it + 1) * stride + row, xi);
          }
        }
      }
    }
    if (aborted) return;
  }
}

// ------------------------------------------------------------------------------------------------
// SpMV-class operators over the natural-order row chunks: y = op(A, x) for rows [row0, row0 + nrows) (ops as in k_spmv2).
// Static round-robin tiles (no inter-CTA dependency: a plain launch), same ring.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void consumer_sum2(double& a, double& b, double* scratch) {   // sum over the 256 consumer threads, valid in thread 0
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { scratch[2 * w] = a; scratch[2 * w + 1] = b; }
  consumer_sync();
  if (w == 0) {
    a = l < kConsumerWarps ? scratch[2 * l] : 0.0;
    b = l < kConsumerWarps ? scratch[2 * l + 1] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
  }
}

template <int LPR, int ITER, int ROWS>
__global__ void __launch_bounds__(kStreamThreads, 3) k_spmv_tma(HybView A, const double* x, const double* __restrict__ b, double* y,
                                                             const unsigned char* __restrict__ rowflag, int op, int mask_dirichlet, int mask_neumann,
                                                             double* __restrict__ partial, int row0, int nrows, int stages) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ double red_scratch[2 * kConsumerWarps];
  constexpr int GPW = 32 / LPR;
  constexpr int TR = kConsumerWarps * GPW * ROWS;
  const unsigned chunk_bytes = (unsigned)A.chunk_bytes;
  const unsigned tile_bytes = TR * chunk_bytes;
  RingCtl* C = ring_setup(smem, stages, tile_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = (nrows + TR - 1) / TR;

  if (warp == kConsumerWarps) {                          // ---- producer
    if (lane != 0) return;
    const unsigned long long pol = policy_evict_first();
    int s = 0;
    unsigned par = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      mbar_wait(&C->empty[s], par ^ 1u);
      const int r0 = row0 + tile * TR;
      const unsigned bytes = (unsigned)min(TR, nrows - tile * TR) * chunk_bytes;
      mbar_arrive_expect_tx(&C->full[s], bytes);
      bulk_g2s(smem + (size_t)s * tile_bytes, A.chunks + (size_t)r0 * chunk_bytes, bytes, &C->full[s], pol);
      if (++s == stages) { s = 0; par ^= 1u; }
    }
    return;
  }

  const int gl = lane % LPR, q = lane / LPR;
  const unsigned gmask = group_mask<LPR>(lane);
  const unsigned long long keep = policy_evict_last();
  double num = 0.0, den = 0.0;
  int s = 0;
  unsigned par = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    mbar_wait(&C->full[s], par);
    const int n = min(TR, nrows - tile * TR);
    const unsigned char* tb = smem + (size_t)s * tile_bytes;
    double v[ROWS][ITER], xx[ROWS][ITER], acc[ROWS];
    int c[ROWS][ITER], row[ROWS];
    bool valid[ROWS];
#pragma unroll
    for (int h = 0; h < ROWS; h++) {
      const int lr = h * (kConsumerWarps * GPW) + warp * GPW + q;
      valid[h] = lr < n;
      row[h] = row0 + tile * TR + lr;
      tile_row_fetch<LPR, ITER>(tb, chunk_bytes, A.W, lr, valid[h], gl, v[h], c[h]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&C->empty[s]);
    if (++s == stages) { s = 0; par ^= 1u; }
#pragma unroll
    for (int h = 0; h < ROWS; h++)
#pragma unroll
      for (int t = 0; t < ITER; t++) xx[h][t] = c[h][t] >= 0 ? ldg_keep(x + c[h][t], keep) : 0.0;
    // the epilogue's own operands (row flag, b_i or the old y_i) are fetched with the gathers, not after the reduction
    int flag_[ROWS];
    double aux_[ROWS];
#pragma unroll
    for (int h = 0; h < ROWS; h++) {
      const bool lead = valid[h] && gl == 0;
      flag_[h] = (lead && rowflag) ? rowflag[row[h]] : 0;
      aux_[h] = !lead ? 0.0 : op == OP_RESID ? b[row[h]] : op == OP_PROLONG ? y[row[h]] : 0.0;
    }
#pragma unroll
    for (int h = 0; h < ROWS; h++) {
      double a = 0.0;
#pragma unroll
      for (int t = 0; t < ITER; t++) a = __dadd_rn(a, __dmul_rn(v[h][t], xx[h][t]));
      acc[h] = a;
    }
    if (A.n_ovf) {                                       // rows longer than W (implicit-Neumann fill-in) keep their tail in a small CSR
#pragma unroll
      for (int h = 0; h < ROWS; h++) {
        if (valid[h] && A.len[row[h]] > A.W) {
          const int o = ovf_find(A, row[h]);
          for (int k = A.ovf_ptr[o] + gl; k < A.ovf_ptr[o + 1]; k += LPR) acc[h] = __dadd_rn(acc[h], __dmul_rn(A.ovf_val[k], x[A.ovf_col[k]]));
        }
      }
    }
#pragma unroll
    for (int h = 0; h < ROWS; h++) acc[h] = group_sum<LPR>(acc[h], gmask);
#pragma unroll
    for (int h = 0; h < ROWS; h++) {
      if (valid[h] && gl == 0) {
        const int flag = flag_[h];
        if (op == OP_SPMV) {
          y[row[h]] = acc[h];
        } else if (op == OP_RESID) {
          const double bi = aux_[h];
          double t = __dsub_rn(bi, acc[h]);
          if (flag == 1) t = 0.0;
          if (y) y[row[h]] = t;
          num += fabs(t);
          den += fabs(bi);
        } else if (op == OP_PROLONG) {
          if (!(mask_dirichlet && flag == 1)) y[row[h]] = __dadd_rn(aux_[h], acc[h]);
        } else {
          double t = acc[h];
          if (flag == 1) t = 0.0;
          if (mask_neumann && flag == 2) t = 0.0;
          y[row[h]] = t;
        }
      }
    }
  }
  if (partial) {
    consumer_sum2(num, den, red_scratch);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = num; partial[2 * blockIdx.x + 1] = den; }
  }
}

int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }

// (lanes per row, entries per lane) table shared with the register-fed kernels
template <class F>
bool dispatch_lanes(int W, int prefer_lpr, F&& f) {
  const int lpr = prefer_lpr ? prefer_lpr : (W >= 48 ? 32 : (W >= 24 ? 16 : 8));
  const int iter = (W + lpr - 1) / lpr;
#define MMG_CASE(L_, I_) if (lpr == L_ && iter == I_) { f(std::integral_constant<int, L_>(), std::integral_constant<int, I_>()); return true; }
  MMG_CASE(32, 2) MMG_CASE(32, 3) MMG_CASE(32, 4)
  MMG_CASE(16, 2) MMG_CASE(16, 3) MMG_CASE(16, 4) MMG_CASE(16, 5)
  MMG_CASE(8, 1) MMG_CASE(8, 2) MMG_CASE(8, 3) MMG_CASE(8, 4) MMG_CASE(8, 5)
#undef MMG_CASE
  return false;
}

struct RingShape { int stages; size_t smem; int ctas_per_sm; };

// Ring depth: as many stages as fit the per-CTA shared-memory budget (ctas_per_sm CTAs share an SM's 227 KB; the rest stays L1 for the gathers)
RingShape ring_shape(unsigned tile_bytes, int ctas_per_sm, int want_stages, int smem_kb) {
  const size_t budget = (size_t)smem_kb * 1024 / (size_t)ctas_per_sm;
  int stages = (int)((budget - sizeof(RingCtl) - 128) / tile_bytes);
  if (want_stages > 0) stages = std::min(stages, want_stages);
  stages = std::max(2, std::min(stages, kMaxStages));
  return RingShape{stages, (size_t)stages * tile_bytes + sizeof(RingCtl), ctas_per_sm};
}

int sm_count(int device) {
  static int cached[64] = {0};
  if (device < 64 && cached[device]) return cached[device];
  int n = 0;
  MMG_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
  if (device < 64) cached[device] = n;
  return n;
}

template <class K>
void allow_smem(K kern, size_t smem, int ctas_per_sm) {
  // per-kernel attributes; setting them again is cheap and idempotent.  The carve-out hint keeps what the ring does not
  // need as L1 for the gathers (the driver would otherwise size shared memory for the occupancy limit, not for our grid).
  MMG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int pct = std::min(100, (int)((100 * (size_t)ctas_per_sm * (smem + 2048) + 228 * 1024 - 1) / (228 * 1024)));
  MMG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
}

}  // namespace

// All colours of all props.iters sweeps of the multicolour smoother over g.mc_chunks.  Returns false when the stencil
// width has no instantiation (the caller falls back to the register-fed kernel).
bool stream_sor_mc(Grid& g) {
  const HybMatrix& L = g.Lap;
  // lanes per row on the big levels: 8 up to n=37 (as k_sor_mc_packed), 16 for the wide stencils (n=70: 4803 vs 3837 GB/s with a warp per row)
  int prefer = g.A < 200000 ? 0 : (L.W > 16 && L.W <= 40) ? 8 : (L.W > 40 && L.W <= 80) ? 16 : 0;
  if (env_int("MMG_TMA_LPR", 0)) prefer = env_int("MMG_TMA_LPR", 0);
  const int rows_pref = env_int("MMG_TMA_ROWS", 2);
  return dispatch_lanes(L.W, prefer, [&](auto Lc, auto I) {
    constexpr int LPR = decltype(Lc)::value, ITER = decltype(I)::value;
    const int rows_used = rows_pref >= 2 ? 2 : 1;
    const unsigned tile_bytes = (unsigned)(kConsumerWarps * (32 / LPR) * rows_used * L.chunk_bytes);
    const int sms = sm_count(g.device);
    auto kern = rows_used == 2 ? k_sor_mc_tma<LPR, ITER, 2> : k_sor_mc_tma<LPR, ITER, 1>;
    // measured on the 4M-row level, n=37 (profiles/r02_tma_sweep.txt): 3 CTAs x 2 stages x 64-row tiles 4789 GB/s; 2 CTAs x 3 stages 4463;
    // 32-row tiles 3204 (2 CTAs) ... 4403 (4 CTAs): the consumers are latency bound on the gathers, so resident warps count
    RingShape rs = ring_shape(tile_bytes, env_int("MMG_TMA_CTAS", 3), env_int("MMG_TMA_STAGES", 0), env_int("MMG_TMA_SMEM_KB", 180));
    allow_smem(kern, rs.smem, rs.ctas_per_sm);
    int blocks_per_sm = 0;
    MMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kStreamThreads, rs.smem));
    MMG_REQUIRE(blocks_per_sm >= 1, MMG_ERR_CUDA, "k_sor_mc_tma does not fit an SM");
    const int blocks = std::min(blocks_per_sm, rs.ctas_per_sm) * sms;
    const int pps = (int)g.mc_colour_ptr.size() - 1;       // phases per sweep: interior colours (+ the boundary / regularisation phase)
    const int nphases = pps * g.props.iters;
    if (g.mc_ctl.n < (size_t)2 * nphases) g.mc_ctl.alloc((size_t)2 * nphases);
    MMG_CUDA(cudaMemsetAsync(g.mc_ctl.p, 0, sizeof(int) * 2 * nphases, g.stream));
    if (L.reg_row >= 0 && g.mc_reg_partial.n < (size_t)g.props.iters * blocks) g.mc_reg_partial.alloc((size_t)g.props.iters * blocks);
    const unsigned char* chunks = g.mc_chunks.p;
    unsigned cb = (unsigned)L.chunk_bytes;
    int W = L.W;
    const int* cp = g.mc_colour_ptr_dev.p;
    int bnd_phase = g.mc_bnd_phase, iters = g.props.iters, ppsv = pps;
    const double* b = g.b.p;
    double* x = g.x.p;
    double omega = g.props.omega;
    int* ctl = g.mc_ctl.p;
    int stages = rs.stages, dynamic = env_int("MMG_TMA_DYNAMIC", 1);
    int* abortp = g.abort_flag.p;
    long long timeout = 4000000000ll;                    // ~2 s of SM clocks per colour barrier
    int debug_flags = env_int("MMG_TMA_DEBUG", 0);      // timing decomposition only (1: no colour barrier, 2: no gathers): results invalid
    HybView A = L.view();
    RegRow reg{L.reg_col.p, L.reg_val.p, L.reg_len, L.reg_row, L.reg_diag, g.mc_reg_partial.p};
    // contiguous share of a phase's tiles per CTA before the ticket-dealt rest, in 1/8 of the even share (profiles/r02_static_share.txt:
    // 4836 -> 4920 GB/s at n=37, 4570 -> 4610 at n=70 with the whole even share static; handing the shares of the CTAs that share
    // an SM to neighbours along the curve as well changes nothing, r02_static_share3.txt)
    int static_8ths = std::max(0, std::min(8, env_int("MMG_TMA_STATIC_8THS", 8)));
    void* args[] = {&chunks, &cb, &W, &cp, &ppsv, &bnd_phase, &iters, &b, &x, &omega, &ctl, &stages, &dynamic, &abortp, &timeout, &debug_flags, &A, &reg, &static_8ths};
    note_kernel(g, "k_sor_mc_tma", LPR, ITER, rows_used);
    MMG_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(blocks), dim3(kStreamThreads), args, rs.smem, g.stream));
  });
}

// Barrier-free TMA-fed sweep over the versioned vectors xs (iters+1 versions of `stride` doubles, version 0 = values_, swept rows
// of the later versions = sentinel): all phases of the call in one cooperative launch.  `peers` non-null: rows next to a cut are
// mirrored into the neighbour ranks' vectors.  False when the stencil width has no instantiation.
bool stream_sor_mc_flow(Grid& g, double* xs, size_t stride, const PeerSends* peers) {
  const HybMatrix& L = g.Lap;
  int prefer = g.A < 200000 ? 0 : (L.W > 16 && L.W <= 40) ? 8 : (L.W > 40 && L.W <= 80) ? 16 : 0;
  if (env_int("MMG_TMA_LPR", 0)) prefer = env_int("MMG_TMA_LPR", 0);
  // measured (profiles/r02_tma_sweep7.txt): 1M rows: 64-row tiles x 3 CTAs 3620 GB/s, 32-row tiles x 4 CTAs 3200 (register-fed
  // k_sor_mc_flow: 2760); 250k rows: 1630 vs 1890 (1720) -- the smaller level wants more, smaller tiles in flight
  const int swept = g.mc_colour_ptr.back();
  const int rows_pref = env_int("MMG_TMAFLOW_ROWS", swept >= 500000 ? 2 : 1);
  const int ctas_pref = env_int("MMG_TMAFLOW_CTAS", swept >= 500000 ? 3 : 4);
  return dispatch_lanes(L.W, prefer, [&](auto Lc, auto I) {
    constexpr int LPR = decltype(Lc)::value, ITER = decltype(I)::value;
    const int rows_used = rows_pref >= 2 ? 2 : 1;
    void* kern = peers ? (rows_used == 2 ? (void*)k_sor_mc_tma_flow<LPR, ITER, 2, true> : (void*)k_sor_mc_tma_flow<LPR, ITER, 1, true>)
                       : (rows_used == 2 ? (void*)k_sor_mc_tma_flow<LPR, ITER, 2, false> : (void*)k_sor_mc_tma_flow<LPR, ITER, 1, false>);
    const unsigned tile_bytes = (unsigned)(kConsumerWarps * (32 / LPR) * rows_used * L.chunk_bytes);
    const int sms = sm_count(g.device);
    RingShape rs = ring_shape(tile_bytes, ctas_pref, env_int("MMG_TMAFLOW_STAGES", 0), env_int("MMG_TMAFLOW_SMEM_KB", 180));
    MMG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs.smem));
    int blocks_per_sm = 0;
    MMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, (const void*)kern, kStreamThreads, rs.smem));
    MMG_REQUIRE(blocks_per_sm >= 1, MMG_ERR_CUDA, "k_sor_mc_tma_flow does not fit an SM");
    const int pps = (int)g.mc_colour_ptr.size() - 1;
    const int TR = kConsumerWarps * (32 / LPR) * rows_used;
    int maxtiles = 1;
    for (int c = 0; c < pps; c++) maxtiles = std::max(maxtiles, (g.mc_colour_ptr[c + 1] - g.mc_colour_ptr[c] + TR - 1) / TR);
    const int blocks = std::min(std::min(blocks_per_sm, rs.ctas_per_sm) * sms, maxtiles);
    const int nphases = pps * g.props.iters;
    if (g.mc_ctl.n < (size_t)2 * nphases) g.mc_ctl.alloc((size_t)2 * nphases);
    MMG_CUDA(cudaMemsetAsync(g.mc_ctl.p, 0, sizeof(int) * 2 * nphases, g.stream));
    const unsigned char* chunks = g.mc_chunks.p;
    unsigned cb = (unsigned)L.chunk_bytes;
    int W = L.W;
    const int* cp = g.mc_colour_ptr_dev.p;
    int iters = g.props.iters, ppsv = pps;
    const double* b = g.b.p;
    size_t st = stride;
    double omega = g.props.omega;
    int* ctl = g.mc_ctl.p;
    // small levels: a colour is a handful of tiles, so a ticket per tile (two dependent atomics per phase and CTA) would be the
    // critical path; tiles are dealt round-robin instead and the producer runs as far ahead as the ring allows
    int stages = rs.stages, dynamic = env_int("MMG_TMAFLOW_DYNAMIC", swept >= 150000 ? 1 : 0);
    int* abortp = g.abort_flag.p;
    long long timeout = 6000000000ll;
    PeerSends ps{};
    if (peers) ps = *peers;
    int l1_first = env_int("MMG_TMAFLOW_L1", 1);
    void* args[] = {&chunks, &cb, &W, &cp, &ppsv, &iters, &b, &xs, &st, &omega, &ctl, &stages, &dynamic, &abortp, &timeout, &ps, &l1_first};
    note_kernel(g, peers ? "k_sor_mc_tma_flow_peer" : "k_sor_mc_tma_flow", LPR, ITER, rows_used);
    MMG_CUDA(cudaLaunchCooperativeKernel(kern, dim3(blocks), dim3(kStreamThreads), args, rs.smem, g.stream));
  });
}

// y = op(M, x) on rows [row0, row0+nrows) through the TMA ring; false when the stencil width has no instantiation
bool stream_spmv(const HybMatrix& M, const double* x, const double* b, double* y, const unsigned char* rowflag, int op, int mask_d, int mask_n,
                 double* partial, int* nblocks_out, int device, cudaStream_t s, int row0, int nrows) {
  int prefer = nrows < 200000 ? 0 : (M.W > 16 && M.W <= 40) ? 8 : (M.W > 40 && M.W <= 80) ? 16 : 0;
  if (env_int("MMG_SPMV_TMA_LPR", 0)) prefer = env_int("MMG_SPMV_TMA_LPR", 0);
  const int rows_pref = env_int("MMG_SPMV_TMA_ROWS", 2);
  return dispatch_lanes(M.W, prefer, [&](auto Lc, auto I) {
    constexpr int LPR = decltype(Lc)::value, ITER = decltype(I)::value;
    const int rows_used = rows_pref >= 2 ? 2 : 1;
    auto kern = rows_used == 2 ? k_spmv_tma<LPR, ITER, 2> : k_spmv_tma<LPR, ITER, 1>;
    const int TR = kConsumerWarps * (32 / LPR) * rows_used;
    const unsigned tile_bytes = (unsigned)(TR * M.chunk_bytes);
    const int sms = sm_count(device);
    RingShape rs = ring_shape(tile_bytes, env_int("MMG_SPMV_TMA_CTAS", 3), env_int("MMG_SPMV_TMA_STAGES", 0), env_int("MMG_SPMV_TMA_SMEM_KB", 180));
    allow_smem(kern, rs.smem, rs.ctas_per_sm);
    const int ntiles = (nrows + TR - 1) / TR;
    const int blocks = std::max(1, std::min(rs.ctas_per_sm * sms, ntiles));
    if (nblocks_out) *nblocks_out = blocks;
    note_kernel_slot(1, "k_spmv_tma", LPR, ITER, rows_used);
    kern<<<blocks, kStreamThreads, rs.smem, s>>>(M.view(), x, b, y, rowflag, op, mask_d, mask_n, partial, row0, nrows, rs.stages);
    MMG_CUDA(cudaGetLastError());
  });
}

}  // namespace mmg
