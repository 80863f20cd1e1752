// Solve-path kernels (sm_100a): residual / interpolation SpMV over the row-chunk HYB layout,
// lexicographic SOR as a dependency-DAG sweep (value-carried ready flags in L2), multicolour SOR,
// Neumann boundary evaluation, the regularisation-row reductions and the small scatter ops.
//
// Reference semantics restated by each kernel are cited inline (paths relative to
// /root/reference/MeshlessPoisson/).  All arithmetic is fp64; multiplications and additions of the
// row sums are issued as separate roundings (__dmul_rn/__dadd_rn) because the reference is built
// without FMA contraction.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstring>

#include "mmg_internal.hpp"
#include "mmg_device.cuh"

#ifndef MMG_FAST_ROWS
#define MMG_FAST_ROWS 4
#endif

namespace mmg {

namespace {

// ------------------------------------------------------------------------------------------------
// y = op(A, x): one LPR-lane group per row, lanes stride over the row chunk.
//   OP_SPMV      y = A x                                  (Eigen sparse*dense, grid.cpp:148, multigrid.cpp:81,102)
//   OP_RESID     y = b - A x, Dirichlet rows -> 0          (Grid::residual grid.cpp:147-151); optional |.|_1 partials
//   OP_PROLONG   y += A x except masked Dirichlet rows     (multigrid.cpp:102-106)
//   OP_RESTRICT  y = A x, Dirichlet rows -> 0, Neumann rows -> 0 when mask_neumann (multigrid.cpp:81-86)
// ------------------------------------------------------------------------------------------------
template <int LPR>
__global__ void __launch_bounds__(kBlock) k_spmv(HybView A, const double* __restrict__ x, const double* __restrict__ b, double* y,
                                                 const unsigned char* __restrict__ rowflag, int op, int mask_dirichlet, int mask_neumann,
                                                 double* __restrict__ partial) {
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR;
  const unsigned gmask = group_mask<LPR>(lane);
  constexpr int GPW = 32 / LPR;  // groups per warp
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  double num = 0.0, den = 0.0;
  for (int row0 = warp * GPW; row0 < A.rows; row0 += nwarps * GPW) {
    const int row = row0 + lane / LPR;
    const bool valid = row < A.rows;
    double acc = 0.0;
    if (valid) {
      const int len = A.len[row];
      const double* __restrict__ v = row_val(A, row);
      const int* __restrict__ c = row_col(A, row);
      const int m = len < A.W ? len : A.W;
      for (int k = gl; k < m; k += LPR) acc = __dadd_rn(acc, __dmul_rn(v[k], __ldg(x + c[k])));
      if (len > A.W) {
        const int o = ovf_find(A, row);
        for (int k = A.ovf_ptr[o] + gl; k < A.ovf_ptr[o + 1]; k += LPR) acc = __dadd_rn(acc, __dmul_rn(A.ovf_val[k], __ldg(x + A.ovf_col[k])));
      }
    }
    acc = group_sum<LPR>(acc, gmask);
    if (valid && gl == 0) {
      const int flag = rowflag ? rowflag[row] : 0;
      if (op == OP_SPMV) {
        y[row] = acc;
      } else if (op == OP_RESID) {
        const double bi = b[row];
        double t = __dsub_rn(bi, acc);
        if (flag == 1) t = 0.0;
        if (y) y[row] = t;
        num += fabs(t);
        den += fabs(bi);
      } else if (op == OP_PROLONG) {
        if (!(mask_dirichlet && flag == 1)) y[row] = __dadd_rn(y[row], acc);
      } else {  // OP_RESTRICT
        double t = acc;
        if (flag == 1) t = 0.0;
        if (mask_neumann && flag == 2) t = 0.0;
        y[row] = t;
      }
    }
  }
  if (partial) {
    block_sum2(num, den);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = num; partial[2 * blockIdx.x + 1] = den; }
  }
}

// partial sums of the regularisation row's off-diagonal dot product: sum_j val[j]*x[col[j]] (grid.cpp:570-576)
__global__ void __launch_bounds__(kBlock) k_regdot(const int* __restrict__ col, const double* __restrict__ val, int len, const double* x,
                                                   double* __restrict__ partial) {
  double s = 0.0, dummy = 0.0;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < len; j += gridDim.x * blockDim.x) s += val[j] * x[col[j]];
  block_sum2(s, dummy);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// single block: fold norm partials (+ regularisation row) into r[N] and *ratio = |r|_1/|b|_1 (multigrid.cpp:112-115)
__global__ void __launch_bounds__(kBlock) k_finish_residual(const double* __restrict__ norm_partial, int n_norm, const double* __restrict__ reg_partial,
                                                            int n_reg, int reg_row, double reg_diag, const double* __restrict__ x,
                                                            const double* __restrict__ b, double* r, double* ratio) {
  double num = 0.0, den = 0.0;
  for (int i = threadIdx.x; i < n_norm; i += blockDim.x) { num += norm_partial[2 * i]; den += norm_partial[2 * i + 1]; }
  block_sum2(num, den);
  __syncthreads();
  double dot = 0.0, dummy = 0.0;
  if (reg_row >= 0) {
    for (int i = threadIdx.x; i < n_reg; i += blockDim.x) dot += reg_partial[i];
    block_sum2(dot, dummy);
  }
  if (threadIdx.x == 0) {
    if (reg_row >= 0) {
      const double t = b[reg_row] - (dot + reg_diag * x[reg_row]);
      if (r) r[reg_row] = t;
      num += fabs(t);
      den += fabs(b[reg_row]);
    }
    if (ratio) *ratio = num / den;
  }
}

// single block: SOR update of the regularisation row (it is the last row of the sweep, grid.cpp:117-143)
__global__ void __launch_bounds__(kBlock) k_finish_sor_reg(const double* __restrict__ reg_partial, int n_reg, int reg_row, double reg_diag, double omega,
                                                           const double* __restrict__ b, const double* x_old, double* x_new) {
  double dot = 0.0, dummy = 0.0;
  for (int i = threadIdx.x; i < n_reg; i += blockDim.x) dot += reg_partial[i];
  block_sum2(dot, dummy);
  if (threadIdx.x == 0) {
    double xi = -dot;
    xi += b[reg_row];
    xi *= omega / reg_diag;
    xi += (1 - omega) * x_old[reg_row];
    x_new[reg_row] = xi;
  }
}

// ------------------------------------------------------------------------------------------------
// Lexicographic SOR sweep (Grid::sor grid.cpp:117-143) as a dependency-DAG sweep.
// x_new is pre-filled by k_sweep_init: rows the sweep skips carry their old value, rows it visits
// carry a sentinel.  Row i reads x_new[c] for c<i (spinning while it is the sentinel) and x_old[c]
// for c>i, so the result is the sequential in-place sweep.  Groups take rows round-robin in
// increasing order and the launch is cooperative (all CTAs resident), hence the lowest unfinished
// row always has a running owner whose dependencies are done: no deadlock.  A clock64 watchdog
// turns any stall into MMG_ERR_TIMEOUT instead of a hung GPU.
// ------------------------------------------------------------------------------------------------
// (k_sor_lex_exact below is the kernel; a first-generation variant with reordered group sums was removed in round 2: it
// deadlocked on Neumann-type grids beyond ~200k rows and was never the default.)
__global__ void __launch_bounds__(kBlock) k_sweep_init(const unsigned char* __restrict__ rowflag, const double* __restrict__ x_old, double* x_new,
                                                       int rows, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) x_new[i] = rowflag[i] != 0 ? x_old[i] : __longlong_as_double((long long)kSentinelBits);
  else if (i < total) x_new[i] = x_old[i];
}

// Multicolour SOR phase: all rows of one colour, in place (same row update as grid.cpp:122-141).
template <int LPR>
__global__ void __launch_bounds__(kBlock) k_sor_mc(HybView A, const int* __restrict__ rows_list, int count, const double* __restrict__ b, double* x,
                                                   double omega) {
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR;
  const unsigned gmask = group_mask<LPR>(lane);
  constexpr int GPW = 32 / LPR;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i0 = warp * GPW; i0 < count; i0 += nwarps * GPW) {
    const int i = i0 + lane / LPR;
    const bool valid = i < count;
    double acc = 0.0, diag = 0.0;
    int row = 0;
    if (valid) {
      row = rows_list[i];
      const int len = A.len[row];
      const double* __restrict__ v = row_val(A, row);
      const int* __restrict__ c = row_col(A, row);
      const int m = len < A.W ? len : A.W;
      for (int k = gl; k < m; k += LPR) {
        const double a = v[k];
        if (k == 0) diag = a;
        else acc = __dsub_rn(acc, __dmul_rn(a, x[c[k]]));
      }
      if (len > A.W) {
        const int o = ovf_find(A, row);
        for (int k = A.ovf_ptr[o] + gl; k < A.ovf_ptr[o + 1]; k += LPR) acc = __dsub_rn(acc, __dmul_rn(A.ovf_val[k], x[A.ovf_col[k]]));
      }
    }
    acc = group_sum<LPR>(acc, gmask);
    if (valid && gl == 0) {
      double xi = __dadd_rn(acc, b[row]);
      xi = __dmul_rn(xi, omega / diag);
      xi = __dadd_rn(xi, __dmul_rn(1 - omega, x[row]));
      x[row] = xi;
    }
  }
}

// Grid::bound_eval_neumann grid.cpp:73-103: x_c = (b_c - sum_{k!=c} a_ck x_k)/a_cc for every Neumann node.
// Neumann rows reference only interior nodes and themselves (kNN exclusion rule grid.cpp:236,244; checked at
// upload), so the list can be evaluated in parallel.
template <int LPR>
__global__ void __launch_bounds__(kBlock) k_bound_eval(HybView A, const int* __restrict__ nodes, int count, const double* __restrict__ b, double* x) {
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR;
  const unsigned gmask = group_mask<LPR>(lane);
  constexpr int GPW = 32 / LPR;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i0 = warp * GPW; i0 < count; i0 += nwarps * GPW) {
    const int i = i0 + lane / LPR;
    const bool valid = i < count;
    double acc = 0.0, diag = 0.0;
    int row = 0;
    if (valid) {
      row = nodes[i];
      const int len = A.len[row];
      const double* __restrict__ v = row_val(A, row);
      const int* __restrict__ c = row_col(A, row);
      const int m = len < A.W ? len : A.W;
      for (int k = gl; k < m; k += LPR) {
        const double a = v[k];
        if (k == 0) diag = a;
        else acc = __dsub_rn(acc, __dmul_rn(x[c[k]], a));
      }
      if (len > A.W) {
        const int o = ovf_find(A, row);
        for (int k = A.ovf_ptr[o] + gl; k < A.ovf_ptr[o + 1]; k += LPR) acc = __dsub_rn(acc, __dmul_rn(x[A.ovf_col[k]], A.ovf_val[k]));
      }
    }
    acc = group_sum<LPR>(acc, gmask);
    if (valid && gl == 0) x[row] = __dadd_rn(b[row], acc) / diag;
  }
}

// ================================================================================================
// Reference-order ("exact") variants.  The reference accumulates every row sum sequentially in
// ascending column order (Eigen's row loop, Grid::sor, Grid::bound_eval_neumann).  A V-cycle
// residual history can only be reproduced to 1e-10 *relative per cycle* all the way down to the
// convergence floor if the sums are rounded identically, so the parity mode keeps that order:
// lanes fetch a row's entries coalesced and form the products in parallel, then the warp folds
// the products one by one in column order (a shuffle broadcast feeding a serial DADD chain).
// Results are bit-identical to the oracle for Dirichlet grids.
// ================================================================================================

// s <- s (+/-) p_k for the 32 products held one per lane, in lane order, first `cnt` lanes only.
// `dcol` handling: when a diagonal product is pending (diag-first storage) it is inserted just
// before the first entry whose column exceeds `row`, i.e. at its ascending-column position.
__device__ __forceinline__ double fold_in_order(double s, double p, int col, int cnt, bool subtract, bool& diag_pending, double pdiag, int row) {
  __syncwarp();
#pragma unroll
  for (int l = 0; l < 32; l++) {
    if (l < cnt) {
      const double pl = __shfl_sync(0xffffffffu, p, l);
      const int cl = __shfl_sync(0xffffffffu, col, l);
      if (diag_pending && cl > row) { s = subtract ? __dsub_rn(s, pdiag) : __dadd_rn(s, pdiag); diag_pending = false; }
      s = subtract ? __dsub_rn(s, pl) : __dadd_rn(s, pl);
    }
  }
  return s;
}

// in-order row dot product; all lanes return the same value.  skip_diag: SOR-style sums leave the diagonal out.
__device__ __forceinline__ double row_dot_in_order(const HybView& A, int row, const double* x, bool diag_first, bool skip_diag, bool subtract, double s0,
                                                   double* diag_out) {
  const int lane = threadIdx.x & 31;
  const int len = A.len[row];
  const double* __restrict__ v = row_val(A, row);
  const int* __restrict__ c = row_col(A, row);
  const int m = len < A.W ? len : A.W;
  double s = s0, pdiag = 0.0;
  bool diag_pending = false;
  int first = 0;
  if (diag_first && m > 0) {
    const double d = v[0];
    if (diag_out) *diag_out = d;
    if (!skip_diag) { pdiag = __dmul_rn(d, x[c[0]]); diag_pending = true; }
    first = 1;
  }
  for (int base = first; base < m; base += 32) {
    const int k = base + lane;
    double p = 0.0;
    int col = 0x7fffffff;
    if (k < m) { col = c[k]; p = __dmul_rn(v[k], x[col]); }
    s = fold_in_order(s, p, col, min(32, m - base), subtract, diag_pending, pdiag, row);
  }
  if (len > A.W) {
    const int o = ovf_find(A, row);
    const int e = A.ovf_ptr[o + 1];
    for (int base = A.ovf_ptr[o]; base < e; base += 32) {
      const int k = base + lane;
      double p = 0.0;
      int col = 0x7fffffff;
      if (k < e) { col = A.ovf_col[k]; p = __dmul_rn(A.ovf_val[k], x[col]); }
      s = fold_in_order(s, p, col, min(32, e - base), subtract, diag_pending, pdiag, row);
    }
  }
  if (diag_pending) s = subtract ? __dsub_rn(s, pdiag) : __dadd_rn(s, pdiag);
  return s;
}

__global__ void __launch_bounds__(kBlock) k_spmv_exact(HybView A, int diag_first, const double* x, const double* __restrict__ b, double* y,
                                                       const unsigned char* __restrict__ rowflag, int op, int mask_dirichlet, int mask_neumann,
                                                       double* __restrict__ partial) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  double num = 0.0, den = 0.0;
  for (int row = warp; row < A.rows; row += nwarps) {
    const double acc = row_dot_in_order(A, row, x, diag_first != 0, false, false, 0.0, nullptr);
    if (lane == 0) {
      const int flag = rowflag ? rowflag[row] : 0;
      if (op == OP_SPMV) {
        y[row] = acc;
      } else if (op == OP_RESID) {
        const double bi = b[row];
        double t = __dsub_rn(bi, acc);
        if (flag == 1) t = 0.0;
        if (y) y[row] = t;
        num += fabs(t);
        den += fabs(bi);
      } else if (op == OP_PROLONG) {
        if (!(mask_dirichlet && flag == 1)) y[row] = __dadd_rn(y[row], acc);
      } else {
        double t = acc;
        if (flag == 1) t = 0.0;
        if (mask_neumann && flag == 2) t = 0.0;
        y[row] = t;
      }
    }
  }
  if (partial) {
    block_sum2(num, den);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = num; partial[2 * blockIdx.x + 1] = den; }
  }
}

// Regularisation row in the reference's summation order: partial[0] = sum_j val[j]*x[col[j]], ascending columns, ONE running
// sum (grid.cpp:126-136 applied to the dense row of :570-576).  The N-term serial DADD chain cannot be parallelised without
// changing the rounding, so the work around it is: (1) k_regdot_products forms every product in parallel (coalesced, the
// gather of x included) into a scratch vector; (2) one warp streams that vector through a shared-memory ring (128-bit loads,
// four 2 KB blocks in flight) and lane 0 runs the chain at the DADD latency.  First version (load -> gather -> 32 shuffles
// per 32 terms) took 36 ms per 1M terms; this one is bounded by ~8 cycles per term.
__global__ void __launch_bounds__(kBlock) k_regdot_products(const int* __restrict__ col, const double* __restrict__ val, int len, const double* x,
                                                            double* __restrict__ prod) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < len; j += gridDim.x * blockDim.x) prod[j] = __dmul_rn(val[j], x[col[j]]);
}
__global__ void __launch_bounds__(32) k_regdot_exact(const double* __restrict__ prod, int len, double* partial) {
  constexpr int BLK = 256, NBUF = 4;                    // doubles per block, blocks in flight
  __shared__ __align__(16) double buf[NBUF][BLK];
  const int lane = threadIdx.x;
  const int nblk = (len + BLK - 1) / BLK;
  auto issue = [&](int blk) {
    double* dst = buf[blk % NBUF];
    const int base = blk * BLK;
#pragma unroll
    for (int i = 0; i < BLK / 64; i++) {                // 32 lanes x 16 bytes x 4 = 2 KB
      const int e = base + (i * 32 + lane) * 2;
      const unsigned d = (unsigned)__cvta_generic_to_shared(dst + (i * 32 + lane) * 2);
      if (e + 1 < len) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(prod + e) : "memory");
      else { dst[(i * 32 + lane) * 2] = e < len ? prod[e] : 0.0; dst[(i * 32 + lane) * 2 + 1] = 0.0; }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int b = 0; b < NBUF - 1 && b < nblk; b++) issue(b);
  double s = 0.0;
  for (int blk = 0; blk < nblk; blk++) {
    if (blk + NBUF - 1 < nblk) issue(blk + NBUF - 1); else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(NBUF - 1) : "memory");
    __syncwarp();
    if (lane == 0) {
      const double2* q = reinterpret_cast<const double2*>(buf[blk % NBUF]);
      const int cnt = min(BLK, len - blk * BLK);        // absent entries are +0.0: s + 0.0 == s for every s except -0.0, which a sum starting at +0.0 never holds
#pragma unroll 8
      for (int i = 0; i < BLK / 2; i++) {
        if (2 * i < cnt) { const double2 v = q[i]; s = __dadd_rn(s, v.x); s = __dadd_rn(s, v.y); }
      }
    }
    __syncwarp();
  }
  if (lane == 0) partial[0] = s;
}

// s <- s - p_l for lanes l = 0..cnt-1 in lane order.  Lanes >= cnt hold 0.0 (x - 0.0 == x bit for bit), so whole
// groups of 8 are folded without per-element predicates: the shuffles hoist, the DSUB chain is the only dependency.
__device__ __forceinline__ double fold_sub32(double s, double p, int cnt) {
  __syncwarp();   // reconverge: after the lane-dependent polling loops the warp is formally diverged, and every __shfl_sync then takes
                  // the compiler's BRA.DIV slow path (~25 cycles per shuffle instead of ~2; measured in profiles/r01_lex_trace.txt)
#pragma unroll
  for (int g8 = 0; g8 < 4; g8++) {
    if (cnt > g8 * 8) {
      double q[8];
#pragma unroll
      for (int l = 0; l < 8; l++) q[l] = __shfl_sync(0xffffffffu, p, g8 * 8 + l);
#pragma unroll
      for (int l = 0; l < 8; l++) s = __dsub_rn(s, q[l]);
    }
  }
  return s;
}

// Lexicographic sweep, reference-order arithmetic: warp per row; products wait on their dependencies
// in parallel, then the row sum is folded in ascending column order.
template <int T>
__global__ void __launch_bounds__(kBlock) k_sor_lex_exact(HybView A, const unsigned char* __restrict__ rowflag, const double* __restrict__ b,
                                                          const double* __restrict__ x_old, double* x_new, double omega, int* abort_flag,
                                                          long long timeout_cycles) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const long long t_start = clock64();
  for (int row = warp; row < A.rows; row += nwarps) {
    if (rowflag[row] != 0) continue;
    const int len = A.len[row];
    const double* __restrict__ v = row_val(A, row);
    const int* __restrict__ c = row_col(A, row);
    const int m = len < A.W ? len : A.W;
    double prod[T];
    double pv[T];
    int pc[T];
    unsigned pend = 0;
#pragma unroll
    for (int t = 0; t < T; t++) {
      const int k = 1 + lane + t * 32;     // slot 0 is the diagonal
      prod[t] = 0.0; pv[t] = 0.0; pc[t] = 0;
      if (k < m) {
        const double a = v[k];
        const int col = c[k];
        if (col > row) prod[t] = __dmul_rn(a, __ldg(x_old + col));
        else { pv[t] = a; pc[t] = col; pend |= 1u << t; }
      }
    }
    const double wd = omega / v[0];
    const double bi = b[row];
    const double xo_term = __dmul_rn(1 - omega, x_old[row]);
    unsigned spins = 0;
    bool aborted = false;
    while (__any_sync(0xffffffffu, pend != 0)) {
      double xv[T];
#pragma unroll
      for (int t = 0; t < T; t++) xv[t] = (pend & (1u << t)) ? ld_relaxed(x_new + pc[t]) : 0.0;
#pragma unroll
      for (int t = 0; t < T; t++)
        if ((pend & (1u << t)) && !is_sentinel(xv[t])) { prod[t] = __dmul_rn(pv[t], xv[t]); pend &= ~(1u << t); }
      if ((++spins & 0xff) == 0) {
        if (*(volatile int*)abort_flag || clock64() - t_start > timeout_cycles) { atomicExch(abort_flag, 1); aborted = true; break; }
      }
    }
    if (aborted) { if (lane == 0) st_relaxed(x_new + row, 0.0); return; }
    double s = 0.0;                          // x_i = 0; x_i -= a_ij * v_j, ascending j (grid.cpp:122-136)
#pragma unroll
    for (int t = 0; t < T; t++) s = fold_sub32(s, prod[t], m - 1 - t * 32);
    if (len > A.W) {  // spill rows: blocking in-order tail
      const int o = ovf_find(A, row);
      const int e = A.ovf_ptr[o + 1];
      for (int base = A.ovf_ptr[o]; base < e; base += 32) {
        const int k = base + lane;
        double p = 0.0;
        if (k < e) {
          const int col = A.ovf_col[k];
          double xv;
          if (col > row) xv = __ldg(x_old + col);
          else {
            xv = ld_relaxed(x_new + col);
            while (is_sentinel(xv)) {
              if (*(volatile int*)abort_flag || clock64() - t_start > timeout_cycles) { atomicExch(abort_flag, 1); xv = 0.0; break; }
              xv = ld_relaxed(x_new + col);
            }
          }
          p = __dmul_rn(A.ovf_val[k], xv);
        }
        s = fold_sub32(s, p, e - base);
      }
    }
    if (lane == 0) {                         // x+=b; x*=w/d; x+=(1-w)x_old (grid.cpp:137-141)
      double xi = __dadd_rn(s, bi);
      xi = __dmul_rn(xi, wd);
      xi = __dadd_rn(xi, xo_term);
      st_relaxed(x_new + row, xi);
    }
  }
}

// Sweep-pipelined lexicographic SOR for grids without Neumann rows (no global reduction between sweeps).
// All `iters` sweeps run inside ONE cooperative launch on versioned vectors xs[0..iters] (xs[0] = input,
// xs[s] = state after sweep s; swept rows start as the sentinel, skipped rows as their constant value).
// Row i of sweep s reads xs[s][c] for c<i and xs[s-1][c] for c>=i, exactly the values the sequential
// in-place sweeps see, so sweep s+1 trails sweep s by one matrix bandwidth instead of a full sweep.
// Warps are dealt to sweeps round-robin (warp w -> sweep 1 + w % iters) and walk their rows in increasing
// order; ordering tasks by key = row + sweep * D (D > bandwidth) shows every dependency has a smaller key and
// every warp visits its tasks in increasing key order, so the smallest unfinished task can always run:
// no deadlock as long as all CTAs are resident (cooperative launch).  Reference-order arithmetic.
template <int T>
__global__ void __launch_bounds__(kBlock) k_sor_lex_pipe(HybView A, const unsigned char* __restrict__ rowflag, const double* __restrict__ b, double* xs,
                                                         size_t stride, int iters, double omega, int* abort_flag, long long timeout_cycles,
                                                         unsigned sleep_ns) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int sweep = 1 + warp % iters;
  const int q = warp / iters, Q = nwarps / iters;
  if (q >= Q) return;                              // leftover warps when nwarps % iters != 0
  const double* x_old = xs + (size_t)(sweep - 1) * stride;
  double* x_new = xs + (size_t)sweep * stride;
  const long long t_start = clock64();
  for (int row = q; row < A.rows; row += Q) {
    if (rowflag[row] != 0) continue;
    const int len = A.len[row];
    const double* __restrict__ v = row_val(A, row);
    const int* __restrict__ c = row_col(A, row);
    const int m = len < A.W ? len : A.W;
    double prod[T];
    double pv[T];
    const double* pp[T];
    unsigned pend = 0;
#pragma unroll
    for (int t = 0; t < T; t++) {
      const int k = 1 + lane + t * 32;             // slot 0 is the diagonal
      prod[t] = 0.0; pv[t] = 0.0; pp[t] = nullptr;
      if (k < m) {
        pv[t] = v[k];
        const int col = c[k];
        pp[t] = (col > row ? x_old : x_new) + col;
        pend |= 1u << t;
      }
    }
    const double wd = omega / v[0];            // off the critical path: known before any dependency resolves
    const double bi = b[row];
    double xo = 0.0;
    bool need_xo = lane == 0;
    unsigned spins = 0;
    bool aborted = false;
    while (true) {
      double xv[T];
#pragma unroll
      for (int t = 0; t < T; t++) xv[t] = (pend & (1u << t)) ? ld_relaxed(pp[t]) : 0.0;   // all polls in flight together
      if (need_xo) { xo = ld_relaxed(x_old + row); need_xo = is_sentinel(xo); }
#pragma unroll
      for (int t = 0; t < T; t++)
        if ((pend & (1u << t)) && !is_sentinel(xv[t])) { prod[t] = __dmul_rn(pv[t], xv[t]); pend &= ~(1u << t); }
      if (!__any_sync(0xffffffffu, pend != 0 || need_xo)) break;
      if (sleep_ns) __nanosleep(sleep_ns);
      if ((++spins & 0xff) == 0) {
        if (*(volatile int*)abort_flag || clock64() - t_start > timeout_cycles) { atomicExch(abort_flag, 1); aborted = true; break; }
      }
    }
    if (aborted) { if (lane == 0) st_relaxed(x_new + row, 0.0); return; }
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < T; t++) s = fold_sub32(s, prod[t], m - 1 - t * 32);
    if (lane == 0) {
      double xi = __dadd_rn(s, bi);
      xi = __dmul_rn(xi, wd);
      xi = __dadd_rn(xi, __dmul_rn(1 - omega, xo));
      st_relaxed(x_new + row, xi);
    }
  }
}

// Chunked variant of the pipelined sweep (the default).  The critical path of the lexicographic DAG runs mostly along
// CONSECUTIVE rows (with 32-row chunks ~83 % of its edges stay inside a chunk), and a hand-off through
// L2 costs ~1.2 us while one through shared memory costs ~0.65 us.  So a CTA of K warps takes a chunk of K consecutive
// rows (warp w <-> row lo+w): a row waits for the rows of its own chunk on an mbarrier (hardware-suspended, one arrival
// per producer) and reads their values from shared memory; everything else (earlier chunks of this sweep, the previous
// sweep) is polled in L2 as before.  CTAs are dealt to
// sweeps round-robin and walk their chunks in increasing order, so the deadlock-freedom argument of
// k_sor_lex_pipe carries over unchanged.  Reference-order arithmetic.
// optional timing trace of one CTA (diagnostics; MMG_LEX_TRACE=1 and mmg_debug_lex_trace())
__device__ long long g_lex_trace[16 * 64];
__device__ int g_lex_trace_on = 0;
#define LEX_STAMP(slot) do { if (tracing && lane == 0) g_lex_trace[(warp * 8 + (slot)) % (16 * 64)] = clock64(); } while (0)

// Neumann-type grids in the pipelined kernel (Grid::sor grid.cpp:104-146 with its bound_eval_neumann after every sweep): the
// dense regularisation row is the LAST row of a sweep and a column of every interior row, so sweep s+1 cannot start before sweep
// s has finished -- the version vectors express exactly that (every row of sweep s+1 polls xs[s][N]).  One auxiliary CTA per
// sweep (a) forms the regularisation row's dot product in the reference's order WHILE the sweep runs: its warps turn the
// entries into products as the rows they read complete, one lane adds them in ascending column order (the serial chain trails
// the sweep instead of following it); (b) evaluates the Neumann boundary rows of that sweep as their stencils complete.
struct LexNeumann {
  int enabled;
  const int* neu_pts; int n_neu;                  // Neumann boundary nodes, boundary-list order
  const int* reg_col; const double* reg_val; int reg_len, reg_row; double reg_diag;
};
constexpr int kRegRing = 32;                        // blocks of 32 products in flight between the product warps and the adding lane

template <int T, int K>
__device__ void lex_aux_cta(const HybView& A, const unsigned char* __restrict__ rowflag, const double* __restrict__ b, const double* x_old, double* x_new,
                            double omega, int* abort_flag, long long timeout_cycles, const LexNeumann& nm, double (*ring)[32], volatile int* ready,
                            volatile int* consumed) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long t_start = clock64();
  auto poll = [&](const double* p) {
    double v = ld_relaxed(p);
    unsigned spins = 0;
    while (is_sentinel(v)) {
      if ((++spins & 0xff) == 0 && (*(volatile int*)abort_flag || clock64() - t_start > timeout_cycles)) { atomicExch(abort_flag, 1); return 0.0; }
      v = ld_relaxed(p);
    }
    return v;
  };
  constexpr int kProd = K / 2;                       // warps 1 .. kProd-1 form products, warps kProd .. K-1 evaluate boundary rows
  if (threadIdx.x == 0) *consumed = 0;
  if (threadIdx.x < kRegRing) ready[threadIdx.x] = 0;
  __syncthreads();
  const int nblk = (nm.reg_len + 31) / 32;
  if (warp == 0) {                                   // the adding lane: sum_j val_j x_j, ascending j, one running sum
    double s = 0.0;
    for (int blk = 0; blk < nblk; blk++) {
      unsigned spins = 0;
      while (ready[blk % kRegRing] != blk + 1) {
        if ((++spins & 0x3ff) == 0 && (*(volatile int*)abort_flag || clock64() - t_start > timeout_cycles)) { atomicExch(abort_flag, 1); break; }
      }
      __syncwarp();
      if (lane == 0) {
        const double2* q = reinterpret_cast<const double2*>(ring[blk % kRegRing]);
        const int cnt = min(32, nm.reg_len - blk * 32);
#pragma unroll
        for (int i = 0; i < 16; i++) {
          if (2 * i < cnt) { const double2 v = q[i]; s = __dadd_rn(s, v.x); if (2 * i + 1 < cnt) s = __dadd_rn(s, v.y); }
        }
        __threadfence_block();
        *consumed = blk + 1;
      }
      __syncwarp();
    }
    if (lane == 0) {                                 // SOR update of the regularisation row (k_finish_sor_reg / grid.cpp:137-141)
      const double dot = __dadd_rn(0.0, s);
      double xi = -dot;
      xi = __dadd_rn(xi, b[nm.reg_row]);
      xi = __dmul_rn(xi, omega / nm.reg_diag);
      xi = __dadd_rn(xi, __dmul_rn(1 - omega, poll(x_old + nm.reg_row)));
      st_relaxed(x_new + nm.reg_row, xi);
    }
  } else if (warp < kProd) {
    for (int blk = warp - 1; blk < nblk; blk += kProd - 1) {
      const int j = blk * 32 + lane;
      double p = 0.0;
      if (j < nm.reg_len) p = __dmul_rn(nm.reg_val[j], poll(x_new + nm.reg_col[j]));
      unsigned spins = 0;
      while (*consumed < blk - kRegRing + 1) {       // the slot still holds an unread block
        if ((++spins & 0x3ff) == 0 && (*(volatile int*)abort_flag || clock64() - t_start > timeout_cycles)) { atomicExch(abort_flag, 1); break; }
      }
      ring[blk % kRegRing][lane] = p;
      __syncwarp();
      __threadfence_block();
      if (lane == 0) ready[blk % kRegRing] = blk + 1;
    }
  } else {                                           // Grid::bound_eval_neumann for this sweep (k_bound_eval_exact arithmetic, grid.cpp:84-98)
    for (int i = warp - kProd; i < nm.n_neu; i += K - kProd) {
      const int row = nm.neu_pts[i];
      const double* __restrict__ v = row_val(A, row);
      const int* __restrict__ c = row_col(A, row);
      const int m = min(A.len[row], A.W);
      double s = b[row];
      for (int base = 1; base < m; base += 32) {     // slot 0 is the diagonal; the rest in ascending column order
        const int k = base + lane;
        double p = 0.0;
        if (k < m) p = __dmul_rn(v[k], poll(x_new + c[k]));
        s = fold_sub32(s, p, min(32, m - base));
      }
      if (lane == 0) st_relaxed(x_new + row, s / v[0]);
    }
  }
}

template <int T, int K>
__global__ void __launch_bounds__(K * 32) k_sor_lex_chunk(HybView A, const unsigned char* __restrict__ rowflag, const double* __restrict__ b, double* xs,
                                                          size_t stride, int iters, double omega, int* abort_flag, long long timeout_cycles, int Qarg,
                                                          LexNeumann nm) {
  __shared__ double xs_s[K];                 // this chunk's new values
  __shared__ unsigned long long bars[K];     // bars[w]: one arrival per in-chunk dependency of row lo+w
  __shared__ unsigned depmask[K];            // bit j of depmask[w]: row lo+w reads row lo+j (j < w)
  __shared__ __align__(16) double prod_s[K][T * 32 + 2];   // per warp: its row's products in column order, for the in-order fold
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Q = Qarg;
  if ((int)blockIdx.x >= Q * iters) {        // auxiliary CTA of a Neumann-type grid: regularisation row + boundary rows of one sweep
    const int aux = (int)blockIdx.x - Q * iters;
    if (!nm.enabled || aux >= iters) return;
    __shared__ __align__(16) double reg_ring[kRegRing][32];
    __shared__ int reg_ready[kRegRing];
    __shared__ int reg_consumed;
    lex_aux_cta<T, K>(A, rowflag, b, xs + (size_t)aux * stride, xs + (size_t)(aux + 1) * stride, omega, abort_flag, timeout_cycles, nm, reg_ring, reg_ready,
                      &reg_consumed);
    return;
  }
  // Dirichlet-type grids: CTAs are dealt to sweeps (sweep s+1 trails sweep s by one matrix bandwidth).  Neumann-type grids: the
  // sweeps cannot overlap (see LexNeumann), so every CTA works on every sweep in turn.
  const int sweep_first = nm.enabled ? 1 : 1 + (int)blockIdx.x % iters;
  const int sweep_last = nm.enabled ? iters : sweep_first;
  const int q = nm.enabled ? (int)blockIdx.x : (int)blockIdx.x / iters;
  const int Qs = nm.enabled ? Q * iters : Q;      // CTAs sharing one sweep
  volatile double* xs_v = xs_s;
  const unsigned bar_me = (unsigned)__cvta_generic_to_shared(&bars[warp]);
  const unsigned bar_lane = (unsigned)__cvta_generic_to_shared(&bars[lane < K ? lane : 0]);
  const long long t_start = clock64();
  const int nchunks = (A.rows + K - 1) / K;
  // bars[w] collects exactly K arrivals per chunk (one from every warp of the CTA): warps that row lo+w does not read
  // arrive at once, the rows it reads arrive when their value is in xs_s.  Initialised once; the phase parity flips per chunk.
  if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_me), "r"(K) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  unsigned parity = 0;
  for (int sweep = sweep_first; sweep <= sweep_last; sweep++) {
  const double* x_old = xs + (size_t)(sweep - 1) * stride;
  double* x_new = xs + (size_t)sweep * stride;
  for (int chunk = q; chunk < nchunks; chunk += Qs, parity ^= 1u) {
    const int lo = chunk * K;
    const int row = lo + warp;
    const bool live = row < A.rows && rowflag[row] == 0;
    const bool tracing = g_lex_trace_on && blockIdx.x == 2 * iters && chunk == q + 20 * Q && warp < 16;
    __syncthreads();                                   // the previous chunk's slots are no longer in use, its barrier phases are complete
    LEX_STAMP(0);
    double prod[T];
    double pv[T];
    const double* pp[T];
    int ps[T];
    unsigned pend_g = 0, in_chunk = 0, mydeps = 0;
    int m = 0;
    double wd = 0.0, bi = 0.0;
    if (live) {
      const int len = A.len[row];
      const double* __restrict__ v = row_val(A, row);
      const int* __restrict__ c = row_col(A, row);
      m = len < A.W ? len : A.W;
#pragma unroll
      for (int t = 0; t < T; t++) {
        const int k = 1 + lane + t * 32;
        prod[t] = 0.0; pv[t] = 0.0; pp[t] = nullptr; ps[t] = 0;
        if (k < m) {
          pv[t] = v[k];
          const int col = c[k];
          if (col >= lo && col < row && rowflag[col] == 0) { ps[t] = col - lo; in_chunk |= 1u << t; mydeps |= 1u << (col - lo); }
          else {   // a Neumann boundary node keeps, during sweep s, the value bound_eval gave it after sweep s-1
            pp[t] = ((col > row || (nm.enabled && rowflag[col] == 2)) ? x_old : x_new) + col; pend_g |= 1u << t;
          }
        }
      }
      wd = omega / v[0];
      bi = b[row];
    }
    const unsigned deps = __reduce_or_sync(0xffffffffu, mydeps);
    if (lane == 0) depmask[warp] = deps;
    __syncthreads();
    const bool notify = live && lane < K && lane > warp && ((depmask[lane] >> warp) & 1u);   // rows of this chunk that read mine
    if (lane < K && !notify) {                                                                // everyone else gets my arrival now
      unsigned long long state;
      asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(state) : "r"(bar_lane) : "memory");
    }
    if (!live) continue;
    double xo = 0.0;
    bool need_xo = lane == 0;
    unsigned spins = 0;
    bool aborted = false;
    LEX_STAMP(1);
    {
      // Watch phase.  Rows complete roughly in index order, so the LAST dependency to resolve is almost always the largest
      // column below the chunk (this sweep) or the largest column above the row (previous sweep).  Only those two words are
      // polled (lanes 0 and 1) until they turn valid: one sector per round instead of one per pending entry, which keeps
      // the SM's load/shuffle pipe free for the warps that are folding (the per-entry polls saturated it: profiles/r01_lex_trace.txt).
      int cnew = -1, cold = -1;
#pragma unroll
      for (int t = 0; t < T; t++)
        if (pend_g & (1u << t)) {
          const int col = (int)(pp[t] - (pp[t] >= x_new ? x_new : x_old));
          if (pp[t] >= x_new) cnew = max(cnew, col); else cold = max(cold, col);
        }
      cnew = __reduce_max_sync(0xffffffffu, cnew);
      cold = __reduce_max_sync(0xffffffffu, cold);
      const double* watch = lane == 0 ? (cnew >= 0 ? x_new + cnew : nullptr) : lane == 1 ? (cold >= 0 ? x_old + cold : nullptr) : nullptr;
      bool waiting = watch != nullptr;
      while (__any_sync(0xffffffffu, waiting)) {
        if (waiting) waiting = is_sentinel(ld_relaxed(watch));
        if ((++spins & 0xff) == 0) {
          if (*(volatile int*)abort_flag || clock64() - t_start > timeout_cycles) { atomicExch(abort_flag, 1); aborted = true; break; }
        }
      }
    }
    LEX_STAMP(2);
    while (!aborted && __any_sync(0xffffffffu, pend_g != 0 || need_xo)) {      // dependencies outside the chunk: L2 polls
      double xv[T];
#pragma unroll
      for (int t = 0; t < T; t++) xv[t] = (pend_g & (1u << t)) ? ld_relaxed(pp[t]) : 0.0;
      if (need_xo) { xo = ld_relaxed(x_old + row); need_xo = is_sentinel(xo); }
#pragma unroll
      for (int t = 0; t < T; t++)
        if ((pend_g & (1u << t)) && !is_sentinel(xv[t])) { prod[t] = __dmul_rn(pv[t], xv[t]); pend_g &= ~(1u << t); }
      if ((++spins & 0xff) == 0) {
        if (*(volatile int*)abort_flag || clock64() - t_start > timeout_cycles) { atomicExch(abort_flag, 1); aborted = true; break; }
      }
    }
    LEX_STAMP(3);
    if (deps && !aborted) {                                        // dependencies inside the chunk: hardware-suspended wait
      unsigned done = 0;
      spins = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"   // %3: suspend-time hint, so the wait really sleeps
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar_me), "r"(parity), "r"(0x989680u) : "memory");
        if (!done && (++spins & 0x3f) == 0) {
          if (*(volatile int*)abort_flag || clock64() - t_start > timeout_cycles) { atomicExch(abort_flag, 1); aborted = true; break; }
        }
      }
#pragma unroll
      for (int t = 0; t < T; t++)
        if (in_chunk & (1u << t)) prod[t] = __dmul_rn(pv[t], xs_v[ps[t]]);
    }
    LEX_STAMP(4);
    double xi = 0.0;
    if (!aborted) {
      // in-order fold fed from shared memory: every lane writes its products, then the whole warp reads them back with
      // broadcast 128-bit loads (18 loads for 36 entries instead of 72 shuffles) and runs the serial DSUB chain
#pragma unroll
      for (int t = 0; t < T; t++) prod_s[warp][lane + t * 32] = prod[t];
      __syncwarp();
      double s = 0.0;
      const int cnt = m - 1;
      const double2* p2 = reinterpret_cast<const double2*>(prod_s[warp]);
#pragma unroll
      for (int t = 0; t < T; t++) {                 // 32 entries at a time: all 16 loads first, then the chain
        if (cnt > t * 32) {
          double2 q[16];
#pragma unroll
          for (int i = 0; i < 16; i++) q[i] = p2[t * 16 + i];
#pragma unroll
          for (int g4 = 0; g4 < 4; g4++) {          // groups of 8 entries; absent entries are 0.0 and s - 0.0 == s
            if (cnt > t * 32 + g4 * 8) {
#pragma unroll
              for (int i = 0; i < 4; i++) { s = __dsub_rn(s, q[g4 * 4 + i].x); s = __dsub_rn(s, q[g4 * 4 + i].y); }
            }
          }
        }
      }
      __syncwarp();
      if (A.n_ovf && A.len[row] > A.W) {          // implicit-Neumann fill-in: the tail holds the row's largest columns, folded last
        const int o = ovf_find(A, row);
        const int e = A.ovf_ptr[o + 1];
        for (int base = A.ovf_ptr[o]; base < e; base += 32) {
          const int k = base + lane;
          double p = 0.0;
          if (k < e) {
            const int col = A.ovf_col[k];
            const double* src = ((col > row || rowflag[col] == 2) ? x_old : x_new) + col;
            double xv = ld_relaxed(src);
            unsigned sp = 0;
            while (is_sentinel(xv)) {
              if ((++sp & 0xff) == 0 && (*(volatile int*)abort_flag || clock64() - t_start > timeout_cycles)) { atomicExch(abort_flag, 1); xv = 0.0; break; }
              xv = ld_relaxed(src);
            }
            p = __dmul_rn(A.ovf_val[k], xv);
          }
          s = fold_sub32(s, p, e - base);
        }
      }
      xi = __dadd_rn(s, bi);
      xi = __dmul_rn(xi, wd);
      xi = __dadd_rn(xi, __dmul_rn(1 - omega, __shfl_sync(0xffffffffu, xo, 0)));
    }
    LEX_STAMP(5);
    if (lane == 0) { xs_v[warp] = xi; st_relaxed(x_new + row, xi); }   // (on abort: a finite value, so nobody behind us hangs)
    __syncwarp();
    if (notify) {
      unsigned long long state;
      asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(state) : "r"(bar_lane) : "memory");
    }
    LEX_STAMP(6);
  }
  }
}

// xs[0] = x; xs[s>=1][i] = skipped(i) ? x[i] : sentinel
__global__ void __launch_bounds__(kBlock) k_pipe_init(const unsigned char* __restrict__ rowflag, const double* __restrict__ x, double* xs, size_t stride,
                                                      int iters, int total, int neumann_rows_are_computed = 0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const double xi = x[i];
  // Dirichlet rows are constants of every version; Neumann boundary rows are re-evaluated after every sweep (grid.cpp:144) when the
  // pipelined lexicographic kernel runs a Neumann-type grid, so their later versions start as "not yet" too
  const bool skipped = rowflag[i] == 1 || (rowflag[i] == 2 && !neumann_rows_are_computed);
  xs[i] = xi;
  for (int s = 1; s <= iters; s++) xs[(size_t)s * stride + i] = skipped ? xi : __longlong_as_double((long long)kSentinelBits);
}

// ------------------------------------------------------------------------------------------------
// Block-lexicographic SOR (throughput mode that keeps the lexicographic character).
// Rows are cut into contiguous blocks of B rows; same-colour blocks do not touch each other, so one
// launch sweeps every block of a colour concurrently.  Inside a block the sweep is the reference's
// ascending-row Gauss-Seidel: row i waits (sentinel in x_work) only for rows j<i of ITS OWN block and
// reads every other column from x, which no block of this colour modifies.  The DAG depth is that of
// one block (independent of N) instead of the whole grid.  Positions are dealt block-interleaved
// (p -> block p % nblk, offset p / nblk) so all blocks advance together; dependencies have smaller p.
// Reference-order arithmetic: bit-identical to the oracle's sor_blocklex.
// ------------------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(kBlock) k_sor_blk(HybView A, const unsigned char* __restrict__ rowflag, const double* __restrict__ b,
                                                    const double* x, double* x_work, const int* __restrict__ phase_blocks, int nblk, int B,
                                                    double omega, int* abort_flag, long long timeout_cycles) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const long long t_start = clock64();
  const long long total = (long long)nblk * B;
  for (long long p = warp; p < total; p += nwarps) {
    const int blk = phase_blocks[(int)(p % nblk)];
    const int lo = blk * B;
    const int row = lo + (int)(p / nblk);
    if (row >= A.rows || rowflag[row] != 0) continue;
    const int len = A.len[row];
    const double* __restrict__ v = row_val(A, row);
    const int* __restrict__ c = row_col(A, row);
    const int m = len < A.W ? len : A.W;
    double prod[T];
    double pv[T];
    int pc[T];
    unsigned pend = 0;
#pragma unroll
    for (int t = 0; t < T; t++) {
      const int k = 1 + lane + t * 32;
      prod[t] = 0.0; pv[t] = 0.0; pc[t] = 0;
      if (k < m) {
        const double a = v[k];
        const int col = c[k];
        if (col < row && col >= lo) { pv[t] = a; pc[t] = col; pend |= 1u << t; }
        else prod[t] = __dmul_rn(a, x[col]);
      }
    }
    const double wd = omega / v[0];
    const double bi = b[row];
    const double xo_term = __dmul_rn(1 - omega, x[row]);
    unsigned spins = 0;
    bool aborted = false;
    while (__any_sync(0xffffffffu, pend != 0)) {
      double xv[T];
#pragma unroll
      for (int t = 0; t < T; t++) xv[t] = (pend & (1u << t)) ? ld_relaxed(x_work + pc[t]) : 0.0;
#pragma unroll
      for (int t = 0; t < T; t++)
        if ((pend & (1u << t)) && !is_sentinel(xv[t])) { prod[t] = __dmul_rn(pv[t], xv[t]); pend &= ~(1u << t); }
      if ((++spins & 0xff) == 0) {
        if (*(volatile int*)abort_flag || clock64() - t_start > timeout_cycles) { atomicExch(abort_flag, 1); aborted = true; break; }
      }
    }
    if (aborted) { if (lane == 0) st_relaxed(x_work + row, 0.0); return; }
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < T; t++) s = fold_sub32(s, prod[t], m - 1 - t * 32);
    if (lane == 0) {
      double xi = __dadd_rn(s, bi);
      xi = __dmul_rn(xi, wd);
      xi = __dadd_rn(xi, xo_term);
      st_relaxed(x_work + row, xi);
    }
  }
}

// before a colour phase: x_work <- sentinel on the rows the phase will write (old value on rows it skips)
__global__ void __launch_bounds__(kBlock) k_blk_init(const unsigned char* __restrict__ rowflag, const double* __restrict__ x, double* x_work,
                                                     const int* __restrict__ phase_blocks, int nblk, int B, int rows) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long long)nblk * B) return;
  const int row = phase_blocks[(int)(p / B)] * B + (int)(p % B);
  if (row < rows) x_work[row] = rowflag[row] != 0 ? x[row] : __longlong_as_double((long long)kSentinelBits);
}
// after a colour phase: publish the block results
__global__ void __launch_bounds__(kBlock) k_blk_merge(const unsigned char* __restrict__ rowflag, double* x, const double* __restrict__ x_work,
                                                      const int* __restrict__ phase_blocks, int nblk, int B, int rows) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long long)nblk * B) return;
  const int row = phase_blocks[(int)(p / B)] * B + (int)(p % B);
  if (row < rows && rowflag[row] == 0) x[row] = x_work[row];
}
// block adjacency bitmap: bit (a, b) set when a swept row of block a stores a swept column of block b != a
__global__ void __launch_bounds__(kBlock) k_blk_adjacency(HybView A, const unsigned char* __restrict__ rowflag, int B, int nb, unsigned* bitmap) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int words = (nb + 31) / 32;
  for (int row = warp; row < A.rows; row += nwarps) {
    if (rowflag[row] != 0) continue;
    const int len = A.len[row];
    const int* __restrict__ c = row_col(A, row);
    const int m = len < A.W ? len : A.W;
    const int a = row / B;
    for (int k = lane; k < m; k += 32) {
      const int col = c[k];
      if (col < A.rows && rowflag[col] == 0) {
        const int bb = col / B;
        if (bb != a) atomicOr(bitmap + (size_t)a * words + (bb >> 5), 1u << (bb & 31));
      }
    }
    if (len > A.W) {
      const int o = ovf_find(A, row);
      for (int k = A.ovf_ptr[o] + lane; k < A.ovf_ptr[o + 1]; k += 32) {
        const int col = A.ovf_col[k];
        if (col < A.rows && rowflag[col] == 0) {
          const int bb = col / B;
          if (bb != a) atomicOr(bitmap + (size_t)a * words + (bb >> 5), 1u << (bb & 31));
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kBlock) k_sor_mc_exact(HybView A, const int* __restrict__ rows_list, int count, const double* __restrict__ b, double* x,
                                                         double omega) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < count; i += nwarps) {
    const int row = rows_list[i];
    double diag = 0.0;
    const double s = row_dot_in_order(A, row, x, true, true, true, 0.0, &diag);
    if (lane == 0) {
      double xi = __dadd_rn(s, b[row]);
      xi = __dmul_rn(xi, omega / diag);
      xi = __dadd_rn(xi, __dmul_rn(1 - omega, x[row]));
      x[row] = xi;
    }
  }
}

// t = b_c; t -= v_k * a_ck ascending k != c; t /= a_cc  (grid.cpp:84-98)
__global__ void __launch_bounds__(kBlock) k_bound_eval_exact(HybView A, const int* __restrict__ nodes, int count, const double* __restrict__ b, double* x) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < count; i += nwarps) {
    const int row = nodes[i];
    double diag = 0.0;
    const double t = row_dot_in_order(A, row, x, true, true, true, b[row], &diag);
    if (lane == 0) x[row] = t / diag;
  }
}

// ================================================================================================
// Throughput ("fast") kernels, second generation.  The first-generation loops above were latency bound
// (ncu: long-scoreboard stalls, ~25-30 % of HBM peak, profiles/r01_*): one row per lane group and a
// load -> gather -> FMA chain per iteration.  Here the per-lane trip count is a compile-time constant,
// two rows are in flight per lane group, all matrix loads are issued before any gather and all gathers
// before any arithmetic, the matrix stream is marked evict-first in L2 and the gathered vector
// evict-last so the 8*N-byte x stays L2 resident under the 12*nnz-byte stream.
// ================================================================================================
// pull the chunk of a row that this lane group will process on its NEXT trip into L2 (one 128-byte line per lane)
template <int LPR>
__device__ __forceinline__ void prefetch_chunk_l2(const HybView& A, int row, int gl) {
  const unsigned char* p = A.chunks + (size_t)row * A.chunk_bytes + (size_t)gl * 128;
  if ((size_t)gl * 128 < A.chunk_bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

template <int LPR, int ITER>
struct RowRegs {
  double v[ITER];
  int c[ITER];
};
template <int LPR, int ITER>
__device__ __forceinline__ void row_fetch(const HybView& A, int row, bool valid, int gl, unsigned long long pol, RowRegs<LPR, ITER>& r, int& len) {
  len = valid ? A.len[row] : 0;
  const double* __restrict__ v = row_val(A, valid ? row : 0);
  const int* __restrict__ c = row_col(A, valid ? row : 0);
  const int m = len < A.W ? len : A.W;
#pragma unroll
  for (int t = 0; t < ITER; t++) {
    const int k = gl + t * LPR;
    const bool ok = k < m;
    r.v[t] = ok ? ldg_stream_f64(v + k, pol) : 0.0;
    r.c[t] = ok ? ldg_stream_s32(c + k, pol) : -1;
  }
}
// acc = (+/-) sum_k v_k * x[c_k]; when skip_first the slot-0 entry (the diagonal) is returned in diag instead
template <int LPR, int ITER, bool SUB, bool DIAG0>
__device__ __forceinline__ double row_accumulate(const RowRegs<LPR, ITER>& r, const double* x, unsigned long long pol, int gl, double& diag) {
  double xx[ITER];
#pragma unroll
  for (int t = 0; t < ITER; t++) xx[t] = r.c[t] >= 0 ? ldg_keep(x + r.c[t], pol) : 0.0;
  double acc = 0.0;
#pragma unroll
  for (int t = 0; t < ITER; t++) {
    if (DIAG0 && t == 0 && gl == 0) { diag = r.v[0]; continue; }
    const double p = __dmul_rn(r.v[t], xx[t]);
    acc = SUB ? __dsub_rn(acc, p) : __dadd_rn(acc, p);
  }
  return acc;
}
template <int LPR, bool SUB>
__device__ __forceinline__ double row_overflow(const HybView& A, int row, int len, const double* x, int gl, double acc) {
  if (len > A.W) {
    const int o = ovf_find(A, row);
    for (int k = A.ovf_ptr[o] + gl; k < A.ovf_ptr[o + 1]; k += LPR) {
      const double p = __dmul_rn(A.ovf_val[k], x[A.ovf_col[k]]);
      acc = SUB ? __dsub_rn(acc, p) : __dadd_rn(acc, p);
    }
  }
  return acc;
}

template <int LPR, int ITER, int ROWS>
__global__ void __launch_bounds__(kBlock) k_spmv2(HybView A, const double* x, const double* __restrict__ b, double* y,
                                                  const unsigned char* __restrict__ rowflag, int op, int mask_dirichlet, int mask_neumann,
                                                  double* __restrict__ partial, int xlen, int row0, int nrows) {
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR;
  const unsigned gmask = group_mask<LPR>(lane);
  constexpr int GPW = 32 / LPR;
  constexpr int TR = (kBlock / 32) * GPW * ROWS;       // rows per CTA tile
  const unsigned long long keep = policy_evict_last(), stream = policy_evict_first();
  double num = 0.0, den = 0.0;
  const int ntiles = (nrows + TR - 1) / TR;            // rows [row0, row0+nrows): the whole matrix, or this rank's block
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    RowRegs<LPR, ITER> r[ROWS];
    int row[ROWS], len[ROWS];
    bool valid[ROWS];
    double acc[ROWS];
#pragma unroll
    for (int h = 0; h < ROWS; h++) {
      row[h] = row0 + tile * TR + (threadIdx.x >> 5) * GPW * ROWS + h * GPW + lane / LPR;
      valid[h] = row[h] < row0 + nrows;
      row_fetch<LPR, ITER>(A, row[h], valid[h], gl, stream, r[h], len[h]);
    }
    double dummy;
#pragma unroll
    for (int h = 0; h < ROWS; h++) acc[h] = row_accumulate<LPR, ITER, false, false>(r[h], x, keep, gl, dummy);
    if (A.n_ovf) {
#pragma unroll
      for (int h = 0; h < ROWS; h++) if (valid[h]) acc[h] = row_overflow<LPR, false>(A, row[h], len[h], x, gl, acc[h]);
    }
#pragma unroll
    for (int h = 0; h < ROWS; h++) acc[h] = group_sum<LPR>(acc[h], gmask);
#pragma unroll
    for (int h = 0; h < ROWS; h++) {
      if (valid[h] && gl == 0) {
        const int flag = rowflag ? rowflag[row[h]] : 0;
        if (op == OP_SPMV) {
          y[row[h]] = acc[h];
        } else if (op == OP_RESID) {
          const double bi = b[row[h]];
          double t = __dsub_rn(bi, acc[h]);
          if (flag == 1) t = 0.0;
          if (y) y[row[h]] = t;
          num += fabs(t);
          den += fabs(bi);
        } else if (op == OP_PROLONG) {
          if (!(mask_dirichlet && flag == 1)) y[row[h]] = __dadd_rn(y[row[h]], acc[h]);
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
    block_sum2(num, den);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = num; partial[2 * blockIdx.x + 1] = den; }
  }
}

template <int LPR, int ITER, int ROWS>
__device__ __forceinline__ void mc_phase(const HybView& A, const int* __restrict__ rows_list, int count, const double* __restrict__ b, double* x,
                                         double omega) {
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR;
  const unsigned gmask = group_mask<LPR>(lane);
  constexpr int GPW = 32 / LPR;
  constexpr int TR = (kBlock / 32) * GPW * ROWS;
  const unsigned long long keep = policy_evict_last(), stream = policy_evict_first();
  const int ntiles = (count + TR - 1) / TR;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    RowRegs<LPR, ITER> r[ROWS];
    int row[ROWS], len[ROWS];
    bool valid[ROWS];
    double acc[ROWS], diag[ROWS];
#pragma unroll
    for (int h = 0; h < ROWS; h++) {
      const int i = tile * TR + (threadIdx.x >> 5) * GPW * ROWS + h * GPW + lane / LPR;
      valid[h] = i < count;
      row[h] = valid[h] ? rows_list[i] : 0;
    }
#pragma unroll
    for (int h = 0; h < ROWS; h++) row_fetch<LPR, ITER>(A, row[h], valid[h], gl, stream, r[h], len[h]);
#pragma unroll
    for (int h = 0; h < ROWS; h++) { diag[h] = 0.0; acc[h] = row_accumulate<LPR, ITER, true, true>(r[h], x, keep, gl, diag[h]); }
    if (A.n_ovf) {
#pragma unroll
      for (int h = 0; h < ROWS; h++) if (valid[h]) acc[h] = row_overflow<LPR, true>(A, row[h], len[h], x, gl, acc[h]);
    }
#pragma unroll
    for (int h = 0; h < ROWS; h++) acc[h] = group_sum<LPR>(acc[h], gmask);
    if (gl == 0) {
#pragma unroll
      for (int h = 0; h < ROWS; h++) {
        if (valid[h]) {
          double xi = __dadd_rn(acc[h], b[row[h]]);
          xi = __dmul_rn(xi, omega / diag[h]);
          xi = __dadd_rn(xi, __dmul_rn(1 - omega, x[row[h]]));
          x[row[h]] = xi;
        }
      }
    }
  }
}

template <int LPR, int ITER, int ROWS>
__global__ void __launch_bounds__(kBlock) k_sor_mc2(HybView A, const int* __restrict__ rows_list, int count, const double* __restrict__ b, double* x,
                                                    double omega, int xlen) {
  mc_phase<LPR, ITER, ROWS>(A, rows_list, count, b, x, omega);
}

// ------------------------------------------------------------------------------------------------
// Colour-major packed copy of the operator (fourth generation of the multicolour sweep).
// ncu without cache flushing (profiles/r01_steady_state_4M.txt) showed the per-colour phases of the natural-order sweep (one cooperative launch, rows of a colour through an index list; removed in round 2) at
// 46 % of DRAM peak but 72 % of the L2 sector-read cap: a colour's rows are every ~15th row of the matrix, so
// (a) the 448-byte row chunks are isolated DRAM bursts, (b) rows_list -> len -> chunk -> gather is a four-deep
// chain of dependent loads, and (c) a CTA's 32 rows lie on a thin arc of one BFS ring, so hardly any gathered sector
// of x is shared between its warps (37 L2 sectors per row).  The packed copy stores the chunks of every colour
// contiguously, in Morton order of the node coordinates inside a colour: a phase streams one contiguous range, a
// CTA tile is a compact 2-D patch whose neighbourhoods overlap in L1, the diagonal slot carries the row index
// (slot 0's column IS the row, diag_first), and padded slots are (0.0, own row), so neither the row list nor the
// length array is read.  Rows of one colour are independent, so their order does not change any result.
// ------------------------------------------------------------------------------------------------
__global__ void k_pack_chunks(const unsigned char* __restrict__ src, size_t chunk_bytes, const int* __restrict__ rows, int count,
                              unsigned char* __restrict__ dst) {
  const int w = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= count) return;
  const uint4* s = reinterpret_cast<const uint4*>(src + (size_t)rows[w] * chunk_bytes);
  uint4* d = reinterpret_cast<uint4*>(dst + (size_t)w * chunk_bytes);
  for (int k = lane; k < (int)(chunk_bytes / 16); k += 32) d[k] = s[k];
}

template <int LPR, int ITER, int ROWS>
__device__ __forceinline__ void mc_phase_packed(const unsigned char* __restrict__ chunks, size_t chunk_bytes, int W, int count,
                                                const double* __restrict__ b, double* x, double omega) {
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR;
  const unsigned gmask = group_mask<LPR>(lane);
  constexpr int GPW = 32 / LPR;
  constexpr int TR = (kBlock / 32) * GPW * ROWS;
  const unsigned long long keep = policy_evict_last(), stream = policy_evict_first();
  const int ntiles = (count + TR - 1) / TR;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    double v[ROWS][ITER], xx[ROWS][ITER], bi[ROWS], acc[ROWS];
    int c[ROWS][ITER];
    bool valid[ROWS];
#pragma unroll
    for (int h = 0; h < ROWS; h++) {
      const int i = tile * TR + (threadIdx.x >> 5) * GPW * ROWS + h * GPW + lane / LPR;
      valid[h] = i < count;
      const double* pv = reinterpret_cast<const double*>(chunks + (size_t)(valid[h] ? i : 0) * chunk_bytes);
      const int* pc = reinterpret_cast<const int*>(pv + W);
#pragma unroll
      for (int t = 0; t < ITER; t++) {
        const int k = gl + t * LPR;
        const bool ok = valid[h] && k < W;
        v[h][t] = ok ? ldg_stream_f64(pv + k, stream) : 0.0;
        c[h][t] = ok ? ldg_stream_s32(pc + k, stream) : -1;
      }
    }
#pragma unroll
    for (int h = 0; h < ROWS; h++)
#pragma unroll
      for (int t = 0; t < ITER; t++) {
        if (c[h][t] != -1) c[h][t] &= 0x3fffffff;          // bit 31: neighbour of a lower colour (k_sor_mc_flow); bit 30: overflow tail / previous colour (mmg_stream.cu)
        xx[h][t] = c[h][t] >= 0 ? ldg_keep(x + c[h][t], keep) : 0.0;
      }
#pragma unroll
    for (int h = 0; h < ROWS; h++) bi[h] = (gl == 0 && valid[h]) ? b[c[h][0]] : 0.0;   // slot 0 is the diagonal: its column is the row
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
        if (valid[h]) {
          double xi = __dadd_rn(acc[h], bi[h]);
          xi = __dmul_rn(xi, omega / v[h][0]);
          xi = __dadd_rn(xi, __dmul_rn(1 - omega, xx[h][0]));     // xx[h][0] on lane 0 is x[row] before the update
          x[c[h][0]] = xi;
        }
      }
    }
  }
}

template <int LPR, int ITER, int ROWS>
__global__ void __launch_bounds__(kBlock) k_sor_mc_packed(const unsigned char* __restrict__ chunks, size_t chunk_bytes, int W,
                                                          const int* __restrict__ colour_ptr, int ncolours, int iters, const double* __restrict__ b,
                                                          double* x, double omega) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  for (int it = 0; it < iters; it++)
    for (int c = 0; c < ncolours; c++) {
      const int first = colour_ptr[c], count = colour_ptr[c + 1] - first;
      mc_phase_packed<LPR, ITER, ROWS>(chunks + (size_t)first * chunk_bytes, chunk_bytes, W, count, b, x, omega);
      grid.sync();
    }
}

// ------------------------------------------------------------------------------------------------
// Barrier-free multicolour sweep (sixth generation): the values are the ready flags.
// The packed kernel loses ~4 us per colour phase to grid.sync() + drain + refill (22 phases per sweep at n=37) and
// another ~12 % to quantisation (7.04 tiles per CTA means some CTAs run 8).  Here every sweep writes its own
// version of the vector, xs[s+1], whose swept rows start as the sentinel NaN of the lexicographic kernels: a row of
// colour c in sweep s reads neighbours of a lower colour from xs[s+1] (bit 31 of the packed column, set once when
// the copy is built) and everything else from xs[s], polling with ld.relaxed.gpu while it sees the sentinel.  No
// barrier, no fence: CTAs run from one colour into the next, a row only waits for the handful of rows it really
// depends on, and those were scheduled a whole phase earlier.  Tiles never span two colours, tiles are dealt
// round-robin in (sweep, colour, tile) order and every CTA walks its tiles in that order, so the lowest unfinished
// tile always has a resident owner whose operands are complete: no deadlock under a cooperative launch; clock64
// watchdog as in the lexicographic kernels.  Arithmetic per row identical to k_sor_mc_packed.
// ------------------------------------------------------------------------------------------------
__global__ void k_mark_lower_colour(unsigned char* chunks, size_t chunk_bytes, int W, int total, const int* __restrict__ colour, int ncolours) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int* pc = reinterpret_cast<int*>(chunks + (size_t)i * chunk_bytes + (size_t)W * 8);
  const int own = colour[pc[0]];                              // slot 0 is the diagonal
  for (int k = 1; k < W; k++) {
    const int col = pc[k];
    const int cn = colour[col];
    unsigned m = (unsigned)col;
    if (cn >= 0 && cn < own) m |= 0x80000000u;                // bit 31: k_sor_mc_flow reads the newer version
    pc[k] = (int)m;
  }
}

// rows longer than the chunk width (implicit-Neumann fill-in) carry bit 30 in their diagonal column: the TMA-fed sweep then
// fetches the tail from the overflow CSR
__global__ void k_mark_overflow(unsigned char* chunks, size_t chunk_bytes, int W, int total, const int* __restrict__ len) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int* pc = reinterpret_cast<int*>(chunks + (size_t)i * chunk_bytes + (size_t)W * 8);
  if (len[pc[0] & 0x3fffffff] > W) pc[0] |= 0x40000000;
}

template <int LPR, int ITER, int ROWS, bool PEER>
__global__ void __launch_bounds__(kBlock) k_sor_mc_flow(const unsigned char* __restrict__ chunks, size_t chunk_bytes, int W,
                                                        const int* __restrict__ colour_ptr, int ncolours, int iters, const double* __restrict__ b,
                                                        double* xs, size_t stride, double omega, int* abort_flag, long long timeout_cycles,
                                                        PeerSends peers) {
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR;
  const unsigned gmask = group_mask<LPR>(lane);
  constexpr int GPW = 32 / LPR;
  constexpr int TR = (kBlock / 32) * GPW * ROWS;
  const unsigned long long stream = policy_evict_first();
  const long long t_start = clock64();
  for (int it = 0; it < iters; it++) {
    const double* xold = xs + (size_t)it * stride;
    double* xnew = xs + (size_t)(it + 1) * stride;
    for (int c = 0; c < ncolours; c++) {
      const int first = colour_ptr[c], count = colour_ptr[c + 1] - first;
      const unsigned char* base = chunks + (size_t)first * chunk_bytes;
      const int ntiles = (count + TR - 1) / TR;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        double v[ROWS][ITER], xx[ROWS][ITER], bi[ROWS], acc[ROWS];
        int cc[ROWS][ITER];
        unsigned pend = 0, newer = 0;
#pragma unroll
        for (int h = 0; h < ROWS; h++) {
          const int i = tile * TR + (threadIdx.x >> 5) * GPW * ROWS + h * GPW + lane / LPR;
          const bool valid = i < count;
          const double* pv = reinterpret_cast<const double*>(base + (size_t)(valid ? i : 0) * chunk_bytes);
          const int* pc = reinterpret_cast<const int*>(pv + W);
#pragma unroll
          for (int t = 0; t < ITER; t++) {
            const int k = gl + t * LPR;
            const bool ok = valid && k < W;
            v[h][t] = ok ? ldg_stream_f64(pv + k, stream) : 0.0;
            cc[h][t] = ok ? ldg_stream_s32(pc + k, stream) : -1;
          }
        }
#pragma unroll
        for (int h = 0; h < ROWS; h++)
#pragma unroll
          for (int t = 0; t < ITER; t++) {
            const int raw = cc[h][t];
            if (raw != -1) {
              cc[h][t] = raw & 0x3fffffff;
              if (raw < 0) newer |= 1u << (h * ITER + t);
              xx[h][t] = PEER ? ld_relaxed_sys((raw < 0 ? xnew : xold) + cc[h][t]) : ld_relaxed((raw < 0 ? xnew : xold) + cc[h][t]);
              if (is_sentinel(xx[h][t])) pend |= 1u << (h * ITER + t);
            } else xx[h][t] = 0.0;
          }
#pragma unroll
        for (int h = 0; h < ROWS; h++) bi[h] = (gl == 0 && cc[h][0] != -1) ? b[cc[h][0]] : 0.0;
        bool aborted = false;
        while (__any_sync(0xffffffffu, pend != 0)) {            // rare: an operand of this tile is still being computed
#pragma unroll
          for (int h = 0; h < ROWS; h++)
#pragma unroll
            for (int t = 0; t < ITER; t++)
              if (pend & (1u << (h * ITER + t))) {
                const double* src = ((newer >> (h * ITER + t)) & 1u ? xnew : xold) + cc[h][t];
                xx[h][t] = PEER ? ld_relaxed_sys(src) : ld_relaxed(src);     // a neighbour GPU stores these with st.relaxed.sys: same scope on both sides
                if (!is_sentinel(xx[h][t])) pend &= ~(1u << (h * ITER + t));
              }
          if (*(volatile int*)abort_flag || clock64() - t_start > timeout_cycles) { atomicExch(abort_flag, 1); aborted = true; break; }
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
            if (cc[h][0] != -1) {
              double xi = __dadd_rn(acc[h], bi[h]);
              xi = __dmul_rn(xi, omega / v[h][0]);
              xi = __dadd_rn(xi, __dmul_rn(1 - omega, xx[h][0]));
              const int row = cc[h][0];
              if (aborted) xi = 0.0;                                // on abort: unblock everyone behind us
              st_relaxed(xnew + row, xi);
              // multi-GPU: a row next to a cut also lands in the neighbour rank's copy of this version, over NVLink
              // (constant indices: a dynamically indexed kernel parameter would be copied to local memory)
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
  }
}

// ------------------------------------------------------------------------------------------------
// Small levels.  Below ~10^4 rows a colour phase is a handful of rows per SM and the cooperative kernel is pure latency:
// ~1.7 us per phase (grid barrier + an L2 round trip each for chunk, gather and store), 160-220 phases per call, 0.4 ms
// per level and cycle -- the four coarsest levels cost as much as the whole 1M-row level.  Here ONE CTA of 1024 threads
// keeps values_ and source_ in shared memory for the whole call, reads the (L2-resident) packed operator with the
// chunk of its next tile prefetched into registers across the block barrier, and separates phases with __syncthreads.
// Same per-row arithmetic as k_sor_mc_packed, so the same bits.
// ------------------------------------------------------------------------------------------------
constexpr int kSmallThreads = 1024;
constexpr int kSmallMaxRows = 512;        // measured per cycle: 256 rows 0.17 ms (barrier-free kernel 0.16), 1024 rows 0.25 (0.17), 3969 rows 0.56 (0.18)

template <int LPR, int ITER>
__global__ void __launch_bounds__(kSmallThreads, 1) k_sor_mc_small(const unsigned char* __restrict__ chunks, size_t chunk_bytes, int W,
                                                                   const int* __restrict__ colour_ptr, int ncolours, int iters,
                                                                   const double* __restrict__ b, double* __restrict__ x, double omega, int xlen) {
  extern __shared__ __align__(16) double small_smem[];
  double* xs = small_smem;
  double* bs = small_smem + xlen;
  for (int i = threadIdx.x; i < xlen; i += kSmallThreads) { xs[i] = x[i]; bs[i] = b[i]; }
  const int lane = threadIdx.x & 31, gl = lane % LPR;
  const unsigned gmask = group_mask<LPR>(lane);
  constexpr int GPW = 32 / LPR;
  constexpr int TR = (kSmallThreads / 32) * GPW;                     // rows per pass of the CTA
  const int slot = (threadIdx.x >> 5) * GPW + lane / LPR;
  const unsigned long long stream = policy_evict_last();            // the operator of a small level lives in L2
  double v[ITER], nv[ITER];
  int c[ITER], nc[ITER];
  auto fetch = [&](int colour, int tile) {
    const int first = colour_ptr[colour], count = colour_ptr[colour + 1] - first;
    const int i = tile * TR + slot;
    const bool valid = i < count;
    const double* pv = reinterpret_cast<const double*>(chunks + (size_t)(first + (valid ? i : 0)) * chunk_bytes);
    const int* pc = reinterpret_cast<const int*>(pv + W);
#pragma unroll
    for (int t = 0; t < ITER; t++) {
      const int k = gl + t * LPR;
      const bool ok = valid && k < W;
      nv[t] = ok ? ldg_stream_f64(pv + k, stream) : 0.0;
      nc[t] = ok ? (ldg_stream_s32(pc + k, stream) & 0x3fffffff) : -1;
    }
  };
  const int nphases = iters * ncolours;
  bool pending = false;
  __syncthreads();
  for (int p = 0; p < nphases; p++) {
    const int col = p % ncolours, ncol = col + 1 == ncolours ? 0 : col + 1;
    const int ntiles = (colour_ptr[col + 1] - colour_ptr[col] + TR - 1) / TR;
    const int nntiles = p + 1 < nphases ? (colour_ptr[ncol + 1] - colour_ptr[ncol] + TR - 1) / TR : 0;
    for (int tile = 0; tile < ntiles; tile++) {
      if (!pending) fetch(col, tile);
#pragma unroll
      for (int t = 0; t < ITER; t++) { v[t] = nv[t]; c[t] = nc[t]; }
      pending = false;
      if (tile + 1 < ntiles) { fetch(col, tile + 1); pending = true; }
      else if (nntiles > 0) { fetch(ncol, 0); pending = true; }      // crosses the barrier: the operator is read-only
      double a = 0.0, x0 = 0.0;
#pragma unroll
      for (int t = 0; t < ITER; t++) {
        const double xv = c[t] >= 0 ? xs[c[t]] : 0.0;
        if (t == 0) x0 = xv;
        if (t == 0 && gl == 0) continue;
        a = __dsub_rn(a, __dmul_rn(v[t], xv));
      }
      a = group_sum<LPR>(a, gmask);
      if (gl == 0 && c[0] >= 0) {
        double xi = __dadd_rn(a, bs[c[0]]);
        xi = __dmul_rn(xi, omega / v[0]);
        xi = __dadd_rn(xi, __dmul_rn(1 - omega, x0));
        xs[c[0]] = xi;
      }
    }
    if (!pending && nntiles > 0) { fetch(ncol, 0); pending = true; }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < xlen; i += kSmallThreads) x[i] = xs[i];
}

// Coarsest levels (<= kResidentMaxBytes of packed operator): the WHOLE operator is copied into shared memory once per call,
// next to values_ and source_, so a colour phase costs a shared-memory round trip and a block barrier (~0.1 us) instead of an
// L2 round trip (~1 us: k_sor_mc_small above can only keep one tile of prefetch in flight).  Same per-row arithmetic, same bits.
constexpr size_t kResidentMaxBytes = 200 * 1024;

template <int LPR, int ITER>
__global__ void __launch_bounds__(kSmallThreads, 1) k_sor_mc_resident(const unsigned char* __restrict__ chunks, size_t chunk_bytes, int W, int total_rows,
                                                                      const int* __restrict__ colour_ptr, int ncolours, int iters,
                                                                      const double* __restrict__ b, double* __restrict__ x, double omega, int xlen) {
  extern __shared__ __align__(16) double small_smem[];
  double* xs = small_smem;
  double* bs = small_smem + xlen;
  unsigned char* mat = reinterpret_cast<unsigned char*>(small_smem + 2 * (size_t)xlen);
  for (int i = threadIdx.x; i < xlen; i += kSmallThreads) { xs[i] = x[i]; bs[i] = b[i]; }
  {
    const uint4* src = reinterpret_cast<const uint4*>(chunks);
    uint4* dst = reinterpret_cast<uint4*>(mat);
    const size_t n16 = (size_t)total_rows * chunk_bytes / 16;
    for (size_t i = threadIdx.x; i < n16; i += kSmallThreads) dst[i] = src[i];
  }
  const int lane = threadIdx.x & 31, gl = lane % LPR;
  const unsigned gmask = group_mask<LPR>(lane);
  constexpr int GPW = 32 / LPR;
  constexpr int TR = (kSmallThreads / 32) * GPW;
  const int slot = (threadIdx.x >> 5) * GPW + lane / LPR;
  const double om1 = 1 - omega;
  __syncthreads();
  const int nphases = iters * ncolours;
  for (int p = 0; p < nphases; p++) {
    const int col = p % ncolours;
    const int first = colour_ptr[col], count = colour_ptr[col + 1] - first;
    for (int i = slot; i < count; i += TR) {            // warp-uniform trip count is not needed: no barrier inside
      const double* pv = reinterpret_cast<const double*>(mat + (size_t)(first + i) * chunk_bytes);
      const int* pc = reinterpret_cast<const int*>(pv + W);
      double v[ITER], xv[ITER];
      int c[ITER];
#pragma unroll
      for (int t = 0; t < ITER; t++) {
        const int k = gl + t * LPR;
        const bool ok = k < W;
        v[t] = ok ? pv[k] : 0.0;
        c[t] = ok ? (pc[k] & 0x3fffffff) : -1;
        xv[t] = ok ? xs[c[t]] : 0.0;
      }
      double a = 0.0;
#pragma unroll
      for (int t = 0; t < ITER; t++) {
        if (t == 0 && gl == 0) continue;
        a = __dsub_rn(a, __dmul_rn(v[t], xv[t]));
      }
      a = group_sum<LPR>(a, gmask);
      if (gl == 0) {
        double xi = __dadd_rn(a, bs[c[0]]);
        xi = __dmul_rn(xi, omega / v[0]);
        xi = __dadd_rn(xi, __dmul_rn(om1, xv[0]));
        xs[c[0]] = xi;
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < xlen; i += kSmallThreads) x[i] = xs[i];
}

__global__ void k_scatter(const int* __restrict__ idx, const double* __restrict__ vals, int count, double* dst, int use_zero) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[idx[i]] = use_zero ? 0.0 : vals[i];
}
__global__ void k_set_one(double* dst, int i, double v) { dst[i] = v; }

int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
int lanes_for_width(int W) { return W >= 48 ? 32 : (W >= 24 ? 16 : 8); }

int grid_for(int work_groups_rows, int lpr, int sm_count) {
  // rows -> warps -> blocks; cap at a few waves of resident CTAs (persistent grid-stride beyond that)
  const int gpw = 32 / lpr;
  const long long warps = ((long long)work_groups_rows + gpw - 1) / gpw;
  long long blocks = (warps * 32 + kBlock - 1) / kBlock;
  const long long cap = (long long)sm_count * 8 * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

int sm_count_of(int device) {
  static int cached[64] = {0};
  if (device < 64 && cached[device]) return cached[device];
  int n = 0;
  MMG_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
  if (device < 64) cached[device] = n;
  return n;
}

void ensure_partials(Grid& g, size_t n) {
  if (g.partials.n < n) g.partials.alloc(n);
}

template <class F>
void dispatch_lpr(int W, F&& f) {
  const int lpr = lanes_for_width(W);
  if (lpr == 32) f(std::integral_constant<int, 32>());
  else if (lpr == 16) f(std::integral_constant<int, 16>());
  else f(std::integral_constant<int, 8>());
}

// (lanes per row, entries per lane) for the second-generation fast kernels; false if the width is outside the table
template <class F>
bool dispatch_lpr_iter(int W, F&& f, int prefer_lpr = 0) {
  int lpr = prefer_lpr ? prefer_lpr : lanes_for_width(W);
  { static int forced = -2; if (forced == -2) { const char* e = getenv("MMG_FAST_LPR"); forced = e ? atoi(e) : 0; } if (forced == 8 || forced == 16 || forced == 32) lpr = forced; }
  const int iter = (W + lpr - 1) / lpr;
#define MMG_CASE(L_, I_) if (lpr == L_ && iter == I_) { f(std::integral_constant<int, L_>(), std::integral_constant<int, I_>()); return true; }
  MMG_CASE(32, 2) MMG_CASE(32, 3) MMG_CASE(32, 4)
  MMG_CASE(16, 2) MMG_CASE(16, 3)
  MMG_CASE(8, 1) MMG_CASE(8, 2) MMG_CASE(8, 3) MMG_CASE(8, 4) MMG_CASE(8, 5)
#undef MMG_CASE
  return false;
}
constexpr int kSpmvRows = MMG_FAST_ROWS;   // rows in flight per lane group: 4 is best for the streaming SpMV kernels (profiles/r01_kernel_rates.txt)
constexpr int kMcRows = 2;                 // ... and 2 for the multicolour sweep, whose gathers do not coalesce
int grid_for2(int rows, int lpr, int sm_count, int per = kSpmvRows) { return grid_for((rows + per - 1) / per, lpr, sm_count); }

void launch_spmv(const HybMatrix& M, const double* x, const double* b, double* y, const unsigned char* rowflag, int op, int mask_d, int mask_n,
                 double* partial, int* nblocks_out, int device, cudaStream_t s, bool exact, int row0 = 0, int nrows = -1) {
  const int sms = sm_count_of(device);
  if (nrows < 0) nrows = M.rows;
  MMG_REQUIRE(!exact || (row0 == 0 && nrows == M.rows), MMG_ERR_STATE, "row-block launches exist for the fast-arithmetic kernels only");
  if (exact) {
    const int blocks = grid_for(M.rows, 32, sms);
    if (nblocks_out) *nblocks_out = blocks;
    k_spmv_exact<<<blocks, kBlock, 0, s>>>(M.view(), M.diag_first ? 1 : 0, x, b, y, rowflag, op, mask_d, mask_n, partial);
    MMG_CUDA(cudaGetLastError());
    note_kernel_slot(1, "k_spmv_exact");
    return;
  }
  if (nrows >= env_int("MMG_SPMV_TMA_MIN_ROWS", 30000) && env_int("MMG_SPMV_TMA", 1) &&
      stream_spmv(M, x, b, y, rowflag, op, mask_d, mask_n, partial, nblocks_out, device, s, row0, nrows))
    return;
  const bool done = dispatch_lpr_iter(M.W, [&](auto L, auto I) {
    constexpr int LPR = decltype(L)::value, ITER = decltype(I)::value;
    int blocks = grid_for2(nrows, LPR, sms);
    if (nblocks_out) *nblocks_out = blocks;
    note_kernel_slot(1, "k_spmv2", LPR, ITER, kSpmvRows);
    k_spmv2<LPR, ITER, kSpmvRows><<<blocks, kBlock, 0, s>>>(M.view(), x, b, y, rowflag, op, mask_d, mask_n, partial, M.cols, row0, nrows);
  });
  MMG_REQUIRE(done || (row0 == 0 && nrows == M.rows), MMG_ERR_STATE, "row-block launch: stencil width outside the fast kernels' table");
  if (!done)
    dispatch_lpr(M.W, [&](auto L) {
      constexpr int LPR = decltype(L)::value;
      const int blocks = grid_for(M.rows, LPR, sms);
      if (nblocks_out) *nblocks_out = blocks;
      k_spmv<LPR><<<blocks, kBlock, 0, s>>>(M.view(), x, b, y, rowflag, op, mask_d, mask_n, partial);
    });
  MMG_CUDA(cudaGetLastError());
}

constexpr int kRegBlocks = 296;

}  // namespace

// ------------------------------------------------------------------------------------------------
// timers
// ------------------------------------------------------------------------------------------------
TimedScope::TimedScope(Grid& g, int cls_, int64_t bytes, int launches) : t(g.timers), cls(cls_), s(g.stream) {
  if (!t) return;
  const int lv = g.level < kMaxLevels ? g.level : kMaxLevels - 1;
  cls = lv * MMG_T_COUNT + cls_;
  t->total_launches += launches;
  t->launches[lv][cls_] += launches;
  t->bytes[lv][cls_] += bytes;
  if (!t->on) return;
  auto get = [&]() {
    cudaEvent_t e;
    if (!t->pool.empty()) { e = t->pool.back(); t->pool.pop_back(); }
    else MMG_CUDA(cudaEventCreate(&e));
    return e;
  };
  e0 = get(); e1 = get();
  MMG_CUDA(cudaEventRecord(e0, s));
}
TimedScope::~TimedScope() {
  if (!t || !t->on || !e0) return;
  cudaEventRecord(e1, s);
  t->pending.push_back({cls, {e0, e1}});
}
void timers_collect(Timers& t) {
  for (auto& p : t.pending) {
    float ms = 0;
    cudaEventSynchronize(p.second.second);
    cudaEventElapsedTime(&ms, p.second.first, p.second.second);
    t.ms[p.first / MMG_T_COUNT][p.first % MMG_T_COUNT] += ms;
    t.pool.push_back(p.second.first);
    t.pool.push_back(p.second.second);
  }
  t.pending.clear();
}

// ------------------------------------------------------------------------------------------------
// host CSR <-> device HYB
// ------------------------------------------------------------------------------------------------
void hyb_from_csr(HybMatrix& M, const HostCsr& A, bool diag_first, bool has_reg, cudaStream_t s) {
  const int R = has_reg ? A.rows - 1 : A.rows;
  M.rows = R; M.cols = A.cols; M.diag_first = diag_first; M.nnz = A.nnz();
  std::vector<int> len(R);
  int maxlen = 0;
  for (int r = 0; r < R; r++) { len[r] = A.ptr[r + 1] - A.ptr[r]; maxlen = std::max(maxlen, len[r]); }
  int W = maxlen;
  if (R > 0) {
    std::vector<int> sorted(len);
    std::nth_element(sorted.begin(), sorted.begin() + R / 2, sorted.end());
    const int med = sorted[R / 2];
    if (maxlen > med + med / 4) {  // a few long rows (implicit-Neumann fill-in): spill them instead of padding every row
      const size_t q = (size_t)(0.985 * (R - 1));
      std::nth_element(sorted.begin(), sorted.begin() + q, sorted.end());
      W = std::max(sorted[q], med);
    }
  }
  if (W < 1) W = 1;
  M.W = W;
  M.chunk_bytes = ((size_t)W * 12 + 31) / 32 * 32;
  std::vector<unsigned char> chunks((size_t)R * M.chunk_bytes, 0);
  std::vector<int> ovf_rows, ovf_ptr(1, 0), ovf_col;
  std::vector<double> ovf_val;
  std::vector<std::pair<int, double>> row;
  for (int r = 0; r < R; r++) {
    row.clear();
    int dpos = -1;
    for (int k = A.ptr[r]; k < A.ptr[r + 1]; k++) {
      if (diag_first && A.idx[k] == r && dpos < 0) dpos = (int)row.size();
      row.push_back({A.idx[k], A.val[k]});
    }
    if (diag_first) {
      MMG_REQUIRE(dpos >= 0 || len[r] == 0, MMG_ERR_ARG, "row " + std::to_string(r) + " of the Laplacian has no diagonal entry");
      if (dpos > 0) std::rotate(row.begin(), row.begin() + dpos, row.begin() + dpos + 1);
    }
    double* v = reinterpret_cast<double*>(chunks.data() + (size_t)r * M.chunk_bytes);
    int* c = reinterpret_cast<int*>(chunks.data() + (size_t)r * M.chunk_bytes + (size_t)W * 8);
    const int m = std::min(len[r], W);
    for (int k = 0; k < m; k++) { v[k] = row[k].second; c[k] = row[k].first; }
    for (int k = m; k < W; k++) { v[k] = 0.0; c[k] = r < A.cols ? r : 0; }
    if (len[r] > W) {
      ovf_rows.push_back(r);
      for (int k = W; k < len[r]; k++) { ovf_col.push_back(row[k].first); ovf_val.push_back(row[k].second); }
      ovf_ptr.push_back((int)ovf_col.size());
    }
  }
  M.chunks.upload(chunks, s);
  M.len.upload(len, s);
  M.n_ovf = (int)ovf_rows.size();
  if (M.n_ovf) {
    M.ovf_rows.upload(ovf_rows, s); M.ovf_ptr.upload(ovf_ptr, s); M.ovf_col.upload(ovf_col, s); M.ovf_val.upload(ovf_val, s);
  }
  M.reg_row = -1; M.reg_len = 0;
  if (has_reg) {
    const int r = A.rows - 1;
    std::vector<int> rc; std::vector<double> rv;
    M.reg_diag = 0;
    for (int k = A.ptr[r]; k < A.ptr[r + 1]; k++) {
      if (A.idx[k] == r) M.reg_diag = A.val[k];
      else { rc.push_back(A.idx[k]); rv.push_back(A.val[k]); }
    }
    M.reg_row = r; M.reg_len = (int)rc.size();
    M.reg_col.upload(rc, s); M.reg_val.upload(rv, s);
  }
  MMG_CUDA(cudaStreamSynchronize(s));  // host staging buffers die here
}

void hyb_to_csr(const HybMatrix& M, HostCsr& A, cudaStream_t s) {
  const int R = M.rows;
  std::vector<unsigned char> chunks = M.chunks.to_host(s);
  std::vector<int> len = M.len.to_host(s);
  std::vector<int> ovf_rows, ovf_ptr, ovf_col; std::vector<double> ovf_val;
  if (M.n_ovf) { ovf_rows = M.ovf_rows.to_host(s); ovf_ptr = M.ovf_ptr.to_host(s); ovf_col = M.ovf_col.to_host(s); ovf_val = M.ovf_val.to_host(s); }
  A.rows = R + (M.reg_row >= 0 ? 1 : 0); A.cols = M.cols;
  A.ptr.assign(A.rows + 1, 0); A.idx.clear(); A.val.clear();
  std::vector<std::pair<int, double>> row;
  size_t o = 0;
  for (int r = 0; r < R; r++) {
    row.clear();
    const double* v = reinterpret_cast<const double*>(chunks.data() + (size_t)r * M.chunk_bytes);
    const int* c = reinterpret_cast<const int*>(chunks.data() + (size_t)r * M.chunk_bytes + (size_t)M.W * 8);
    const int m = std::min(len[r], M.W);
    for (int k = 0; k < m; k++) row.push_back({c[k], v[k]});
    if (len[r] > M.W) {
      for (int k = ovf_ptr[o]; k < ovf_ptr[o + 1]; k++) row.push_back({ovf_col[k], ovf_val[k]});
      o++;
    }
    if (M.diag_first && !row.empty()) {  // put the diagonal back in column order
      auto d = row.front();
      row.erase(row.begin());
      auto it = std::lower_bound(row.begin(), row.end(), d, [](const std::pair<int, double>& a, const std::pair<int, double>& b) { return a.first < b.first; });
      row.insert(it, d);
    }
    for (auto& e : row) { A.idx.push_back(e.first); A.val.push_back(e.second); }
    A.ptr[r + 1] = (int)A.idx.size();
  }
  if (M.reg_row >= 0) {
    std::vector<int> rc = M.reg_col.to_host(s); std::vector<double> rv = M.reg_val.to_host(s);
    bool placed = false;
    for (size_t k = 0; k < rc.size(); k++) {
      if (!placed && rc[k] > M.reg_row) { A.idx.push_back(M.reg_row); A.val.push_back(M.reg_diag); placed = true; }
      A.idx.push_back(rc[k]); A.val.push_back(rv[k]);
    }
    if (!placed) { A.idx.push_back(M.reg_row); A.val.push_back(M.reg_diag); }
    A.ptr[R + 1] = (int)A.idx.size();
  }
}

// ------------------------------------------------------------------------------------------------
// operations
// ------------------------------------------------------------------------------------------------
// off-diagonal dot product of the regularisation row; returns how many partials were written
static int launch_regdot(Grid& g, const double* x, double* reg_partial) {
  const HybMatrix& L = g.Lap;
  if (L.reg_row < 0) return 0;
  if (g.exact) {
    if (g.reg_prod.n < (size_t)L.reg_len) g.reg_prod.alloc((size_t)L.reg_len);
    k_regdot_products<<<std::max(1, std::min(kRegBlocks * 4, (L.reg_len + kBlock - 1) / kBlock)), kBlock, 0, g.stream>>>(L.reg_col.p, L.reg_val.p, L.reg_len, x,
                                                                                                                     g.reg_prod.p);
    k_regdot_exact<<<1, 32, 0, g.stream>>>(g.reg_prod.p, L.reg_len, reg_partial);
    MMG_CUDA(cudaGetLastError());
    return 1;
  }
  k_regdot<<<kRegBlocks, kBlock, 0, g.stream>>>(L.reg_col.p, L.reg_val.p, L.reg_len, x, reg_partial);
  MMG_CUDA(cudaGetLastError());
  return kRegBlocks;
}

static void residual_impl(Grid& g, double* r_out, double* ratio_dev) {
  const HybMatrix& L = g.Lap;
  MMG_REQUIRE(g.have_laplacian, MMG_ERR_STATE, "laplaceMat_ has not been built or uploaded");
  const int sms = sm_count_of(g.device);
  const int maxblocks = sms * 32 + 8;
  ensure_partials(g, (size_t)2 * maxblocks + kRegBlocks);
  double* norm_partial = g.partials.p;
  double* reg_partial = g.partials.p + 2 * maxblocks;
  int nb = 0;
  {
    TimedScope ts(g, MMG_T_RESIDUAL, L.matrix_bytes() + (int64_t)g.A * 24, 2 + (L.reg_row >= 0));
    launch_spmv(L, g.x.p, g.b.p, r_out, g.rowflag.p, OP_RESID, 0, 0, norm_partial, &nb, g.device, g.stream, g.exact);
    const int n_reg = launch_regdot(g, g.x.p, reg_partial);
    k_finish_residual<<<1, kBlock, 0, g.stream>>>(norm_partial, nb, reg_partial, n_reg, L.reg_row, L.reg_diag, g.x.p, g.b.p, r_out, ratio_dev);
    MMG_CUDA(cudaGetLastError());
  }
}

void op_residual(Grid& g, double* r_out) { residual_impl(g, r_out, nullptr); }
void op_residual_norm(Grid& g, double* ratio_dev) { residual_impl(g, nullptr, ratio_dev); }

void op_bound_eval_neumann(Grid& g) {
  if (g.neu_pts.n == 0) return;
  MMG_REQUIRE(g.have_laplacian, MMG_ERR_STATE, "laplaceMat_ has not been built or uploaded");
  const int sms = sm_count_of(g.device);
  TimedScope ts(g, MMG_T_OTHER, (int64_t)g.neu_pts.n * (12 * g.Lap.W + 24));
  if (g.exact) {
    k_bound_eval_exact<<<grid_for((int)g.neu_pts.n, 32, sms), kBlock, 0, g.stream>>>(g.Lap.view(), g.neu_pts.p, (int)g.neu_pts.n, g.b.p, g.x.p);
  } else {
    dispatch_lpr(g.Lap.W, [&](auto Lc) {
      constexpr int LPR = decltype(Lc)::value;
      k_bound_eval<LPR><<<grid_for((int)g.neu_pts.n, LPR, sms), kBlock, 0, g.stream>>>(g.Lap.view(), g.neu_pts.p, (int)g.neu_pts.n, g.b.p, g.x.p);
    });
  }
  MMG_CUDA(cudaGetLastError());
}

void op_boundary_op(Grid& g, int coarse) {  // grid.cpp:42-51
  if (g.dir_pts.n == 0) return;
  TimedScope ts(g, MMG_T_OTHER, (int64_t)g.dir_pts.n * 20);
  k_scatter<<<((int)g.dir_pts.n + kBlock - 1) / kBlock, kBlock, 0, g.stream>>>(g.dir_pts.p, g.dir_vals.p, (int)g.dir_pts.n, g.x.p, coarse);
  MMG_CUDA(cudaGetLastError());
}

void op_modify_coeff_neumann(Grid& g, int coarse) {  // grid.cpp:62-72
  TimedScope ts(g, MMG_T_OTHER, (int64_t)g.neu_pts.n * 20 + 8, 1 + (g.neu_pts.n ? 1 : 0));
  if (g.neu_pts.n) k_scatter<<<((int)g.neu_pts.n + kBlock - 1) / kBlock, kBlock, 0, g.stream>>>(g.neu_pts.p, g.neu_vals.p, (int)g.neu_pts.n, g.b.p, coarse);
  k_set_one<<<1, 1, 0, g.stream>>>(g.b.p, (int)g.b.n - 1, 0.0);
  MMG_CUDA(cudaGetLastError());
}

void op_fix_vector_bound_coarse(Grid& g, double* vec_dev) {  // grid.cpp:197-205
  if (g.dir_pts.n == 0) return;
  TimedScope ts(g, MMG_T_OTHER, (int64_t)g.dir_pts.n * 12);
  k_scatter<<<((int)g.dir_pts.n + kBlock - 1) / kBlock, kBlock, 0, g.stream>>>(g.dir_pts.p, nullptr, (int)g.dir_pts.n, vec_dev, 1);
  MMG_CUDA(cudaGetLastError());
}

void op_zero_values(Grid& g) {
  TimedScope ts(g, MMG_T_OTHER, (int64_t)g.A * 8);
  g.x.zero(g.stream);
}

void op_spmv(const HybMatrix& M, const double* x_dev, double* y_dev, Grid& ctx, int timer_class) {
  TimedScope ts(ctx, timer_class, M.matrix_bytes() + (int64_t)M.rows * 8 + (int64_t)M.cols * 8);
  launch_spmv(M, x_dev, nullptr, y_dev, nullptr, OP_SPMV, 0, 0, nullptr, nullptr, ctx.device, ctx.stream, ctx.exact);
}

void op_restrict(Grid& fine, Grid& coarse, const HybMatrix& R, const double* fine_res_dev) {
  MMG_REQUIRE(R.rows == coarse.n && R.cols == fine.n, MMG_ERR_STATE, "restriction matrix shape does not match the grids");
  {
    TimedScope ts(fine, MMG_T_RESTRICT, R.matrix_bytes() + (int64_t)coarse.n * 8 + (int64_t)fine.n * 8);
    // fix_vector_bound_coarse on the coarse source and, if the FINE grid is Neumann, modify_coeff_neumann("coarse"):
    // both are masks on the output rows (multigrid.cpp:82-86)
    launch_spmv(R, fine_res_dev, nullptr, coarse.b.p, coarse.rowflag.p, OP_RESTRICT, 1, fine.neumann ? 1 : 0, nullptr, nullptr, fine.device, fine.stream, fine.exact);
  }
  if (fine.neumann) {  // source_(rows-1)=0 (multigrid.cpp:84) and the same store inside modify_coeff_neumann (grid.cpp:71)
    TimedScope ts(fine, MMG_T_OTHER, 8);
    k_set_one<<<1, 1, 0, fine.stream>>>(coarse.b.p, (int)coarse.b.n - 1, 0.0);
    MMG_CUDA(cudaGetLastError());
  }
}

void op_prolong_correct(Grid& fine, Grid& coarse, const HybMatrix& P) {
  MMG_REQUIRE(P.rows == fine.n && P.cols == coarse.n, MMG_ERR_STATE, "prolongation matrix shape does not match the grids");
  TimedScope ts(fine, MMG_T_PROLONG, P.matrix_bytes() + (int64_t)fine.n * 16 + (int64_t)coarse.n * 8);
  launch_spmv(P, coarse.x.p, nullptr, fine.x.p, fine.rowflag.p, OP_PROLONG, fine.neumann ? 0 : 1, 0, nullptr, nullptr, fine.device, fine.stream, fine.exact);
}

// ---- SOR -----------------------------------------------------------------------------------------
static int lex_block_cap(int sms) { static int v = -2; if (v == -2) v = env_int("MMG_LEX_BLOCKS", 0); return v > 0 ? v : 2 * sms; }
static unsigned lex_sleep_ns() { static int v = -2; if (v == -2) v = env_int("MMG_LEX_SLEEP_NS", 0); return (unsigned)v; }

template <int T, int K>
static void launch_lex_chunk(Grid& g, size_t stride) {
  const int iters = g.props.iters;
  { const int on = env_int("MMG_LEX_TRACE", 0); MMG_CUDA(cudaMemcpyToSymbolAsync(g_lex_trace_on, &on, sizeof(int), 0, cudaMemcpyHostToDevice, g.stream)); }
  int blocks_per_sm = 0;
  MMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_sor_lex_chunk<T, K>, K * 32, 0));
  const int sms = sm_count_of(g.device);
  const bool neu = g.neumann;
  int blocks = std::min(blocks_per_sm * sms, env_int("MMG_LEX_CHUNK_BLOCKS", 1 << 20)) - (neu ? iters : 0);   // room for one auxiliary CTA per sweep
  const int need = ((g.Lap.rows + K - 1) / K) * iters;
  if (blocks > need) blocks = need;
  if (blocks < iters) blocks = iters;
  int Q = blocks / iters;
  blocks = Q * iters + (neu ? iters : 0);
  LexNeumann nm{};
  if (neu) nm = LexNeumann{1, g.neu_pts.p, (int)g.neu_pts.n, g.Lap.reg_col.p, g.Lap.reg_val.p, g.Lap.reg_len, g.Lap.reg_row, g.Lap.reg_diag};
  HybView A = g.Lap.view();
  const unsigned char* rf = g.rowflag.p;
  const double* b = g.b.p;
  double* xs = g.xs.p;
  size_t st = stride;
  int it = iters;
  double omega = g.props.omega;
  int* abortp = g.abort_flag.p;
  long long timeout = 6000000000ll;
  void* args[] = {&A, &rf, &b, &xs, &st, &it, &omega, &abortp, &timeout, &Q, &nm};
  note_kernel(g, "k_sor_lex_chunk", T, K);
  MMG_CUDA(cudaLaunchCooperativeKernel((void*)k_sor_lex_chunk<T, K>, dim3(blocks), dim3(K * 32), args, 0, g.stream));
}

template <int T>
static void launch_lex_pipe(Grid& g) {
  const int iters = g.props.iters;
  const size_t stride = ((size_t)g.A + 63) / 64 * 64;
  if (g.xs.n < stride * (iters + 1)) g.xs.alloc(stride * (iters + 1));
  k_pipe_init<<<(g.A + kBlock - 1) / kBlock, kBlock, 0, g.stream>>>(g.rowflag.p, g.x.p, g.xs.p, stride, iters, g.A, g.neumann ? 1 : 0);
  MMG_CUDA(cudaGetLastError());
  const int chunk = g.neumann ? 32 : env_int("MMG_LEX_CHUNK", 32);  // rows per CTA chunk of the chunked sweep (0 = warp-per-row pipelined kernel), DESIGN.md §5
  if constexpr (T > 4) { MMG_REQUIRE(!g.neumann, MMG_ERR_STATE, "pipelined lexicographic sweep on a Neumann-type grid: stencil too wide"); }
  if constexpr (T <= 4) {
    if (chunk > 0) {
      if (chunk == 8) launch_lex_chunk<T, 8>(g, stride);
      else if (chunk == 32) launch_lex_chunk<T, 32>(g, stride);
      else launch_lex_chunk<T, 16>(g, stride);
      MMG_CUDA(cudaMemcpyAsync(g.x.p, g.xs.p + (size_t)iters * stride, sizeof(double) * g.A, cudaMemcpyDeviceToDevice, g.stream));
      return;
    }
  }
  int blocks_per_sm = 0;
  MMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_sor_lex_pipe<T>, kBlock, 0));
  const int sms = sm_count_of(g.device);
  int blocks = std::min(blocks_per_sm * sms, lex_block_cap(sms) * std::min(iters, 3));
  const int need = grid_for(g.Lap.rows, 32, 1 << 20) * iters;
  if (blocks > need) blocks = need;
  if (blocks * (kBlock / 32) < iters) blocks = (iters + kBlock / 32 - 1) / (kBlock / 32);
  HybView A = g.Lap.view();
  const unsigned char* rf = g.rowflag.p;
  const double* b = g.b.p;
  double* xs = g.xs.p;
  size_t st = stride;
  int it = iters;
  double omega = g.props.omega;
  int* abortp = g.abort_flag.p;
  long long timeout = 6000000000ll;
  unsigned sleep_ns = lex_sleep_ns();
  void* args[] = {&A, &rf, &b, &xs, &st, &it, &omega, &abortp, &timeout, &sleep_ns};
  note_kernel(g, "k_sor_lex_pipe", T);
  MMG_CUDA(cudaLaunchCooperativeKernel((void*)k_sor_lex_pipe<T>, dim3(blocks), dim3(kBlock), args, 0, g.stream));
  MMG_CUDA(cudaMemcpyAsync(g.x.p, g.xs.p + (size_t)iters * stride, sizeof(double) * g.A, cudaMemcpyDeviceToDevice, g.stream));
}

// A Neumann-type grid runs in the pipelined kernel when its stencil fits the chunked instantiations (W-1 <= 128 entries), the
// operator is stored diagonal-first and no Neumann boundary row spills into the overflow CSR (they never do: a boundary row has
// exactly n entries, grid.cpp:520-548).
static bool lex_neumann_pipelinable(Grid& g) {
  if (g.lex_neu_ok < 0) {
    bool ok = g.Lap.reg_row >= 0 && g.Lap.diag_first && g.Lap.W - 1 <= 128 && env_int("MMG_LEX_NEUMANN_PIPE", 1);
    if (ok && g.Lap.n_ovf) {
      std::vector<int> ovf = g.Lap.ovf_rows.to_host(g.stream);
      for (int r : ovf) if (g.bcflags[r] == 2) { ok = false; break; }
    }
    g.lex_neu_ok = ok ? 1 : 0;
  }
  return g.lex_neu_ok == 1;
}

// all props.iters lexicographic sweeps of a grid in one pipelined launch
static bool sor_lex_pipelined(Grid& g) {
  const int Te = (g.Lap.W - 1 + 31) / 32;
  if (Te <= 1) launch_lex_pipe<1>(g);
  else if (Te <= 2) launch_lex_pipe<2>(g);
  else if (Te <= 3) launch_lex_pipe<3>(g);
  else if (Te <= 4) launch_lex_pipe<4>(g);
  else if (Te <= 6) launch_lex_pipe<6>(g);
  else if (Te <= 8) launch_lex_pipe<8>(g);
  else return false;
  return true;
}

template <class K>
static void launch_lex_kernel(Grid& g, K kernel, int LPR, const double* x_old, double* x_new) {
  int blocks_per_sm = 0;
  MMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, kBlock, 0));
  const int sms = sm_count_of(g.device);
  int blocks = blocks_per_sm * sms;
  const int cap = lex_block_cap(sms);            // fewer pollers: the sweep is latency bound, idle warps only load L2
  if (blocks > cap) blocks = cap;
  const int need = grid_for(g.Lap.rows, LPR, 1 << 20);
  if (blocks > need) blocks = need;
  HybView A = g.Lap.view();
  const unsigned char* rf = g.rowflag.p;
  const double* b = g.b.p;
  double omega = g.props.omega;
  int* abortp = g.abort_flag.p;
  long long timeout = 6000000000ll;  // ~3 s of SM clocks
  void* args[] = {&A, &rf, &b, &x_old, &x_new, &omega, &abortp, &timeout};
  MMG_CUDA(cudaLaunchCooperativeKernel((void*)kernel, dim3(blocks), dim3(kBlock), args, 0, g.stream));
}
template <int T>
static void launch_lex_exact(Grid& g, const double* x_old, double* x_new) { launch_lex_kernel(g, k_sor_lex_exact<T>, 32, x_old, x_new); }

static void sor_lex_sweep(Grid& g) {
  const HybMatrix& L = g.Lap;
  const int W = L.W;
  const double* x_old = g.x.p;
  double* x_new = g.x_alt.p;
  k_sweep_init<<<(g.A + kBlock - 1) / kBlock, kBlock, 0, g.stream>>>(g.rowflag.p, x_old, x_new, L.rows, g.A);
  MMG_CUDA(cudaGetLastError());
  // The row kernel folds in the reference's column order in both arithmetic modes (a lexicographic sweep is latency bound, the
  // in-order fold costs nothing extra); MMG_ARITH_FAST only replaces the in-order sum of the dense regularisation row -- one
  // warp adding ~N terms one after the other -- by a tree reduction (launch_regdot).
  const int Te = (W - 1 + 31) / 32;
#define LEXE_CASE(T_) if (Te <= T_) { launch_lex_exact<T_>(g, x_old, x_new); } else
  LEXE_CASE(1) LEXE_CASE(2) LEXE_CASE(3) LEXE_CASE(4) LEXE_CASE(6) LEXE_CASE(8)
  { throw Error(MMG_ERR_ARG, "stencil width " + std::to_string(W) + " is outside the lexicographic kernel's dispatch table"); }
#undef LEXE_CASE
  note_kernel(g, "k_sor_lex_exact", Te);
  if (L.reg_row >= 0) {
    double* reg_partial = g.partials.p;
    const int n_reg = launch_regdot(g, x_new, reg_partial);
    k_finish_sor_reg<<<1, kBlock, 0, g.stream>>>(reg_partial, n_reg, L.reg_row, L.reg_diag, g.props.omega, g.b.p, x_old, x_new);
    MMG_CUDA(cudaGetLastError());
  }
  std::swap(g.x.p, g.x_alt.p);
}

// block colouring: adjacency bitmap on the device, first-fit in ascending block order on the host (nb is small)
void build_block_colouring(Grid& g) {
  const HybMatrix& L = g.Lap;
  const int B = g.block_size, R = L.rows;
  MMG_REQUIRE(B >= 32, MMG_ERR_ARG, "block size must be at least 32 rows");
  const int nb = (R + B - 1) / B, words = (nb + 31) / 32;
  DevBuf<unsigned> bitmap;
  bitmap.alloc((size_t)nb * words);
  bitmap.zero(g.stream);
  k_blk_adjacency<<<grid_for(R, 32, sm_count_of(g.device)), kBlock, 0, g.stream>>>(L.view(), g.rowflag.p, B, nb, bitmap.p);
  MMG_CUDA(cudaGetLastError());
  std::vector<unsigned> bm = bitmap.to_host(g.stream);
  auto adjacent = [&](int a, int b) { return ((bm[(size_t)a * words + (b >> 5)] >> (b & 31)) & 1u) || ((bm[(size_t)b * words + (a >> 5)] >> (a & 31)) & 1u); };
  g.blk_colour.assign(nb, -1);
  int ncol = 0;
  std::vector<int> mark;
  for (int a = 0; a < nb; a++) {
    for (int b = 0; b < a; b++)
      if (adjacent(a, b)) {
        if ((int)mark.size() <= g.blk_colour[b]) mark.resize(g.blk_colour[b] + 1, -1);
        mark[g.blk_colour[b]] = a;
      }
    int c = 0;
    while (c < (int)mark.size() && mark[c] == a) c++;
    g.blk_colour[a] = c;
    ncol = std::max(ncol, c + 1);
  }
  g.n_blk_colours = ncol;
  g.blk_phase_ptr.assign(ncol + 1, 0);
  for (int a = 0; a < nb; a++) g.blk_phase_ptr[g.blk_colour[a] + 1]++;
  for (int c = 0; c < ncol; c++) g.blk_phase_ptr[c + 1] += g.blk_phase_ptr[c];
  std::vector<int> order(nb), pos(g.blk_phase_ptr.begin(), g.blk_phase_ptr.end() - 1);
  for (int a = 0; a < nb; a++) order[pos[g.blk_colour[a]]++] = a;
  g.blk_phase_blocks.upload(order, g.stream);
  g.sync();
  g.have_blocks = true;
}

template <int T>
static void launch_blk(Grid& g, const int* phase_blocks, int nblk) {
  int blocks_per_sm = 0;
  MMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_sor_blk<T>, kBlock, 0));
  const int sms = sm_count_of(g.device);
  int blocks = std::min(blocks_per_sm * sms, env_int("MMG_BLK_BLOCKS", 6 * sms));
  const long long need = ((long long)nblk * g.block_size * 32 + kBlock - 1) / kBlock;
  if (blocks > need) blocks = (int)need;
  HybView A = g.Lap.view();
  const unsigned char* rf = g.rowflag.p;
  const double* b = g.b.p;
  const double* x = g.x.p;
  double* xw = g.x_alt.p;
  int B = g.block_size;
  double omega = g.props.omega;
  int* abortp = g.abort_flag.p;
  long long timeout = 6000000000ll;
  void* args[] = {&A, &rf, &b, &x, &xw, &phase_blocks, &nblk, &B, &omega, &abortp, &timeout};
  MMG_CUDA(cudaLaunchCooperativeKernel((void*)k_sor_blk<T>, dim3(blocks), dim3(kBlock), args, 0, g.stream));
}

static void sor_blk_sweep(Grid& g) {
  const HybMatrix& L = g.Lap;
  const int B = g.block_size;
  const int Te = (L.W - 1 + 31) / 32;
  for (int c = 0; c < g.n_blk_colours; c++) {
    const int first = g.blk_phase_ptr[c], nblk = g.blk_phase_ptr[c + 1] - first;
    if (nblk == 0) continue;
    const int* pb = g.blk_phase_blocks.p + first;
    const long long total = (long long)nblk * B;
    const int nb = (int)((total + kBlock - 1) / kBlock);
    k_blk_init<<<nb, kBlock, 0, g.stream>>>(g.rowflag.p, g.x.p, g.x_alt.p, pb, nblk, B, L.rows);
    if (Te <= 1) launch_blk<1>(g, pb, nblk);
    else if (Te <= 2) launch_blk<2>(g, pb, nblk);
    else if (Te <= 3) launch_blk<3>(g, pb, nblk);
    else if (Te <= 4) launch_blk<4>(g, pb, nblk);
    else if (Te <= 6) launch_blk<6>(g, pb, nblk);
    else if (Te <= 8) launch_blk<8>(g, pb, nblk);
    else throw Error(MMG_ERR_ARG, "stencil width outside the block-lexicographic kernel's dispatch table");
    k_blk_merge<<<nb, kBlock, 0, g.stream>>>(g.rowflag.p, g.x.p, g.x_alt.p, pb, nblk, B, L.rows);
    MMG_CUDA(cudaGetLastError());
  }
}

static void sor_mc_sweep(Grid& g) {
  const HybMatrix& L = g.Lap;
  const int sms = sm_count_of(g.device);
  const int ncol_rows = L.reg_row >= 0 ? g.n_colours - 1 : g.n_colours;  // the regularisation row is the last colour
  if (g.exact) {
    for (int c = 0; c < ncol_rows; c++) {
      const int first = g.colour_ptr[c], count = g.colour_ptr[c + 1] - first;
      if (count == 0) continue;
      k_sor_mc_exact<<<grid_for(count, 32, sms), kBlock, 0, g.stream>>>(L.view(), g.colour_rows.p + first, count, g.b.p, g.x.p, g.props.omega);
    }
  } else {
    const bool done = dispatch_lpr_iter(L.W, [&](auto Lc, auto I) {
      constexpr int LPR = decltype(Lc)::value, ITER = decltype(I)::value;
      for (int c = 0; c < ncol_rows; c++) {
        const int first = g.colour_ptr[c], count = g.colour_ptr[c + 1] - first;
        if (count == 0) continue;
        k_sor_mc2<LPR, ITER, kMcRows><<<grid_for2(count, LPR, sms, kMcRows), kBlock, 0, g.stream>>>(L.view(), g.colour_rows.p + first, count, g.b.p, g.x.p,
                                                                                                   g.props.omega, g.A);
      }
    });
    if (!done)
      dispatch_lpr(L.W, [&](auto Lc) {
        constexpr int LPR = decltype(Lc)::value;
        for (int c = 0; c < ncol_rows; c++) {
          const int first = g.colour_ptr[c], count = g.colour_ptr[c + 1] - first;
          if (count == 0) continue;
          k_sor_mc<LPR><<<grid_for(count, LPR, sms), kBlock, 0, g.stream>>>(L.view(), g.colour_rows.p + first, count, g.b.p, g.x.p, g.props.omega);
        }
      });
  }
  MMG_CUDA(cudaGetLastError());
  if (L.reg_row >= 0) {
    double* reg_partial = g.partials.p;
    const int n_reg = launch_regdot(g, g.x.p, reg_partial);
    k_finish_sor_reg<<<1, kBlock, 0, g.stream>>>(reg_partial, n_reg, L.reg_row, L.reg_diag, g.props.omega, g.b.p, g.x.p, g.x.p);
    MMG_CUDA(cudaGetLastError());
  }
}

// launch one of the register-fed whole-call kernels (grid barrier / barrier-free) over the packed copy
template <class Launch>
static bool launch_packed_family(Grid& g, Launch&& launch) {
  const HybMatrix& L = g.Lap;
  return dispatch_lpr_iter(L.W, launch, (g.A >= 200000 && L.W > 16 && L.W <= 40) ? 8 : 0);   // 8 lanes per row on the big levels: 4 rows per gather instruction share lines (+5 %)
}

void op_sor(Grid& g, int smoother) {
  MMG_REQUIRE(g.have_laplacian, MMG_ERR_STATE, "laplaceMat_ has not been built or uploaded");
  const HybMatrix& L = g.Lap;
  ensure_partials(g, (size_t)kRegBlocks + 2 * (sm_count_of(g.device) * 32 + 8));
  if (smoother == MMG_SMOOTHER_MULTICOLOUR && !g.have_colours) build_colouring(g);
  const int64_t call_bytes = (L.matrix_bytes() + (int64_t)g.A * 28) * g.props.iters;
  if (smoother == MMG_SMOOTHER_LEXICOGRAPHIC && g.props.iters >= 1 && !env_int("MMG_LEX_NO_PIPE", 0) &&
      ((!g.neumann && g.Lap.n_ovf == 0 && g.Lap.W <= 257) || (g.neumann && g.exact && lex_neumann_pipelinable(g)))) {
    // (Neumann-type grids with MMG_ARITH_FAST take the per-sweep path below: its tree-reduced regularisation row makes a sweep
    // 10 % shorter than the pipelined kernel's in-order chain -- measured 0.66 vs 0.73 s per V-cycle at 1M nodes)
    TimedScope ts(g, MMG_T_SOR, call_bytes, 3);
    sor_lex_pipelined(g);
    return;
  }
  if (smoother == MMG_SMOOTHER_BLOCK_LEXICOGRAPHIC) {
    MMG_REQUIRE(!g.neumann, MMG_ERR_STATE, "the block-lexicographic smoother is implemented for grids without Neumann boundaries");
    if (!g.have_blocks) build_block_colouring(g);
  }
  // Throughput mode: every colour phase of every sweep of the call in ONE launch over the colour-major packed copy.
  //   <= 512 rows              k_sor_mc_small   one CTA, vectors in shared memory
  //   <= MMG_MC_FLOW_MAX_ROWS  k_sor_mc_flow    barrier-free, the values are the ready flags (latency-bound levels)
  //   above, and every grid with a Neumann boundary: k_sor_mc_tma, the TMA-fed ring with one counter barrier per phase
  //   (mmg_stream.cu); k_sor_mc_packed (register-fed, grid.sync) is the fallback when the stencil width has no instantiation.
  if (smoother == MMG_SMOOTHER_MULTICOLOUR && !g.exact && L.diag_first && g.props.iters >= 1 && (int)g.hx.size() == g.n && g.mc_row1 < 0 &&
      env_int("MMG_MC_PACKED", 1)) {
    ensure_mc_pack(g);
    const bool plain = !g.neumann && L.n_ovf == 0;      // no regularisation row, no boundary evaluation, no overflow tails
    const size_t resident_bytes = (size_t)g.mc_colour_ptr.back() * L.chunk_bytes + (size_t)g.A * 16;
    if (plain && resident_bytes <= kResidentMaxBytes && env_int("MMG_MC_RESIDENT", 1)) {
      bool ok = false;
      {
        TimedScope ts(g, MMG_T_SOR, call_bytes, 1);
        ok = dispatch_lpr_iter(L.W, [&](auto Lc, auto I) {
          constexpr int LPR = decltype(Lc)::value, ITER = decltype(I)::value;
          auto kern = k_sor_mc_resident<LPR, ITER>;
          note_kernel(g, "k_sor_mc_resident", LPR, ITER);
          MMG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kResidentMaxBytes));
          kern<<<1, kSmallThreads, resident_bytes, g.stream>>>(g.mc_chunks.p, L.chunk_bytes, L.W, g.mc_colour_ptr.back(), g.mc_colour_ptr_dev.p, g.n_colours,
                                                                g.props.iters, g.b.p, g.x.p, g.props.omega, g.A);
          MMG_CUDA(cudaGetLastError());
        });
      }
      if (ok) return;
    }
    if (plain && g.A <= std::min(kSmallMaxRows, env_int("MMG_MC_SMALL_MAX", kSmallMaxRows)) && env_int("MMG_MC_SMALL", 1)) {
      bool ok = false;
      {
        TimedScope ts(g, MMG_T_SOR, call_bytes, 1);
        ok = dispatch_lpr_iter(L.W, [&](auto Lc, auto I) {
          constexpr int LPR = decltype(Lc)::value, ITER = decltype(I)::value;
          auto kern = k_sor_mc_small<LPR, ITER>;
          note_kernel(g, "k_sor_mc_small", LPR, ITER);
          const size_t smem = (size_t)g.A * 16;
          static_assert(kSmallMaxRows * 16 <= 48 * 1024, "beyond 48 KB the kernel needs cudaFuncAttributeMaxDynamicSharedMemorySize");
          kern<<<1, kSmallThreads, smem, g.stream>>>(g.mc_chunks.p, L.chunk_bytes, L.W, g.mc_colour_ptr_dev.p, g.n_colours, g.props.iters, g.b.p, g.x.p,
                                                      g.props.omega, g.A);
          MMG_CUDA(cudaGetLastError());
        });
      }
      if (ok) return;
    }
    if (plain && g.A <= env_int("MMG_MC_FLOW_MAX_ROWS", 1500000) && env_int("MMG_MC_FLOW", 1)) {
      bool ok = false;
      {
        TimedScope ts(g, MMG_T_SOR, call_bytes, 3);
        const int iters = g.props.iters;
        const size_t stride = ((size_t)g.A + 63) / 64 * 64;
        if (g.xs.n < stride * (iters + 1)) g.xs.alloc(stride * (iters + 1));
        if (g.A >= env_int("MMG_MC_TMAFLOW_MIN_ROWS", 2000) && env_int("MMG_MC_TMAFLOW", 1)) {     // the same sweep, operator fed through the TMA ring
          k_pipe_init<<<(g.A + kBlock - 1) / kBlock, kBlock, 0, g.stream>>>(g.rowflag.p, g.x.p, g.xs.p, stride, iters, g.A);
          MMG_CUDA(cudaGetLastError());
          ok = stream_sor_mc_flow(g, g.xs.p, stride, nullptr);
          if (ok) MMG_CUDA(cudaMemcpyAsync(g.x.p, g.xs.p + (size_t)iters * stride, sizeof(double) * g.A, cudaMemcpyDeviceToDevice, g.stream));
        }
        if (!ok) ok = launch_packed_family(g, [&](auto Lc, auto I) {
          constexpr int LPR = decltype(Lc)::value, ITER = decltype(I)::value;
          k_pipe_init<<<(g.A + kBlock - 1) / kBlock, kBlock, 0, g.stream>>>(g.rowflag.p, g.x.p, g.xs.p, stride, iters, g.A);
          MMG_CUDA(cudaGetLastError());
          const int rows_used = env_int("MMG_MC_FLOW_ROWS", 1) == 1 ? 1 : 2;   // latency-bound levels: one row per lane group keeps more warps resident
          void* kern = rows_used == 1 ? (void*)k_sor_mc_flow<LPR, ITER, 1, false> : (void*)k_sor_mc_flow<LPR, ITER, 2, false>;
          note_kernel(g, "k_sor_mc_flow", LPR, ITER, rows_used);
          int blocks_per_sm = 0;
          MMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kBlock, 0));
          const int sms = sm_count_of(g.device);
          int maxcount = 0;
          for (int c = 0; c < g.n_colours; c++) maxcount = std::max(maxcount, g.mc_colour_ptr[c + 1] - g.mc_colour_ptr[c]);
          int blocks = std::min(blocks_per_sm * sms, grid_for2(maxcount, LPR, sms, rows_used));
          const unsigned char* chunks = g.mc_chunks.p;
          size_t cb = L.chunk_bytes, st = stride;
          int W = L.W;
          const int* cp = g.mc_colour_ptr_dev.p;
          int nc = g.n_colours, itn = iters;
          const double* b = g.b.p;
          double* xs = g.xs.p;
          double omega = g.props.omega;
          int* abortp = g.abort_flag.p;
          long long timeout = 6000000000ll;
          PeerSends peers{};
          void* args[] = {&chunks, &cb, &W, &cp, &nc, &itn, &b, &xs, &st, &omega, &abortp, &timeout, &peers};
          MMG_CUDA(cudaLaunchCooperativeKernel(kern, dim3(blocks), dim3(kBlock), args, 0, g.stream));
          MMG_CUDA(cudaMemcpyAsync(g.x.p, g.xs.p + (size_t)iters * stride, sizeof(double) * g.A, cudaMemcpyDeviceToDevice, g.stream));
        });
      }
      if (ok) return;
    }
    if (env_int("MMG_MC_TMA", 1) || !plain) {
      bool ok = false;
      {
        TimedScope ts(g, MMG_T_SOR, call_bytes, 2);
        ok = stream_sor_mc(g);
      }
      if (ok) return;
    }
    if (plain) {
      bool done = false;
      {
        TimedScope ts(g, MMG_T_SOR, call_bytes, 1);
        const int rows_pref = env_int("MMG_MC_ROWS", 2);
        done = launch_packed_family(g, [&](auto Lc, auto I) {
          constexpr int LPR = decltype(Lc)::value, ITER = decltype(I)::value;
          void* kern = rows_pref >= 4 ? (void*)k_sor_mc_packed<LPR, ITER, 4> : rows_pref == 1 ? (void*)k_sor_mc_packed<LPR, ITER, 1> : (void*)k_sor_mc_packed<LPR, ITER, 2>;
          const int rows_used = rows_pref >= 4 ? 4 : rows_pref == 1 ? 1 : 2;
          note_kernel(g, "k_sor_mc_packed", LPR, ITER, rows_used);
          int blocks_per_sm = 0;
          MMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kBlock, 0));
          const int sms = sm_count_of(g.device);
          int maxcount = 0;
          for (int c = 0; c < g.n_colours; c++) maxcount = std::max(maxcount, g.mc_colour_ptr[c + 1] - g.mc_colour_ptr[c]);
          int blocks = std::min(blocks_per_sm * sms, grid_for2(maxcount, LPR, sms, rows_used));
          const unsigned char* chunks = g.mc_chunks.p;
          size_t cb = L.chunk_bytes;
          int W = L.W;
          const int* cp = g.mc_colour_ptr_dev.p;
          int nc = g.n_colours, iters = g.props.iters;
          const double* b = g.b.p;
          double* x = g.x.p;
          double omega = g.props.omega;
          void* args[] = {&chunks, &cb, &W, &cp, &nc, &iters, &b, &x, &omega};
          MMG_CUDA(cudaLaunchCooperativeKernel(kern, dim3(blocks), dim3(kBlock), args, 0, g.stream));
        });
      }
      if (done) return;
    }
  }
  // per-phase launches: the reference-order (parity) kernels, partitioned levels, and anything the fused paths do not cover
  for (int it = 0; it < g.props.iters; it++) {
    {
      const int launches = smoother == MMG_SMOOTHER_MULTICOLOUR ? g.n_colours + 1 : smoother == MMG_SMOOTHER_BLOCK_LEXICOGRAPHIC ? 3 * g.n_blk_colours : 2 + (L.reg_row >= 0 ? 2 : 0);
      TimedScope ts(g, MMG_T_SOR, L.matrix_bytes() + (int64_t)g.A * 28, launches);
      if (smoother == MMG_SMOOTHER_MULTICOLOUR) sor_mc_sweep(g);
      else if (smoother == MMG_SMOOTHER_BLOCK_LEXICOGRAPHIC) sor_blk_sweep(g);
      else sor_lex_sweep(g);
    }
    op_bound_eval_neumann(g);  // grid.cpp:144
  }
}

std::string& last_kernel_slot(int slot) {
  static thread_local std::string names[2];
  return names[slot & 1];
}

void debug_lex_trace(long long* out, int n) {
  MMG_CUDA(cudaDeviceSynchronize());
  MMG_CUDA(cudaMemcpyFromSymbol(out, g_lex_trace, sizeof(long long) * std::min(n, 16 * 64)));
}

// ================================================================================================
// Fractional-step explicit operators (fractionalStepGrid.cpp:101-154)
// ================================================================================================
namespace {
// u_hat = u + dt*(-(u*u_x + v*u_y) + mu/rho*del2_u)   (fractionalStepGrid.cpp:109, 121)
__global__ void k_fs_hat(int n, const double* __restrict__ w, const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ wx,
                         const double* __restrict__ wy, const double* __restrict__ lap, double dt, double mu_over_rho, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = w[i] + dt * (-(u[i] * wx[i] + v[i] * wy[i]) + mu_over_rho * lap[i]);
}
// source_.head(N) = rho/dt * (Dx u_hat + Dy v_hat)   (:127)
__global__ void k_fs_ppe_interior(int n, const double* __restrict__ a, const double* __restrict__ b, double rho_over_dt, double* __restrict__ src) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) src[i] = rho_over_dt * (a[i] + b[i]);
}
// boundary nodes: source_ = nx*dpdx + ny*dpdy, dpdx = -rho/dt*(u - u_hat)   (:132-143)
__global__ void k_fs_ppe_boundary(int count, const int* __restrict__ pts, const double* __restrict__ u, const double* __restrict__ v,
                                  const double* __restrict__ uh, const double* __restrict__ vh, const double* __restrict__ nx,
                                  const double* __restrict__ ny, double neg_rho_over_dt, double* src) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  const int c = pts[j];
  const double dpdx = neg_rho_over_dt * (u[c] - uh[c]);
  const double dpdy = neg_rho_over_dt * (v[c] - vh[c]);
  src[c] = nx[c] * dpdx + ny[c] * dpdy;
}
// u = u_hat - dt/rho * (Dx p)   (:147, 150)
__global__ void k_fs_correct(int n, const double* __restrict__ hat, const double* __restrict__ grad, double dt_over_rho, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = hat[i] - dt_over_rho * grad[i];
}
__global__ void __launch_bounds__(kBlock) k_fs_absdiff(int n, const double* __restrict__ a, const double* __restrict__ b, double* partial) {
  double s = 0.0, dummy = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += fabs(a[i] - b[i]);
  block_sum2(s, dummy);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
void fs_spmv(Grid& g, const HybMatrix& M, const double* x, double* y) {
  launch_spmv(M, x, nullptr, y, nullptr, OP_SPMV, 0, 0, nullptr, nullptr, g.device, g.stream, g.exact);
}
}  // namespace

void fs_calc_hat(Grid& g, int component) {
  Grid::FracStep& F = *g.fs;
  MMG_REQUIRE(F.have_ops, MMG_ERR_STATE, "the fractional-step operators have not been built or uploaded");
  const int N = g.n, nb = (N + kBlock - 1) / kBlock;
  TimedScope ts(g, MMG_T_OTHER, 6 * F.Dx.matrix_bytes() + (int64_t)N * 8 * 14, 8);
  for (int w = 0; w < 2; w++) {   // w = 0: u_hat from u; w = 1: v_hat from v
    if (component != MMG_FS_BOTH && component != w) continue;
    const double* field = F.vec[w].p;
    fs_spmv(g, F.Dx, field, F.t0.p);
    fs_spmv(g, F.Dy, field, F.t1.p);
    fs_spmv(g, F.Lap, field, F.t2.p);
    k_fs_hat<<<nb, kBlock, 0, g.stream>>>(N, field, F.vec[0].p, F.vec[1].p, F.t0.p, F.t1.p, F.t2.p, F.dt, F.mu / F.rho, F.vec[4 + w].p);
  }
  MMG_CUDA(cudaGetLastError());
}

void fs_set_ppe_source(Grid& g) {
  Grid::FracStep& F = *g.fs;
  MMG_REQUIRE(F.have_ops, MMG_ERR_STATE, "the fractional-step operators have not been built or uploaded");
  const int N = g.n, nb = (N + kBlock - 1) / kBlock;
  TimedScope ts(g, MMG_T_OTHER, 2 * F.Dx.matrix_bytes() + (int64_t)N * 8 * 6, 4);
  fs_spmv(g, F.Dx, F.vec[4].p, F.t0.p);
  fs_spmv(g, F.Dy, F.vec[5].p, F.t1.p);
  k_fs_ppe_interior<<<nb, kBlock, 0, g.stream>>>(N, F.t0.p, F.t1.p, F.rho / F.dt, g.b.p);
  const int m = (int)F.bnd_pts.n;
  if (m) k_fs_ppe_boundary<<<(m + kBlock - 1) / kBlock, kBlock, 0, g.stream>>>(m, F.bnd_pts.p, F.vec[0].p, F.vec[1].p, F.vec[4].p, F.vec[5].p, F.nx.p, F.ny.p,
                                                                                -F.rho / F.dt, g.b.p);
  MMG_CUDA(cudaGetLastError());
}

void fs_correct(Grid& g, int component) {
  Grid::FracStep& F = *g.fs;
  MMG_REQUIRE(F.have_ops, MMG_ERR_STATE, "the fractional-step operators have not been built or uploaded");
  const int N = g.n, nb = (N + kBlock - 1) / kBlock;
  TimedScope ts(g, MMG_T_OTHER, 2 * F.Dx.matrix_bytes() + (int64_t)N * 8 * 6, 4);
  if (component == MMG_FS_BOTH || component == MMG_FS_U) {
    fs_spmv(g, F.Dx, g.x.p, F.t0.p);      // values_->head(N): the operators have N columns
    k_fs_correct<<<nb, kBlock, 0, g.stream>>>(N, F.vec[4].p, F.t0.p, F.dt / F.rho, F.vec[0].p);
  }
  if (component == MMG_FS_BOTH || component == MMG_FS_V) {
    fs_spmv(g, F.Dy, g.x.p, F.t0.p);
    k_fs_correct<<<nb, kBlock, 0, g.stream>>>(N, F.vec[5].p, F.t0.p, F.dt / F.rho, F.vec[1].p);
  }
  MMG_CUDA(cudaGetLastError());
}

double fs_residual(Grid& g) {   // (*u - *u_hat).lpNorm<1>() / N   (:152-154)
  Grid::FracStep& F = *g.fs;
  const int blocks = 256;
  ensure_partials(g, blocks);
  k_fs_absdiff<<<blocks, kBlock, 0, g.stream>>>(g.n, F.vec[0].p, F.vec[4].p, g.partials.p);
  MMG_CUDA(cudaGetLastError());
  std::vector<double> h(blocks);
  g.partials.download(h.data(), blocks, g.stream);
  double s = 0;
  for (double t : h) s += t;
  return s / g.n;
}

// ================================================================================================
// Multi-GPU: row-block partition of the large levels, halo exchange of contiguous ranges
// ================================================================================================
namespace {
__global__ void __launch_bounds__(kBlock) k_col_range(HybView A, int row0, int nrows, int* out2) {
  int mn = 0x7fffffff, mx = -1;
  for (int r = row0 + blockIdx.x * blockDim.x + threadIdx.x; r < row0 + nrows; r += gridDim.x * blockDim.x) {
    const int len = A.len[r];
    const int* __restrict__ c = row_col(A, r);
    const int m = len < A.W ? len : A.W;
    for (int k = 0; k < m; k++) { mn = min(mn, c[k]); mx = max(mx, c[k]); }
  }
  mn = __reduce_min_sync(0xffffffffu, mn);
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0) { atomicMin(out2, mn); atomicMax(out2 + 1, mx); }
}
__global__ void k_ratio(const double* sums, double* ratio) { *ratio = sums[0] / sums[1]; }
__global__ void __launch_bounds__(kBlock) k_sum_partials(const double* __restrict__ partial, int n, double* sums) {
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { a += partial[2 * i]; b += partial[2 * i + 1]; }
  block_sum2(a, b);
  if (threadIdx.x == 0) { sums[0] = a; sums[1] = b; }
}

// [min col, max col + 1) over rows [row0, row0+nrows) of M, for every rank's block
std::vector<std::pair<int, int>> column_needs(Grid& ctx, const HybMatrix& M, const std::vector<int>& row_bounds) {
  MMG_REQUIRE(M.n_ovf == 0 && M.reg_row < 0, MMG_ERR_STATE, "multi-GPU partition supports operators without spill rows (Dirichlet grids)");
  const int world = (int)row_bounds.size() - 1;
  DevBuf<int> out;
  out.alloc(2 * world);
  std::vector<int> init(2 * world);
  for (int r = 0; r < world; r++) { init[2 * r] = 0x7fffffff; init[2 * r + 1] = -1; }
  out.upload(init, ctx.stream);
  for (int r = 0; r < world; r++) {
    const int n = row_bounds[r + 1] - row_bounds[r];
    if (n > 0) k_col_range<<<std::min(1024, (n + kBlock - 1) / kBlock), kBlock, 0, ctx.stream>>>(M.view(), row_bounds[r], n, out.p + 2 * r);
  }
  MMG_CUDA(cudaGetLastError());
  std::vector<int> h = out.to_host(ctx.stream);
  std::vector<std::pair<int, int>> need(world);
  for (int r = 0; r < world; r++) need[r] = h[2 * r + 1] < 0 ? std::make_pair(0, 0) : std::make_pair(h[2 * r], h[2 * r + 1] + 1);
  return need;
}
}  // namespace

void dist_setup(Solver& s) {
  const int L = (int)s.grids.size(), W = s.world;
  peer_teardown(s);                      // a repeated set-up (new threshold / communicator) starts from scratch
  for (Grid* gp : s.grids) { gp->mc_row0 = 0; gp->mc_row1 = -1; gp->mc_packed = false; gp->mc_chunks.release(); }
  s.dist.clear();
  s.dist.resize(L);
  if (s.sums.n < 2) s.sums.alloc(2);
  for (int l = 0; l < L; l++) {
    Grid& g = *s.grids[l];
    LevelDist& D = s.dist[l];
    MMG_REQUIRE(!g.neumann, MMG_ERR_STATE, "multi-GPU partition is implemented for grids without Neumann boundaries");
    D.partitioned = W > 1 && g.n >= s.part_threshold;
    D.bounds.resize(W + 1);
    partition_bounds(g.n, W, D.bounds.data());
    if (!D.partitioned) continue;
    plan_build(D.x_plan, s.rank, W, column_needs(g, g.Lap, D.bounds), D.bounds);
    if (!g.have_colours) build_colouring(g);
    D.colour_sub.resize(g.n_colours);
    for (int c = 0; c < g.n_colours; c++) {
      const int* b = g.colour_rows_host.data() + g.colour_ptr[c];
      const int* e = g.colour_rows_host.data() + g.colour_ptr[c + 1];
      const int* lo = std::lower_bound(b, e, D.bounds[s.rank]);
      const int* hi = std::lower_bound(b, e, D.bounds[s.rank + 1]);
      D.colour_sub[c] = {(int)(lo - g.colour_rows_host.data()), (int)(hi - lo)};
    }
  }
  for (int l = 1; l < L; l++) {
    LevelDist& D = s.dist[l];
    if (!D.partitioned) continue;
    Grid& fine = *s.grids[l];
    Grid& coarse = *s.grids[l - 1];
    // restriction: its output rows (coarse rows) are split like the coarse level's vectors; it reads this level's residual
    D.r_bounds = s.dist[l - 1].bounds;
    plan_build(D.r_plan, s.rank, W, column_needs(fine, *s.restrict_[l], D.r_bounds), D.bounds);
    // prolongation: output rows = my rows of this level; it reads the coarser level's values_
    if (s.dist[l - 1].partitioned) plan_build(D.p_plan, s.rank, W, column_needs(fine, *s.prolong_[l - 1], D.bounds), s.dist[l - 1].bounds);
    (void)coarse;
  }
  if (W > 1 && env_int("MMG_DIST_PEER", 1)) peer_setup(s);
  // the one-time packed copies are built here, not lazily inside the first sweep: every rank then enters its first peer-memory
  // sweep at the same time (a rank still packing would let its neighbours' watchdogs run)
  for (int l = 0; l < L; l++) {
    LevelDist& D = s.dist[l];
    Grid& g = *s.grids[l];
    if (!D.partitioned || !D.peer_ready || g.Lap.n_ovf != 0 || !g.Lap.diag_first) continue;
    g.mc_row0 = D.bounds[s.rank]; g.mc_row1 = D.bounds[s.rank + 1];
    g.mc_packed = false;
    ensure_mc_pack(g);
  }
  s.dist_ready = true;
}

void dist_residual(Solver& s, int level) {
  Grid& g = *s.grids[level];
  const LevelDist& D = s.dist[level];
  if (!D.partitioned) { op_residual(g, g.r.p); return; }
  const int lo = D.bounds[s.rank], n = D.bounds[s.rank + 1] - lo;
  TimedScope ts(g, MMG_T_RESIDUAL, (g.Lap.matrix_bytes() + (int64_t)g.A * 24) / s.world, 1);
  launch_spmv(g.Lap, g.x.p, g.b.p, g.r.p, g.rowflag.p, OP_RESID, 0, 0, nullptr, nullptr, g.device, g.stream, false, lo, n);
}

void dist_residual_norm(Solver& s, int level, double* ratio_dev) {
  Grid& g = *s.grids[level];
  const LevelDist& D = s.dist[level];
  if (!D.partitioned) { op_residual_norm(g, ratio_dev); return; }
  const int lo = D.bounds[s.rank], n = D.bounds[s.rank + 1] - lo;
  ensure_partials(g, (size_t)2 * (sm_count_of(g.device) * 32 + 8) + kRegBlocks);
  int nb = 0;
  {
    TimedScope ts(g, MMG_T_RESIDUAL, (g.Lap.matrix_bytes() + (int64_t)g.A * 24) / s.world, 3);
    launch_spmv(g.Lap, g.x.p, g.b.p, nullptr, g.rowflag.p, OP_RESID, 0, 0, g.partials.p, &nb, g.device, g.stream, false, lo, n);
    k_sum_partials<<<1, kBlock, 0, g.stream>>>(g.partials.p, nb, s.sums.p);
    MMG_CUDA(cudaGetLastError());
  }
  allreduce_sum(s, s.sums.p, 2);
  k_ratio<<<1, 1, 0, g.stream>>>(s.sums.p, ratio_dev);
  MMG_CUDA(cudaGetLastError());
}

// ---- peer-memory smoother -------------------------------------------------------------------------
// both sets: versions 1..iters of every swept row = sentinel (run once, before any rank can store into them)
__global__ void k_peer_init_all(const unsigned char* __restrict__ rowflag, double* xs, size_t stride, int versions, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total || rowflag[i] != 0) return;
  for (int set = 0; set < 2; set++)
    for (int v = 1; v < versions; v++) xs[((size_t)set * versions + v) * stride + i] = __longlong_as_double((long long)kSentinelBits);
}
// start of a smoothing call on set `cur`: version 0 = values_; rows the sweep skips are constant in every version;
// swept rows of versions >= 1 already hold the sentinel (or a neighbour's early store) -- they were reset one call ago,
// which is what this launch does for the other set: nobody can touch that set before my boundary rows of THIS call
// have reached the neighbours (they need them to finish their call), so the reset cannot overwrite live data.
__global__ void k_peer_call_init(const unsigned char* __restrict__ rowflag, const double* __restrict__ x, double* cur, double* nxt, size_t stride,
                                 int iters, int versions, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const double xi = x[i];
  const bool skipped = rowflag[i] != 0;
  cur[i] = xi;
  for (int v = 1; v <= iters; v++)
    if (skipped) cur[(size_t)v * stride + i] = xi;
  // every version of the other set, not only 1..iters: props.iters may grow again (up to peer_iters) before the next dist_setup
  for (int v = 1; v <= versions; v++) nxt[(size_t)v * stride + i] = skipped ? xi : __longlong_as_double((long long)kSentinelBits);
}
// End of a smoothing call: my rows only waited for the halo values THEY read, so the last version of a halo row of a
// higher colour may still be in flight from the neighbour.  Wait for each (the sentinel is the "not yet" flag) and
// copy it into values_.  The neighbour's sweep never waits for this kernel, so it cannot deadlock.
__global__ void k_peer_collect(const double* last, double* x, int offset, int count, int* abort_flag, long long timeout_cycles) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const long long t_start = clock64();
  double v = ld_relaxed_sys(last + offset + i);
  while (is_sentinel(v)) {
    if (*(volatile int*)abort_flag || clock64() - t_start > timeout_cycles) { atomicExch(abort_flag, 1); v = 0.0; break; }
    v = ld_relaxed_sys(last + offset + i);
  }
  x[offset + i] = v;
}
void peer_init_sets(Grid& g, LevelDist& D) {
  k_peer_init_all<<<(g.A + kBlock - 1) / kBlock, kBlock, 0, g.stream>>>(g.rowflag.p, D.peer_xs.p, D.peer_stride, D.peer_iters + 1, g.A);
  MMG_CUDA(cudaGetLastError());
}

// All sweeps of one smoothing call on a partitioned level in ONE launch per rank, no collective inside: the barrier-free
// sweep over this rank's rows, whose stores next to a cut also go into the neighbour's version vectors over NVLink.
static bool dist_sor_peer(Solver& s, Grid& g, LevelDist& D) {
  const HybMatrix& L = g.Lap;
  const int iters = g.props.iters;
  if (!D.peer_ready || iters > D.peer_iters || iters < 1 || L.n_ovf != 0 || !L.diag_first) return false;
  if (g.mc_row1 < 0 || !g.mc_packed) {
    g.mc_row0 = D.bounds[s.rank]; g.mc_row1 = D.bounds[s.rank + 1];
    g.mc_packed = false;
    ensure_mc_pack(g);
  }
  const size_t stride = D.peer_stride, set_elems = (size_t)(D.peer_iters + 1) * stride;
  double* cur = D.peer_xs.p + (size_t)D.peer_parity * set_elems;
  double* nxt = D.peer_xs.p + (size_t)(D.peer_parity ^ 1) * set_elems;
  TimedScope ts(g, MMG_T_SOR, (L.matrix_bytes() + (int64_t)g.A * 28) * iters / s.world, 3);
  k_peer_call_init<<<(g.A + kBlock - 1) / kBlock, kBlock, 0, g.stream>>>(g.rowflag.p, g.x.p, cur, nxt, stride, iters, D.peer_iters, g.A);
  MMG_CUDA(cudaGetLastError());
  bool tma_done = false;
  if (g.mc_colour_ptr.back() >= env_int("MMG_MC_TMAFLOW_MIN_ROWS", 2000) && env_int("MMG_MC_TMAFLOW", 1)) {
    PeerSends peers{};
    peers.n = D.n_sends;
    for (int k = 0; k < D.n_sends; k++) {
      peers.lo[k] = D.send_lo[k]; peers.hi[k] = D.send_hi[k];
      peers.base[k] = D.send_base[k] + (size_t)D.peer_parity * set_elems;
    }
    tma_done = stream_sor_mc_flow(g, cur, stride, &peers);
  }
  const bool ok = tma_done || dispatch_lpr_iter(L.W, [&](auto Lc, auto I) {
    constexpr int LPR = decltype(Lc)::value, ITER = decltype(I)::value;
    const int rows_used = env_int("MMG_MC_FLOW_ROWS", 1) == 1 ? 1 : 2;
    void* kern = rows_used == 1 ? (void*)k_sor_mc_flow<LPR, ITER, 1, true> : (void*)k_sor_mc_flow<LPR, ITER, 2, true>;
    int blocks_per_sm = 0;
    MMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kBlock, 0));
    const int sms = sm_count_of(g.device);
    int maxcount = 0;
    for (int c = 0; c < g.n_colours; c++) maxcount = std::max(maxcount, g.mc_colour_ptr[c + 1] - g.mc_colour_ptr[c]);
    int blocks = std::min(blocks_per_sm * sms, grid_for2(std::max(1, maxcount), LPR, sms, rows_used));
    const unsigned char* chunks = g.mc_chunks.p;
    size_t cb = L.chunk_bytes, st = stride;
    int W = L.W;
    const int* cp = g.mc_colour_ptr_dev.p;
    int nc = g.n_colours, itn = iters;
    const double* b = g.b.p;
    double* xs = cur;
    double omega = g.props.omega;
    int* abortp = g.abort_flag.p;
    long long timeout = 6000000000ll;
    PeerSends peers{};
    peers.n = D.n_sends;
    for (int k = 0; k < D.n_sends; k++) {
      peers.lo[k] = D.send_lo[k]; peers.hi[k] = D.send_hi[k];
      peers.base[k] = D.send_base[k] + (size_t)D.peer_parity * set_elems;
    }
    void* args[] = {&chunks, &cb, &W, &cp, &nc, &itn, &b, &xs, &st, &omega, &abortp, &timeout, &peers};
    MMG_CUDA(cudaLaunchCooperativeKernel(kern, dim3(blocks), dim3(kBlock), args, 0, g.stream));
  }, (g.A >= 200000 && L.W > 16 && L.W <= 40) ? 8 : 0);
  MMG_REQUIRE(ok, MMG_ERR_STATE, "stencil width outside the fast kernels' table");
  // values_ <- last version: my block plus the halo ranges the neighbours mirrored into my copy
  const double* last = cur + (size_t)iters * stride;
  const int lo = D.bounds[s.rank], n = D.bounds[s.rank + 1] - lo;
  MMG_CUDA(cudaMemcpyAsync(g.x.p + lo, last + lo, sizeof(double) * n, cudaMemcpyDeviceToDevice, g.stream));
  for (const ExchangePlan::Msg& m : D.x_plan.recvs) {
    if (m.count <= 0) continue;
    k_peer_collect<<<(m.count + kBlock - 1) / kBlock, kBlock, 0, g.stream>>>(last, g.x.p, m.offset, m.count, g.abort_flag.p, 6000000000ll);
    MMG_CUDA(cudaGetLastError());
  }
  D.peer_parity ^= 1;
  return true;
}

void dist_sor(Solver& s, int level) {
  Grid& g = *s.grids[level];
  LevelDist& D = s.dist[level];
  if (!D.partitioned) { op_sor(g, s.smoother); return; }
  MMG_REQUIRE(s.smoother == MMG_SMOOTHER_MULTICOLOUR && !g.exact, MMG_ERR_STATE,
              "a partitioned level needs the multicolour smoother with fast arithmetic: the lexicographic sweep is a pipeline across ranks, not a partition");
  if (dist_sor_peer(s, g, D)) return;
  const HybMatrix& L = g.Lap;
  const int sms = sm_count_of(g.device);
  for (int it = 0; it < g.props.iters; it++) {
    TimedScope ts(g, MMG_T_SOR, (L.matrix_bytes() + (int64_t)g.A * 28) / s.world, g.n_colours);
    for (int c = 0; c < g.n_colours; c++) {
      const int first = D.colour_sub[c].first, count = D.colour_sub[c].second;
      if (count > 0) {
        const bool ok = dispatch_lpr_iter(L.W, [&](auto Lc, auto I) {
          constexpr int LPR = decltype(Lc)::value, ITER = decltype(I)::value;
          k_sor_mc2<LPR, ITER, kMcRows><<<grid_for2(count, LPR, sms, kMcRows), kBlock, 0, g.stream>>>(L.view(), g.colour_rows.p + first, count, g.b.p,
                                                                                                           g.x.p, g.props.omega, g.A);
        });
        MMG_REQUIRE(ok, MMG_ERR_STATE, "stencil width outside the fast kernels' table");
      }
      plan_execute(s, D.x_plan, g.x.p);     // neighbours need this colour's new values before their next phase
    }
    MMG_CUDA(cudaGetLastError());
  }
}

void dist_restrict(Solver& s, int level) {
  Grid& fine = *s.grids[level];
  Grid& coarse = *s.grids[level - 1];
  const LevelDist& D = s.dist[level];
  if (!D.partitioned) { op_restrict(fine, coarse, *s.restrict_[level], fine.r.p); return; }
  const HybMatrix& R = *s.restrict_[level];
  plan_execute(s, D.r_plan, fine.r.p);
  const int lo = D.r_bounds[s.rank], n = D.r_bounds[s.rank + 1] - lo;
  {
    TimedScope ts(fine, MMG_T_RESTRICT, (R.matrix_bytes() + (int64_t)coarse.n * 8 + (int64_t)fine.n * 8) / s.world, 1);
    if (n > 0) launch_spmv(R, fine.r.p, nullptr, coarse.b.p, coarse.rowflag.p, OP_RESTRICT, 1, 0, nullptr, nullptr, fine.device, fine.stream, false, lo, n);
  }
  if (!s.dist[level - 1].partitioned) allgather_blocks(s, coarse.b.p, D.r_bounds);   // the coarser level is replicated: everyone needs all of its source
}

void dist_prolong_correct(Solver& s, int level) {
  Grid& fine = *s.grids[level];
  Grid& coarse = *s.grids[level - 1];
  const LevelDist& D = s.dist[level];
  if (!D.partitioned) { op_prolong_correct(fine, coarse, *s.prolong_[level - 1]); return; }
  const HybMatrix& P = *s.prolong_[level - 1];
  if (s.dist[level - 1].partitioned) plan_execute(s, D.p_plan, coarse.x.p);
  const int lo = D.bounds[s.rank], n = D.bounds[s.rank + 1] - lo;
  {
    TimedScope ts(fine, MMG_T_PROLONG, (P.matrix_bytes() + (int64_t)fine.n * 16 + (int64_t)coarse.n * 8) / s.world, 1);
    launch_spmv(P, coarse.x.p, nullptr, fine.x.p, fine.rowflag.p, OP_PROLONG, 1, 0, nullptr, nullptr, fine.device, fine.stream, false, lo, n);
  }
  plan_execute(s, D.x_plan, fine.x.p);      // corrected values next to the cut
}

// ---- schedules (integer artefacts) ----------------------------------------------------------------
void build_colouring(Grid& g) {
  // First-fit in ascending row order on the structurally symmetrised graph of the rows the sweep visits;
  // the regularisation row takes the last colour (colouring contract: DESIGN.md §5).
  HostCsr A;
  hyb_to_csr(g.Lap, A, g.stream);
  const int R = A.rows;
  const int reg = g.Lap.reg_row;
  auto swept = [&](int i) { return i == reg || g.bcflags[i] == 0; };
  std::vector<int> tcnt(R + 1, 0);
  for (int i = 0; i < R; i++) {
    if (i == reg || !swept(i)) continue;
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++) { const int j = A.idx[k]; if (j != reg && j != i && swept(j)) tcnt[j + 1]++; }
  }
  for (int i = 0; i < R; i++) tcnt[i + 1] += tcnt[i];
  std::vector<int> tidx(tcnt[R]), pos(tcnt.begin(), tcnt.end() - 1);
  for (int i = 0; i < R; i++) {
    if (i == reg || !swept(i)) continue;
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++) { const int j = A.idx[k]; if (j != reg && j != i && swept(j)) tidx[pos[j]++] = i; }
  }
  std::vector<int> colour(R, -1), mark;
  int ncol = 0;
  for (int i = 0; i < R; i++) {
    if (i == reg || !swept(i)) continue;
    auto touch = [&](int j) {
      if (j < i && j != reg && colour[j] >= 0) {
        if ((int)mark.size() <= colour[j]) mark.resize(colour[j] + 1, -1);
        mark[colour[j]] = i;
      }
    };
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++) touch(A.idx[k]);
    for (int k = tcnt[i]; k < tcnt[i + 1]; k++) touch(tidx[k]);
    int c = 0;
    while (c < (int)mark.size() && mark[c] == i) c++;
    colour[i] = c;
    ncol = std::max(ncol, c + 1);
  }
  if (reg >= 0) { colour[reg] = ncol; ncol++; }
  g.colour_host = colour;
  g.n_colours = ncol;
  g.colour_ptr.assign(ncol + 1, 0);
  for (int i = 0; i < R; i++) if (colour[i] >= 0) g.colour_ptr[colour[i] + 1]++;
  for (int c = 0; c < ncol; c++) g.colour_ptr[c + 1] += g.colour_ptr[c];
  std::vector<int> rows(g.colour_ptr[ncol]), cp(g.colour_ptr.begin(), g.colour_ptr.end() - 1);
  for (int i = 0; i < R; i++) if (colour[i] >= 0) rows[cp[colour[i]]++] = i;
  g.colour_rows.upload(rows, g.stream);
  g.colour_rows_host = rows;
  g.colour_ptr_dev.upload(g.colour_ptr, g.stream);
  MMG_CUDA(cudaStreamSynchronize(g.stream));
  g.have_colours = true;
  g.mc_packed = false;
  g.mc_chunks.release();
}

void ensure_mc_pack(Grid& g) {
  if (g.mc_packed) return;
  MMG_REQUIRE(g.have_colours, MMG_ERR_STATE, "ensure_mc_pack: colouring missing");
  // rows the copy covers: everything, or this rank's block of a partitioned level (mc_row0/mc_row1)
  const int r0 = g.mc_row1 < 0 ? 0 : g.mc_row0, r1 = g.mc_row1 < 0 ? g.A : g.mc_row1;
  std::vector<int> rows;
  // interior colours; the regularisation row (last colour of a Neumann-type grid) has no chunk: the TMA-fed sweep handles it as
  // a reduction inside the extra phase that also evaluates the Neumann boundary rows (last range of the packed copy)
  const int n_int = g.Lap.reg_row >= 0 ? g.n_colours - 1 : g.n_colours;
  g.mc_colour_ptr.assign(n_int + 1, 0);
  for (int c = 0; c < n_int; c++) {
    for (int k = g.colour_ptr[c]; k < g.colour_ptr[c + 1]; k++) { const int r = g.colour_rows_host[k]; if (r >= r0 && r < r1) rows.push_back(r); }
    g.mc_colour_ptr[c + 1] = (int)rows.size();
  }
  const int total_int = (int)rows.size();
  g.mc_bnd_phase = -1;
  if (g.neumann) {
    MMG_REQUIRE(g.mc_row1 < 0, MMG_ERR_STATE, "partitioned levels with Neumann boundaries are not supported");
    for (const Boundary& bd : g.boundaries)
      if (bd.type == MMG_BC_NEUMANN) rows.insert(rows.end(), bd.pts.begin(), bd.pts.end());
    g.mc_colour_ptr.push_back((int)rows.size());
    g.mc_bnd_phase = n_int;
  }
  g.mc_colour_ptr_dev.upload(g.mc_colour_ptr, g.stream);
  const int total = (int)rows.size();
  if (env_int("MMG_MC_ORDER", 1) == 1 && total > 0) {
    double x0 = g.hx[0], x1 = g.hx[0], y0 = g.hy[0], y1 = g.hy[0];
    for (int i = 0; i < g.n; i++) { x0 = std::min(x0, g.hx[i]); x1 = std::max(x1, g.hx[i]); y0 = std::min(y0, g.hy[i]); y1 = std::max(y1, g.hy[i]); }
    const double sx = x1 > x0 ? 65535.0 / (x1 - x0) : 0.0, sy = y1 > y0 ? 65535.0 / (y1 - y0) : 0.0;
    auto spread = [](uint32_t v) {
      v &= 0xFFFFu;
      v = (v | (v << 8)) & 0x00FF00FFu; v = (v | (v << 4)) & 0x0F0F0F0Fu; v = (v | (v << 2)) & 0x33333333u; v = (v | (v << 1)) & 0x55555555u;
      return v;
    };
    std::vector<std::pair<uint32_t, int>> keyed;
    for (int c = 0; c < n_int; c++) {
      const int a = g.mc_colour_ptr[c], e = g.mc_colour_ptr[c + 1];
      keyed.clear();
      for (int k = a; k < e; k++) {
        const int r = rows[k];
        const uint32_t qx = (uint32_t)((g.hx[r] - x0) * sx), qy = (uint32_t)((g.hy[r] - y0) * sy);
        keyed.push_back({spread(qx) | (spread(qy) << 1), r});
      }
      std::sort(keyed.begin(), keyed.end());
      for (int k = a; k < e; k++) rows[k] = keyed[k - a].second;
    }
  }
  DevBuf<int> drows;
  drows.upload(rows, g.stream);
  g.mc_chunks.alloc((size_t)total * g.Lap.chunk_bytes);
  if (total > 0) {
    const long long threads = (long long)total * 32;
    k_pack_chunks<<<(unsigned)((threads + kBlock - 1) / kBlock), kBlock, 0, g.stream>>>(g.Lap.chunks.p, g.Lap.chunk_bytes, drows.p, total, g.mc_chunks.p);
    MMG_CUDA(cudaGetLastError());
    DevBuf<int> dcol;
    dcol.upload(g.colour_host, g.stream);
    k_mark_lower_colour<<<(total_int + kBlock - 1) / kBlock, kBlock, 0, g.stream>>>(g.mc_chunks.p, g.Lap.chunk_bytes, g.Lap.W, total_int, dcol.p, n_int);
    MMG_CUDA(cudaGetLastError());
    if (g.Lap.n_ovf) {
      k_mark_overflow<<<(total + kBlock - 1) / kBlock, kBlock, 0, g.stream>>>(g.mc_chunks.p, g.Lap.chunk_bytes, g.Lap.W, total, g.Lap.len.p);
      MMG_CUDA(cudaGetLastError());
    }
    MMG_CUDA(cudaStreamSynchronize(g.stream));
  }
  MMG_CUDA(cudaStreamSynchronize(g.stream));
  g.mc_packed = true;
}

void compute_lex_levels(Grid& g, std::vector<int>& level, int& n_levels) {
  HostCsr A;
  hyb_to_csr(g.Lap, A, g.stream);
  const int R = A.rows, reg = g.Lap.reg_row;
  level.assign(R, -1);
  n_levels = 0;
  for (int i = 0; i < R; i++) {
    const bool swept = i == reg || g.bcflags[i] == 0;
    if (!swept) continue;
    int l = 0;
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++) { const int j = A.idx[k]; if (j < i && level[j] >= 0) l = std::max(l, level[j] + 1); }
    level[i] = l;
    n_levels = std::max(n_levels, l + 1);
  }
}

}  // namespace mmg
