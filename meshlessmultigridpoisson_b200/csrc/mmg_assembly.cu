// RBF-FD assembly on the device (sm_100a): exact cell-grid kNN with the reference's
// (distance, index) ordering, batched dense full-pivot LU of the PHS+polynomial saddle systems
// (one CTA per stencil, matrix resident in shared memory), and direct emission of row chunks.
//
// Arithmetic contract: the same sequence of IEEE operations as the oracle's restatement of
// grid.cpp:263-424,687-712 and of Eigen's FullPivLU (no FMA contraction — the library is built with
// --fmad=false; fp64 '/' and sqrt are IEEE in CUDA).  The only places that cannot be bit-identical
// to a CPU libm are pow(r,3), pow(x,int) and pow(D,-1/2): they are evaluated to <0.5000001 ulp with
// double-double arithmetic, which agrees with glibc's pow except where glibc itself mis-rounds.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <queue>

#include "mmg_internal.hpp"

namespace mmg {

namespace {

struct CellGrid {
  double x0, y0, cs;
  int nx, ny;
  const int* start;  // nx*ny+1
  const int* ids;    // n
};

struct AsmState {
  DevBuf<int> cell_start, cell_ids, dflags;
  double x0 = 0, y0 = 0, cs = 1;
  int nx = 0, ny = 0;
  bool cells_valid = false;
  // deriv_normal_coeffs_ (grid.cpp:520-548), host side: one entry per Neumann node in boundary-list order
  std::vector<int> dn_point;
  std::vector<double> dn_w;   // stencil entries per node
  std::vector<int> dn_nb;
  DevBuf<int> err;
  CellGrid view() const { return CellGrid{x0, y0, cs, nx, ny, cell_start.p, cell_ids.p}; }
};

AsmState& state(Grid& g) {
  if (!g.asm_state) g.asm_state = new AsmState();
  return *static_cast<AsmState*>(g.asm_state);
}

void build_cells(Grid& g) {
  AsmState& st = state(g);
  const int N = g.n;
  double minX = g.hx[0], maxX = minX, minY = g.hy[0], maxY = minY;
  for (int i = 0; i < N; i++) {
    minX = std::min(minX, g.hx[i]); maxX = std::max(maxX, g.hx[i]);
    minY = std::min(minY, g.hy[i]); maxY = std::max(maxY, g.hy[i]);
  }
  const double area = std::max((maxX - minX) * (maxY - minY), 1e-300);
  st.x0 = minX; st.y0 = minY;
  st.cs = 2.0 * std::sqrt(area / N);
  st.nx = (int)std::floor((maxX - minX) / st.cs) + 1;
  st.ny = (int)std::floor((maxY - minY) / st.cs) + 1;
  std::vector<int> start((size_t)st.nx * st.ny + 1, 0), cell(N), ids(N);
  for (int i = 0; i < N; i++) {
    const int cx = std::min(st.nx - 1, std::max(0, (int)std::floor((g.hx[i] - st.x0) / st.cs)));
    const int cy = std::min(st.ny - 1, std::max(0, (int)std::floor((g.hy[i] - st.y0) / st.cs)));
    cell[i] = cy * st.nx + cx;
    start[cell[i] + 1]++;
  }
  for (size_t c = 0; c + 1 < start.size(); c++) start[c + 1] += start[c];
  std::vector<int> pos(start.begin(), start.end() - 1);
  for (int i = 0; i < N; i++) ids[pos[cell[i]]++] = i;
  st.cell_start.upload(start, g.stream);
  st.cell_ids.upload(ids, g.stream);
  st.dflags.upload(g.bcflags, g.stream);
  if (!st.err.p) { st.err.alloc(1); st.err.zero(g.stream); }
  g.sync();
  st.cells_valid = true;
}

// ------------------------------------------------------------------------------------------------
// exact kNN: one warp per query, candidates from a growing square of cells, bitonic sort on the
// reference's key (sqrt(dx^2+dy^2), index)  (grid.cpp:216-260; tie-break = std::pair ordering)
// ------------------------------------------------------------------------------------------------
constexpr int KNN_WARPS = 4;
constexpr int KNN_CAP = 1024;
constexpr unsigned long long KEY_EXCLUDED = 0x7FF0000000000000ull;  // +inf: sorts after every real distance
constexpr unsigned long long KEY_PAD = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ bool key_less(unsigned long long ka, int ia, unsigned long long kb, int ib) { return ka < kb || (ka == kb && ia < ib); }

__global__ void __launch_bounds__(KNN_WARPS * 32) k_knn(CellGrid cg, const double* __restrict__ px, const double* __restrict__ py,
                                                        const int* __restrict__ bcflag, int m, const double* __restrict__ qx,
                                                        const double* __restrict__ qy, const int* __restrict__ qflag, int neumann, int k,
                                                        int* __restrict__ out, int* err) {
  __shared__ unsigned long long skey[KNN_WARPS][KNN_CAP];
  __shared__ int sidx[KNN_WARPS][KNN_CAP];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * KNN_WARPS + w;
  if (q >= m) return;
  unsigned long long* key = skey[w];
  int* idx = sidx[w];
  const double x = qx[q], y = qy[q];
  const bool excl = neumann && qflag && qflag[q] != 0;
  const int cx = min(cg.nx - 1, max(0, (int)floor((x - cg.x0) / cg.cs)));
  const int cy = min(cg.ny - 1, max(0, (int)floor((y - cg.y0) / cg.cs)));
  int rc = 2;
  while (true) {
    const int xlo = max(cx - rc, 0), xhi = min(cx + rc, cg.nx - 1), ylo = max(cy - rc, 0), yhi = min(cy + rc, cg.ny - 1);
    int count = 0;
    for (int yy = ylo; yy <= yhi; yy++) {
      const int p0 = cg.start[yy * cg.nx + xlo], p1 = cg.start[yy * cg.nx + xhi + 1];
      for (int p = p0 + lane; p - lane < p1; p += 32) {
        const bool have = p < p1;
        int i = 0;
        unsigned long long kk = 0;
        if (have) {
          i = cg.ids[p];
          const double dx = x - px[i], dy = y - py[i];
          const double d = sqrt(dx * dx + dy * dy);   // fmad off: two roundings + add + sqrt, as the reference
          kk = (unsigned long long)__double_as_longlong(d);
        }
        const unsigned ball = __ballot_sync(0xffffffffu, have);
        const int slot = count + __popc(ball & ((1u << lane) - 1u));
        if (have && slot < KNN_CAP) { key[slot] = kk; idx[slot] = i; }
        count += __popc(ball);
      }
    }
    if (count > KNN_CAP) { if (lane == 0) atomicExch(err, 1); return; }
    __syncwarp();
    if (excl) {  // samePoint = last index at distance exactly 0; every other flagged node is excluded (grid.cpp:236,244)
      int same = -1;
      for (int j = lane; j < count; j += 32) if (key[j] == 0ull) same = max(same, idx[j]);
      for (int o = 16; o > 0; o >>= 1) same = max(same, __shfl_xor_sync(0xffffffffu, same, o));
      for (int j = lane; j < count; j += 32) if (idx[j] != same && bcflag[idx[j]] != 0) key[j] = KEY_EXCLUDED;
    }
    int P = 32;
    while (P < count) P <<= 1;
    for (int j = count + lane; j < P; j += 32) { key[j] = KEY_PAD; idx[j] = 0x7fffffff; }
    __syncwarp();
    for (int size = 2; size <= P; size <<= 1)
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = lane; t < (P >> 1); t += 32) {
          const int lo = ((t / stride) * stride << 1) + (t % stride), hi = lo + stride;
          const bool up = ((lo & size) == 0);
          const unsigned long long ka = key[lo], kb = key[hi];
          const int ia = idx[lo], ib = idx[hi];
          const bool sw = up ? key_less(kb, ib, ka, ia) : key_less(ka, ia, kb, ib);
          if (sw) { key[lo] = kb; key[hi] = ka; idx[lo] = ib; idx[hi] = ia; }
        }
        __syncwarp();
      }
    // every point outside the visited square is farther than `safe`
    const double inf = __longlong_as_double(0x7FF0000000000000ll);
    const double sx0 = (cx - rc <= 0) ? inf : x - (cg.x0 + (cx - rc) * cg.cs);
    const double sx1 = (cx + rc >= cg.nx - 1) ? inf : (cg.x0 + (cx + rc + 1) * cg.cs) - x;
    const double sy0 = (cy - rc <= 0) ? inf : y - (cg.y0 + (cy - rc) * cg.cs);
    const double sy1 = (cy + rc >= cg.ny - 1) ? inf : (cg.y0 + (cy + rc + 1) * cg.cs) - y;
    const double safe = fmin(fmin(sx0, sx1), fmin(sy0, sy1));
    const bool enough = count >= k && key[k - 1] < KEY_EXCLUDED;
    const bool done = enough && (__longlong_as_double((long long)key[k - 1]) < safe * (1 - 1e-12) || safe == inf);
    if (done) break;
    if (safe == inf) { if (lane == 0) atomicExch(err, 2); return; }  // fewer admissible points than k
    rc++;
    __syncwarp();
  }
  for (int j = lane; j < k; j += 32) out[(size_t)q * k + j] = idx[j];
}

// ------------------------------------------------------------------------------------------------
// near-correctly-rounded helpers (double-double); FMA is used explicitly, it is not contraction
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void two_prod(double a, double b, double& hi, double& lo) { hi = a * b; lo = __fma_rn(a, b, -hi); }
__device__ __forceinline__ void dd_mul_d(double ah, double al, double b, double& rh, double& rl) {
  double ph, pl;
  two_prod(ah, b, ph, pl);
  pl = __fma_rn(al, b, pl);
  rh = ph + pl;
  rl = pl - (rh - ph);
}
__device__ double powi_cr(double x, int e) {  // x^e, e >= 0, pow(0,0)=1 like std::pow
  if (e == 0) return 1.0;
  double h = x, l = 0.0;
  for (int i = 1; i < e; i++) dd_mul_d(h, l, x, h, l);
  return h + l;
}
__device__ double rsqrt_cr(double D) {  // D^(-1/2)
  const double r = 1.0 / sqrt(D);
  double th, tl, ph, pl;
  two_prod(r, r, th, tl);
  two_prod(D, th, ph, pl);
  pl = __fma_rn(D, tl, pl);
  const double e = (1.0 - ph) - pl;
  return __fma_rn(0.5 * r, e, r);
}

// ------------------------------------------------------------------------------------------------
// batched weights: one CTA per stencil
// ------------------------------------------------------------------------------------------------
enum { W_LAPLACE = 0, W_DX = 1, W_DY = 2, W_INTERP = 3, W_NORMAL = 4 };

struct WeightJob {
  const double* px; const double* py;     // base-grid points
  const int* nb;                          // [systems][n] neighbour ids (kNN order)
  const int* ids;                         // eval node ids (nullable)
  const double* ex; const double* ey;     // eval points when ids == nullptr
  const double* nrmx; const double* nrmy; // per-node normals (W_NORMAL)
  int systems, n, m, S, polyDeg, mode;
  // outputs (either raw or chunk)
  double* w_out;                          // [systems][n] (nullable)
  unsigned char* chunks; size_t chunk_bytes; int W; int* len; double* diags; int diag_first;  // chunk path (row = system index or ids[sys])
};

struct PivotRec { double v; int rc; };  // rc = (c << 16) | r

__device__ __forceinline__ bool piv_better(double v, int rc, double bv, int brc) { return v > bv || (v == bv && rc < brc); }

template <int TY>
__global__ void __launch_bounds__(32 * TY) k_weights(WeightJob J) {
  extern __shared__ double sm[];
  const int S = J.S, n = J.n, m = J.m;
  double* A = sm;                       // S*S column-major
  double* rhs = A + (size_t)S * S;      // 2*S
  double* sx = rhs + 2 * S;             // n+1 scaled x (eval last)
  double* sy = sx + (n + 1);
  double* red = sy + (n + 1);           // 4*TY reduction scratch
  int* rowT = reinterpret_cast<int*>(red + 4 * TY);
  int* colT = rowT + S;
  int* nbs = colT + S;                  // n
  int* misc = nbs + n;                  // [0]=pivot rc, [1]=nonzero, [2] rank
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx, NT = 32 * TY;
  const int sys = blockIdx.x;
  const int node = J.ids ? J.ids[sys] : -1;
  const double evx = J.ids ? J.px[node] : J.ex[sys];
  const double evy = J.ids ? J.py[node] : J.ey[sys];

  // ---- shifting_scaling (general_computation_functions.cpp:82-107)
  for (int i = tid; i < n; i += NT) {
    const int id = J.nb[(size_t)sys * n + i];
    nbs[i] = id;
    sx[i] = J.px[id];
    sy[i] = J.py[id];
  }
  __syncthreads();
  {
    double mnx = sx[0], mxx = sx[0], mny = sy[0], mxy = sy[0];
    for (int i = tid; i < n; i += NT) { mnx = fmin(mnx, sx[i]); mxx = fmax(mxx, sx[i]); mny = fmin(mny, sy[i]); mxy = fmax(mxy, sy[i]); }
    for (int o = 16; o > 0; o >>= 1) {
      mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
      mny = fmin(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if (tx == 0) { red[4 * ty] = mnx; red[4 * ty + 1] = mxx; red[4 * ty + 2] = mny; red[4 * ty + 3] = mxy; }
    __syncthreads();
    mnx = red[0]; mxx = red[1]; mny = red[2]; mxy = red[3];
    for (int t = 1; t < TY; t++) { mnx = fmin(mnx, red[4 * t]); mxx = fmax(mxx, red[4 * t + 1]); mny = fmin(mny, red[4 * t + 2]); mxy = fmax(mxy, red[4 * t + 3]); }
    __syncthreads();
    const double scale = fmax(mxx - mnx, mxy - mny);
    for (int i = tid; i < n; i += NT) { sx[i] = (sx[i] - mnx) / scale; sy[i] = (sy[i] - mny) / scale; }
    if (tid == 0) { sx[n] = (evx - mnx) / scale; sy[n] = (evy - mny) / scale; red[0] = scale; }
    __syncthreads();
  }
  const double scale = red[0];
  const double xE = sx[n], yE = sy[n];
  __syncthreads();

  // ---- buildCoeffMatrix (grid.cpp:263-299): [Phi P; P^T 0]
  for (int e = tid; e < S * S; e += NT) A[e] = 0.0;
  __syncthreads();
  for (int j = ty; j < n; j += TY)
    for (int i = tx; i <= j; i += 32) {
      const double dx = sx[i] - sx[j], dy = sy[i] - sy[j];
      const double r = sqrt(dx * dx + dy * dy);
      const double a = powi_cr(r, 3);
      A[i + (size_t)j * S] = a;
      A[j + (size_t)i * S] = a;
    }
  for (int row = tid; row < n; row += NT) {
    int col = n;
    const double x = sx[row], y = sy[row];
    for (int p = 0; p <= J.polyDeg; p++)
      for (int q = 0; q <= p; q++) {
        const double pc = powi_cr(x, p - q) * powi_cr(y, q);
        A[row + (size_t)col * S] = pc;
        A[col + (size_t)row * S] = pc;
        col++;
      }
  }
  // ---- right-hand sides (grid.cpp:317-333, 356-372, 394-416, 699-709)
  const int nrhs = J.mode == W_NORMAL ? 2 : 1;
  for (int i = tid; i < 2 * S; i += NT) rhs[i] = 0.0;
  __syncthreads();
  for (int pass = 0; pass < nrhs; pass++) {
    const int mode = J.mode == W_NORMAL ? (pass == 0 ? W_DX : W_DY) : J.mode;
    double* b = rhs + pass * S;
    for (int i = tid; i < n; i += NT) {
      const double xR = sx[i], yR = sy[i];
      if (mode == W_LAPLACE) {
        const double D = (xE * xE - 2 * xE * xR + xR * xR + yE * yE - 2 * yE * yR + yR * yR);
        if (D > 0) {
          const double t1 = 2 * xE - 2 * xR, t2 = 2 * yE - 2 * yR;
          b[i] = (t1 * t1 + t2 * t2) * 1.5 * 0.5 * rsqrt_cr(D) + 6.0 * sqrt(D);
        }
      } else if (mode == W_DX || mode == W_DY) {
        if (i > 0) {
          const double dx = sx[i] - xE, dy = sy[i] - yE;
          const double r = sqrt(dx * dx + dy * dy);
          b[i] = 3.0 * r * (mode == W_DX ? (xE - xR) : (yE - yR));
        }
      } else {  // W_INTERP
        const double dx = xE - sx[i], dy = yE - sy[i];
        b[i] = powi_cr(sqrt(dx * dx + dy * dy), 3);
      }
    }
    if (tid == 0) {
      int row = n;
      for (int p = 0; p <= J.polyDeg; p++)
        for (int q = 0; q <= p; q++) {
          double t = 0;
          if (mode == W_LAPLACE) {
            if (p - q - 2 >= 0) t += (p - q) * (p - q - 1) * powi_cr(xE, p - q - 2) * powi_cr(yE, q);
            if (q - 2 >= 0) t += q * (q - 1) * powi_cr(xE, p - q) * powi_cr(yE, q - 2);
          } else if (mode == W_DX) {
            if (p - q - 1 >= 0) t += (p - q) * powi_cr(xE, p - q - 1) * powi_cr(yE, q);
          } else if (mode == W_DY) {
            if (q - 1 >= 0) t += q * powi_cr(xE, p - q) * powi_cr(yE, q - 1);
          } else {
            t = powi_cr(xE, p - q) * powi_cr(yE, q);
          }
          b[row++] = t;
        }
    }
  }
  __syncthreads();

  // ---- FullPivLU::computeInPlace: pivot = first strict max |a| in a column-major scan of the trailing block
  double bv = -1.0;
  int brc = 0x7fffffff;
  for (int c = ty; c < S; c += TY)
    for (int r = tx; r < S; r += 32) {
      const double v = fabs(A[r + (size_t)c * S]);
      const int rc = (c << 16) | r;
      if (piv_better(v, rc, bv, brc)) { bv = v; brc = rc; }
    }
  int nonzero = S;
  double maxpivot = 0.0;
  for (int k = 0; k < S; k++) {
    // block arg-max of (bv, brc)
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int orc = __shfl_xor_sync(0xffffffffu, brc, o);
      if (piv_better(ov, orc, bv, brc)) { bv = ov; brc = orc; }
    }
    if (tx == 0) { red[2 * ty] = bv; reinterpret_cast<int*>(red + 2 * ty + 1)[0] = brc; }
    __syncthreads();
    bv = red[0]; brc = reinterpret_cast<int*>(red + 1)[0];
    for (int t = 1; t < TY; t++) {
      const double ov = red[2 * t];
      const int orc = reinterpret_cast<int*>(red + 2 * t + 1)[0];
      if (piv_better(ov, orc, bv, brc)) { bv = ov; brc = orc; }
    }
    __syncthreads();
    if (bv == 0.0) {
      nonzero = k;
      for (int i = k + tid; i < S; i += NT) { rowT[i] = i; colT[i] = i; }
      break;
    }
    if (bv > maxpivot) maxpivot = bv;
    const int pr = brc & 0xffff, pcn = brc >> 16;
    if (tid == 0) { rowT[k] = pr; colT[k] = pcn; }
    if (pr != k) for (int c = tid; c < S; c += NT) { const double t = A[k + (size_t)c * S]; A[k + (size_t)c * S] = A[pr + (size_t)c * S]; A[pr + (size_t)c * S] = t; }
    __syncthreads();
    if (pcn != k) for (int r = tid; r < S; r += NT) { const double t = A[r + (size_t)k * S]; A[r + (size_t)k * S] = A[r + (size_t)pcn * S]; A[r + (size_t)pcn * S] = t; }
    __syncthreads();
    bv = -1.0; brc = 0x7fffffff;
    if (k < S - 1) {
      const double piv = A[k + (size_t)k * S];
      __syncthreads();
      for (int r = k + 1 + tid; r < S; r += NT) A[r + (size_t)k * S] /= piv;
      __syncthreads();
      for (int c = k + 1 + ty; c < S; c += TY) {
        const double u = A[k + (size_t)c * S];
        for (int r = k + 1 + tx; r < S; r += 32) {
          const double a = A[r + (size_t)c * S] - A[r + (size_t)k * S] * u;
          A[r + (size_t)c * S] = a;
          const double v = fabs(a);
          const int rc = (c << 16) | r;
          if (piv_better(v, rc, bv, brc)) { bv = v; brc = rc; }
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  // ---- FullPivLU::_solve_impl, rank at threshold eps*size*|maxpivot|
  if (tid == 0) {
    const double thr = fabs(maxpivot) * (2.220446049250313e-16 * S);
    int rank = 0;
    for (int i = 0; i < nonzero; i++) rank += (fabs(A[i + (size_t)i * S]) > thr);
    misc[2] = rank;
    for (int pass = 0; pass < nrhs; pass++) {
      double* c = rhs + pass * S;
      for (int k = 0; k < S; k++) { const double t = c[k]; c[k] = c[rowT[k]]; c[rowT[k]] = t; }
    }
  }
  __syncthreads();
  const int rank = misc[2];
  for (int j = 0; j < S; j++) {  // unit-lower solve, column oriented
    for (int pass = 0; pass < nrhs; pass++) {
      double* c = rhs + pass * S;
      const double cj = c[j];
      for (int i = j + 1 + tid; i < S; i += NT) c[i] -= A[i + (size_t)j * S] * cj;
    }
    __syncthreads();
  }
  for (int j = rank - 1; j >= 0; j--) {  // upper solve on the rank x rank corner
    if (tid < nrhs) rhs[tid * S + j] /= A[j + (size_t)j * S];
    __syncthreads();
    for (int pass = 0; pass < nrhs; pass++) {
      double* c = rhs + pass * S;
      const double cj = c[j];
      for (int i = tid; i < j; i += NT) c[i] -= A[i + (size_t)j * S] * cj;
    }
    __syncthreads();
  }
  if (tid == 0) {
    for (int pass = 0; pass < nrhs; pass++) {
      double* c = rhs + pass * S;
      for (int i = rank; i < S; i++) c[i] = 0.0;
      for (int k = S - 1; k >= 0; k--) { const double t = c[k]; c[k] = c[colT[k]]; c[colT[k]] = t; }
    }
  }
  __syncthreads();
  // ---- scale the kept weights (grid.cpp:337-340, 419-422)
  for (int i = tid; i < n; i += NT) {
    double w;
    if (J.mode == W_LAPLACE) w = rhs[i] / (scale * scale);
    else if (J.mode == W_DX || J.mode == W_DY) w = rhs[i] / scale;
    else if (J.mode == W_NORMAL) {  // grid.cpp:535-537: w = wx*nx; w += ny*wy
      double a = rhs[i] / scale;
      a *= J.nrmx[node];
      a += J.nrmy[node] * (rhs[S + i] / scale);
      w = a;
    } else w = rhs[i];
    rhs[i] = w;
  }
  __syncthreads();
  if (J.w_out)
    for (int i = tid; i < n; i += NT) J.w_out[(size_t)sys * n + i] = rhs[i];
  if (J.chunks) {  // row chunk: columns ascending (setFromTriplets order), diagonal first when asked
    const int row = node >= 0 ? node : sys;
    double* cv = reinterpret_cast<double*>(J.chunks + (size_t)row * J.chunk_bytes);
    int* cc = reinterpret_cast<int*>(J.chunks + (size_t)row * J.chunk_bytes + (size_t)J.W * 8);
    for (int i = tid; i < n; i += NT) {
      const int col = nbs[i];
      int pos = 0;
      for (int j = 0; j < n; j++) pos += (nbs[j] < col);
      if (J.diag_first) {
        if (col == row) { pos = 0; if (J.diags) J.diags[row] = rhs[i]; }
        else if (col < row) pos += 1;
      }
      cv[pos] = rhs[i];
      cc[pos] = col;
    }
    for (int i = n + tid; i < J.W; i += NT) { cv[i] = 0.0; cc[i] = row < 0 ? 0 : min(row, 0x7ffffffe); }
    if (tid == 0 && J.len) J.len[row] = n;
  }
}

size_t weights_smem(int S, int n, int TY) { return sizeof(double) * ((size_t)S * S + 2 * S + 2 * (n + 1) + 4 * TY) + sizeof(int) * (2 * S + n + 4); }

void launch_weights(Grid& g, WeightJob& J) {
  if (J.systems == 0) return;
  MMG_REQUIRE(g.props.rbfExp == 3, MMG_ERR_ARG, "only the PHS r^3 kernel (rbfExp == 3, the reference's setting) is implemented on the device");
  MMG_REQUIRE(J.n <= 160 && J.S < 65535, MMG_ERR_ARG, "stencil too large for the batched LU kernel");
  auto go = [&](auto tyc) {
    constexpr int TY = decltype(tyc)::value;
    const size_t smem = weights_smem(J.S, J.n, TY);
    static size_t configured = 0;
    if (smem > configured) {
      MMG_CUDA(cudaFuncSetAttribute(k_weights<TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 200 * 1024)));
      configured = std::max<size_t>(smem, 200 * 1024);
    }
    MMG_REQUIRE(smem <= 227 * 1024, MMG_ERR_ARG, "local RBF system does not fit in shared memory");
    k_weights<TY><<<J.systems, dim3(32, TY), smem, g.stream>>>(J);
  };
  if (J.S <= 40) go(std::integral_constant<int, 2>());
  else if (J.S <= 64) go(std::integral_constant<int, 4>());
  else go(std::integral_constant<int, 8>());
  MMG_CUDA(cudaGetLastError());
}

void knn_device(Grid& g, int m, const double* qx_dev, const double* qy_dev, const int* qflag_dev, int neumann, int k, int* out_dev) {
  AsmState& st = state(g);
  if (!st.cells_valid) build_cells(g);
  MMG_REQUIRE(k >= 1 && k <= g.n && k <= KNN_CAP / 2, MMG_ERR_ARG, "kNearestNeighbors: k out of range");
  if (m == 0) return;
  k_knn<<<(m + KNN_WARPS - 1) / KNN_WARPS, KNN_WARPS * 32, 0, g.stream>>>(st.view(), g.px.p, g.py.p, st.dflags.p, m, qx_dev, qy_dev, qflag_dev, neumann, k,
                                                                        out_dev, st.err.p);
  MMG_CUDA(cudaGetLastError());
  int e = 0;
  st.err.download(&e, 1, g.stream);
  if (e) {
    st.err.zero(g.stream);
    throw Error(e == 1 ? MMG_ERR_STATE : MMG_ERR_ARG, e == 1 ? "kNearestNeighbors: candidate buffer overflow (point density far from uniform)"
                                                             : "kNearestNeighbors: fewer admissible points than the stencil size");
  }
}

int poly_terms(int p) { return (p + 1) * (p + 2) / 2; }
int stencil_of(int p) { return (int)(2.5 * (p + 1) * (p + 2) / 2); }  // grid.cpp:267

// Eigen setFromTriplets semantics on the host for the O(sqrt N)-fill Neumann path
struct Trip { int r, c; double v; };
void csr_from_triplets(HostCsr& A, int rows, int cols, const std::vector<Trip>& t) {
  A.rows = rows; A.cols = cols;
  std::vector<int64_t> cnt(rows + 1, 0);
  for (const Trip& e : t) cnt[e.r + 1]++;
  for (int r = 0; r < rows; r++) cnt[r + 1] += cnt[r];
  std::vector<int64_t> pos(cnt.begin(), cnt.end() - 1);
  std::vector<std::pair<int, double>> tmp(t.size());
  for (const Trip& e : t) tmp[pos[e.r]++] = {e.c, e.v};
  A.ptr.assign(rows + 1, 0); A.idx.clear(); A.val.clear();
  for (int r = 0; r < rows; r++) {
    auto b = tmp.begin() + cnt[r], e = tmp.begin() + cnt[r + 1];
    std::stable_sort(b, e, [](const std::pair<int, double>& p, const std::pair<int, double>& q) { return p.first < q.first; });
    for (auto it = b; it != e;) {
      double s = it->second;
      auto jt = it + 1;
      for (; jt != e && jt->first == it->first; ++jt) s = s + jt->second;
      A.idx.push_back(it->first); A.val.push_back(s);
      it = jt;
    }
    A.ptr[r + 1] = (int)A.idx.size();
  }
}

}  // namespace

void asm_release(Grid& g) {
  delete static_cast<AsmState*>(g.asm_state);
  g.asm_state = nullptr;
}

void asm_knn_points(Grid& g, int m, const double* qx, const double* qy, const int* qflag, int neumann, int k, int* out_host) {
  DevBuf<double> dx, dy;
  DevBuf<int> df, dout;
  dx.upload(qx, m, g.stream); dy.upload(qy, m, g.stream);
  if (qflag) df.upload(qflag, m, g.stream);
  dout.alloc((size_t)m * k);
  knn_device(g, m, dx.p, dy.p, qflag ? df.p : nullptr, neumann, k, dout.p);
  dout.download(out_host, (size_t)m * k, g.stream);
}

// Grid::rcm_order_points grid.cpp:713-776.  kNN on the device; the BFS (a queue algorithm whose visiting order is the
// result) runs on the host over the downloaded lists.
void asm_rcm_order_points(Grid& g) {
  AsmState& st = state(g);
  const int N = g.n, k = g.props.stencilSize;
  st.cells_valid = false;
  build_cells(g);
  DevBuf<int> dout;
  dout.alloc((size_t)N * k);
  knn_device(g, N, g.px.p, g.py.p, st.dflags.p, g.neumann ? 1 : 0, k, dout.p);
  std::vector<int> nn = dout.to_host(g.stream);
  dout.release();
  std::vector<std::vector<int>> extra;  // appended adjacency of interior rows next to Neumann nodes (grid.cpp:722-740)
  const bool augment = g.neumann && g.implicit;
  if (augment) {
    extra.resize(N);
    std::vector<int> adj;
    for (int i = 0; i < N; i++) {
      if (g.bcflags[i] != 0) continue;
      adj.assign(nn.begin() + (size_t)i * k, nn.begin() + (size_t)(i + 1) * k);
      for (size_t j = 0; j < adj.size(); j++) {
        const int nbn = adj[j];
        if (g.bcflags[nbn] != 2) continue;
        for (int t = 0; t < k; t++) {
          const int cand = nn[(size_t)nbn * k + t];
          if (std::find(adj.begin(), adj.end(), cand) == adj.end()) adj.push_back(cand);
        }
      }
      extra[i].assign(adj.begin() + k, adj.end());
    }
  }
  std::vector<char> seen(N, 0);
  std::vector<int> visit;
  visit.reserve(N);
  std::queue<int> q;
  seen[0] = 1; q.push(0);
  while (!q.empty()) {  // general_computation_functions.cpp:108-129
    const int cur = q.front(); q.pop();
    visit.push_back(cur);
    for (int t = 0; t < k; t++) { const int a = nn[(size_t)cur * k + t]; if (!seen[a]) { seen[a] = 1; q.push(a); } }
    if (augment) for (int a : extra[cur]) if (!seen[a]) { seen[a] = 1; q.push(a); }
  }
  MMG_REQUIRE((int)visit.size() == N, MMG_ERR_STATE, "rcm_order_points: BFS from node 0 did not reach every node (the reference reads past the end here)");
  std::reverse(visit.begin(), visit.end());
  g.order = visit;
  std::vector<double> nx(N), ny(N), nnx(N), nny(N), src(g.A);
  std::vector<int> nf(N), old2new(N);
  g.b.download(src.data(), g.A, g.stream);
  std::vector<double> nsrc = src;
  for (int i = 0; i < N; i++) {
    const int o = visit[i];
    nx[i] = g.hx[o]; ny[i] = g.hy[o]; nf[i] = g.bcflags[o]; nnx[i] = g.hnx[o]; nny[i] = g.hny[o]; nsrc[i] = src[o];
    old2new[o] = i;
  }
  g.hx = nx; g.hy = ny; g.bcflags = nf; g.hnx = nnx; g.hny = nny;
  for (Boundary& b : g.boundaries) for (int& p : b.pts) p = old2new[p];
  g.b.upload(nsrc.data(), g.A, g.stream);
  g.px.upload(g.hx, g.stream); g.py.upload(g.hy, g.stream);
  g.sync();
  st.cells_valid = false;
}

static void node_weights(Grid& g, int mode, int m, const int* ids_dev, const int* nb_dev, double* w_dev, const double* nrmx, const double* nrmy) {
  WeightJob J{};
  J.px = g.px.p; J.py = g.py.p; J.nb = nb_dev; J.ids = ids_dev; J.nrmx = nrmx; J.nrmy = nrmy;
  J.systems = m; J.n = g.props.stencilSize; J.m = poly_terms(g.props.polyDeg); J.S = J.n + J.m; J.polyDeg = g.props.polyDeg; J.mode = mode;
  J.w_out = w_dev;
  launch_weights(g, J);
}

void asm_weights(Grid& g, int which, int m, const int* ids, double* w, int* nb) {
  AsmState& st = state(g);
  if (!st.cells_valid) build_cells(g);
  MMG_REQUIRE(g.props.stencilSize == stencil_of(g.props.polyDeg), MMG_ERR_ARG, "stencilSize must equal (int)(2.5*(p+1)(p+2)/2) (grid.cpp:267 vs :309)");
  const int n = g.props.stencilSize;
  std::vector<double> qx(m), qy(m);
  std::vector<int> qf(m);
  for (int i = 0; i < m; i++) {
    MMG_REQUIRE(ids[i] >= 0 && ids[i] < g.n, MMG_ERR_ARG, "node id out of range");
    qx[i] = g.hx[ids[i]]; qy[i] = g.hy[ids[i]]; qf[i] = g.bcflags[ids[i]] != 0;
  }
  DevBuf<double> dqx, dqy, dw;
  DevBuf<int> dqf, dnb, dids;
  dqx.upload(qx, g.stream); dqy.upload(qy, g.stream); dqf.upload(qf, g.stream); dids.upload(ids, m, g.stream);
  dnb.alloc((size_t)m * n); dw.alloc((size_t)m * n);
  knn_device(g, m, dqx.p, dqy.p, dqf.p, g.neumann ? 1 : 0, n, dnb.p);
  const int mode = which == MMG_MAT_LAPLACE ? W_LAPLACE : which == MMG_MAT_DERIVX ? W_DX : which == MMG_MAT_DERIVY ? W_DY : -1;
  MMG_REQUIRE(mode >= 0, MMG_ERR_ARG, "weights: which must be MMG_MAT_LAPLACE, MMG_MAT_DERIVX or MMG_MAT_DERIVY");
  node_weights(g, mode, m, dids.p, dnb.p, dw.p, nullptr, nullptr);
  dw.download(w, (size_t)m * n, g.stream);
  dnb.download(nb, (size_t)m * n, g.stream);
}

void asm_point_interp_weights(Grid& g, int m, const double* px, const double* py, int polyDeg, double* w, int* nb) {
  AsmState& st = state(g);
  if (!st.cells_valid) build_cells(g);
  const int n = stencil_of(polyDeg);
  DevBuf<double> dqx, dqy, dw;
  DevBuf<int> dnb;
  dqx.upload(px, m, g.stream); dqy.upload(py, m, g.stream);
  dnb.alloc((size_t)m * n); dw.alloc((size_t)m * n);
  knn_device(g, m, dqx.p, dqy.p, nullptr, 0, n, dnb.p);
  WeightJob J{};
  J.px = g.px.p; J.py = g.py.p; J.nb = dnb.p; J.ex = dqx.p; J.ey = dqy.p;
  J.systems = m; J.n = n; J.m = poly_terms(polyDeg); J.S = J.n + J.m; J.polyDeg = polyDeg; J.mode = W_INTERP; J.w_out = dw.p;
  launch_weights(g, J);
  dw.download(w, (size_t)m * n, g.stream);
  dnb.download(nb, (size_t)m * n, g.stream);
}

// Multigrid::buildInterpMatrix multigrid.cpp:17-33: row t = interpolation weights from base-grid nodes to target point t
void asm_build_interp(Grid& base, Grid& target, int polyDeg, HybMatrix& M) {
  AsmState& st = state(base);
  if (!st.cells_valid) build_cells(base);
  MMG_REQUIRE(base.device == target.device, MMG_ERR_ARG, "buildInterpMatrix: grids live on different devices");
  const int n = stencil_of(polyDeg), T = target.n;
  DevBuf<int> dnb;
  dnb.alloc((size_t)T * n);
  knn_device(base, T, target.px.p, target.py.p, nullptr, 0, n, dnb.p);
  M.rows = T; M.cols = base.n; M.W = n; M.diag_first = false; M.nnz = (int64_t)T * n;
  M.chunk_bytes = ((size_t)n * 12 + 31) / 32 * 32;
  M.chunks.alloc((size_t)T * M.chunk_bytes);
  M.len.alloc(T);
  M.n_ovf = 0; M.reg_row = -1; M.reg_len = 0;
  WeightJob J{};
  J.px = base.px.p; J.py = base.py.p; J.nb = dnb.p; J.ex = target.px.p; J.ey = target.py.p;
  J.systems = T; J.n = n; J.m = poly_terms(polyDeg); J.S = J.n + J.m; J.polyDeg = polyDeg; J.mode = W_INTERP;
  J.chunks = M.chunks.p; J.chunk_bytes = M.chunk_bytes; J.W = n; J.len = M.len.p; J.diag_first = 0;
  launch_weights(base, J);
  base.sync();
}

// FractionalStepGrid::build_derivX_mat / build_derivY_mat / build_uv_laplace_mat (fractionalStepGrid.cpp:60-100): N x N, one
// stencil row per node (no boundary-condition rows), the same neighbour lists for the three operators
void asm_build_fs_operators(Grid& g) {
  MMG_REQUIRE(g.fs != nullptr, MMG_ERR_STATE, "mmg_grid_fs_init has not been called");
  AsmState& st = state(g);
  if (!st.cells_valid) build_cells(g);
  MMG_REQUIRE(g.props.stencilSize == stencil_of(g.props.polyDeg), MMG_ERR_ARG, "stencilSize must equal (int)(2.5*(p+1)(p+2)/2)");
  const int N = g.n, n = g.props.stencilSize;
  DevBuf<int> dnb, dids, dqf;
  dnb.alloc((size_t)N * n);
  std::vector<int> qf(N), ids(N);
  for (int i = 0; i < N; i++) { qf[i] = g.bcflags[i] != 0; ids[i] = i; }
  dqf.upload(qf, g.stream); dids.upload(ids, g.stream);
  knn_device(g, N, g.px.p, g.py.p, dqf.p, g.neumann ? 1 : 0, n, dnb.p);
  HybMatrix* mats[3] = {&g.fs->Dx, &g.fs->Dy, &g.fs->Lap};
  const int modes[3] = {W_DX, W_DY, W_LAPLACE};
  for (int o = 0; o < 3; o++) {
    HybMatrix& M = *mats[o];
    M = HybMatrix();
    M.rows = N; M.cols = N; M.W = n; M.diag_first = false; M.nnz = (int64_t)N * n;
    M.chunk_bytes = ((size_t)n * 12 + 31) / 32 * 32;
    M.chunks.alloc((size_t)N * M.chunk_bytes);
    M.len.alloc(N);
    WeightJob J{};
    J.px = g.px.p; J.py = g.py.p; J.nb = dnb.p; J.ids = dids.p;
    J.systems = N; J.n = n; J.m = poly_terms(g.props.polyDeg); J.S = J.n + J.m; J.polyDeg = g.props.polyDeg; J.mode = modes[o];
    J.chunks = M.chunks.p; J.chunk_bytes = M.chunk_bytes; J.W = n; J.len = M.len.p; J.diag_first = 0;
    launch_weights(g, J);
  }
  g.sync();
  g.fs->have_ops = true;
}

// Grid::build_deriv_normal_bound grid.cpp:520-548
void asm_build_deriv_normal_bound(Grid& g) {
  AsmState& st = state(g);
  if (!st.cells_valid) build_cells(g);
  st.dn_point.clear();
  for (const Boundary& b : g.boundaries)
    if (b.type == MMG_BC_NEUMANN) st.dn_point.insert(st.dn_point.end(), b.pts.begin(), b.pts.end());
  const int m = (int)st.dn_point.size(), n = g.props.stencilSize;
  st.dn_w.assign((size_t)m * n, 0.0); st.dn_nb.assign((size_t)m * n, 0);
  if (m == 0) return;
  std::vector<double> qx(m), qy(m);
  std::vector<int> qf(m, 1);
  for (int i = 0; i < m; i++) { qx[i] = g.hx[st.dn_point[i]]; qy[i] = g.hy[st.dn_point[i]]; }
  DevBuf<double> dqx, dqy, dw, dnx, dny;
  DevBuf<int> dqf, dnb, dids;
  dqx.upload(qx, g.stream); dqy.upload(qy, g.stream); dqf.upload(qf, g.stream); dids.upload(st.dn_point, g.stream);
  dnx.upload(g.hnx, g.stream); dny.upload(g.hny, g.stream);
  dnb.alloc((size_t)m * n); dw.alloc((size_t)m * n);
  knn_device(g, m, dqx.p, dqy.p, dqf.p, g.neumann ? 1 : 0, n, dnb.p);
  node_weights(g, W_NORMAL, m, dids.p, dnb.p, dw.p, dnx.p, dny.p);
  dw.download(st.dn_w.data(), (size_t)m * n, g.stream);
  dnb.download(st.dn_nb.data(), (size_t)m * n, g.stream);
}

// Grid::build_laplacian grid.cpp:549-663
void asm_build_laplacian(Grid& g) {
  AsmState& st = state(g);
  if (!st.cells_valid) build_cells(g);
  MMG_REQUIRE(g.props.stencilSize == stencil_of(g.props.polyDeg), MMG_ERR_ARG, "stencilSize must equal (int)(2.5*(p+1)(p+2)/2) (grid.cpp:267 vs :386)");
  const int N = g.n, n = g.props.stencilSize;
  DevBuf<int> dnb, dids;
  dnb.alloc((size_t)N * n);
  {
    std::vector<int> qf(N);
    for (int i = 0; i < N; i++) qf[i] = g.bcflags[i] != 0;
    DevBuf<int> dqf;
    dqf.upload(qf, g.stream);
    knn_device(g, N, g.px.p, g.py.p, dqf.p, g.neumann ? 1 : 0, n, dnb.p);
  }
  std::vector<int> ids(N);
  std::iota(ids.begin(), ids.end(), 0);
  dids.upload(ids, g.stream);
  if (!g.neumann) {
    // Dirichlet grid: every row is a plain stencil row -> emit row chunks straight from the LU kernel
    HybMatrix& M = g.Lap;
    M = HybMatrix();
    M.rows = N; M.cols = N; M.W = n; M.diag_first = true; M.nnz = (int64_t)N * n;
    M.chunk_bytes = ((size_t)n * 12 + 31) / 32 * 32;
    M.chunks.alloc((size_t)N * M.chunk_bytes);
    M.len.alloc(N);
    DevBuf<double> ddiag;
    ddiag.alloc(N);
    WeightJob J{};
    J.px = g.px.p; J.py = g.py.p; J.nb = dnb.p; J.ids = dids.p;
    J.systems = N; J.n = n; J.m = poly_terms(g.props.polyDeg); J.S = J.n + J.m; J.polyDeg = g.props.polyDeg; J.mode = W_LAPLACE;
    J.chunks = M.chunks.p; J.chunk_bytes = M.chunk_bytes; J.W = n; J.len = M.len.p; J.diags = ddiag.p; J.diag_first = 1;
    launch_weights(g, J);
    g.diags.assign(g.A, 0.0);
    ddiag.download(g.diags.data(), N, g.stream);
    g.nbc = HostCsr();
    g.have_laplacian = true; g.have_colours = false; g.have_blocks = false; g.lex_neu_ok = -1;
    return;
  }
  // Neumann / mixed grid: weights from the device, triplet bookkeeping of grid.cpp:553-661 on the host
  DevBuf<double> dw;
  dw.alloc((size_t)N * n);
  node_weights(g, W_LAPLACE, N, dids.p, dnb.p, dw.p, nullptr, nullptr);
  std::vector<double> W = dw.to_host(g.stream);
  std::vector<int> NB = dnb.to_host(g.stream);
  dw.release(); dnb.release();
  std::vector<Trip> trip, bnd;
  trip.reserve((size_t)N * (n + 2));
  g.diags.assign(g.A, 0.0);
  for (int i = 0; i < N; i++) {
    if (g.bcflags[i] != 2) {
      for (int j = 0; j < n; j++) {
        const int c = NB[(size_t)i * n + j];
        const double w = W[(size_t)i * n + j];
        trip.push_back(Trip{i, c, w});
        if (g.bcflags[i] == 0 && g.bcflags[c] == 2) bnd.push_back(Trip{i, c, w});
        if (i == c) g.diags[i] = w;
      }
      trip.push_back(Trip{i, N, 1.0});
    }
  }
  for (int i = 0; i < N + 1; i++)
    if (i == N || g.bcflags[i] != 2) trip.push_back(Trip{N, i, 1.0});
  {
    size_t expect = 0;
    for (const Boundary& b : g.boundaries) if (b.type == MMG_BC_NEUMANN) expect += b.pts.size();
    MMG_REQUIRE(st.dn_point.size() == expect, MMG_ERR_STATE, "build_laplacian: build_deriv_normal_bound() must run first on a Neumann grid");
  }
  for (size_t t = 0; t < st.dn_point.size(); t++)
    for (int j = 0; j < n; j++) {
      const int c = st.dn_nb[t * n + j];
      trip.push_back(Trip{st.dn_point[t], c, st.dn_w[t * n + j]});
      if (st.dn_point[t] == c) g.diags[st.dn_point[t]] = st.dn_w[t * n + j];
    }
  HostCsr A;
  csr_from_triplets(A, g.A, g.A, trip);
  csr_from_triplets(g.nbc, g.A, g.A, bnd);
  if (g.implicit) {  // grid.cpp:598-661
    std::vector<std::pair<int, double>> rowBnd;
    for (int i = 0; i < A.rows - 1; i++) {
      if (g.bcflags[i] != 0) continue;
      rowBnd.clear();
      for (int j = A.ptr[i]; j < A.ptr[i + 1]; j++)
        if (A.idx[j] != A.rows - 1 && g.bcflags[A.idx[j]] == 2) rowBnd.push_back({A.idx[j], A.val[j]});
      for (auto& e : rowBnd) {
        const int j_col = e.first;
        const double A_ij = e.second, A_jj = g.diags[j_col];
        for (int kk = A.ptr[j_col]; kk < A.ptr[j_col + 1]; kk++) {
          if (A.idx[kk] == j_col) continue;
          trip.push_back(Trip{i, A.idx[kk], -A.val[kk] * A_ij / A_jj});
        }
        trip.push_back(Trip{i, j_col, -A_ij});
      }
    }
    csr_from_triplets(A, g.A, g.A, trip);
  }
  hyb_from_csr(g.Lap, A, true, true, g.stream);
  g.have_laplacian = true; g.have_colours = false; g.have_blocks = false; g.lex_neu_ok = -1;
}

}  // namespace mmg
