// Internal structures of libmmg (sm_100a).  Host-side C++ only; kernels live in the .cu files.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mmg.h"

#define MMG_STR2(x) #x
#define MMG_STR(x) MMG_STR2(x)

namespace mmg {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define MMG_CUDA(call)                                                                                         \
  do {                                                                                                         \
    cudaError_t e__ = (call);                                                                                  \
    if (e__ != cudaSuccess)                                                                                    \
      throw ::mmg::Error(MMG_ERR_CUDA, std::string(#call) + " failed: " + cudaGetErrorString(e__) + " at " + \
                                           __FILE__ + ":" + std::to_string(__LINE__));                          \
  } while (0)
#define MMG_REQUIRE(cond, code, msg) \
  do {                               \
    if (!(cond)) throw ::mmg::Error(code, msg); \
  } while (0)

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) MMG_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
  }
  void zero(cudaStream_t s) { if (n) MMG_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  void upload(const T* h, size_t count, cudaStream_t s) {
    if (count != n) alloc(count);
    if (count) MMG_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T>& h, cudaStream_t s) { upload(h.data(), h.size(), s); }
  void download(T* h, size_t count, cudaStream_t s) const {
    if (count) MMG_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    MMG_CUDA(cudaStreamSynchronize(s));
  }
  std::vector<T> to_host(cudaStream_t s) const { std::vector<T> h(n); download(h.data(), n, s); return h; }
};

// Host CSR as Eigen leaves it: compressed rows, columns ascending, explicit zeros kept.
struct HostCsr {
  int rows = 0, cols = 0;
  std::vector<int> ptr, idx;
  std::vector<double> val;
  int64_t nnz() const { return (int64_t)idx.size(); }
};

// Device view of a "row-chunk HYB" matrix:
//   row r owns one 32-byte-aligned chunk [ W x fp64 values | W x int32 columns | pad ], its first
//   min(len,W) entries; when diag_first the diagonal sits in slot 0 and the rest stay in ascending
//   column order.  Rows longer than W spill the tail into a small CSR (ovf_*), found by binary search
//   over ovf_rows.  The dense regularisation row of a Neumann grid (grid.cpp:570-576) is kept apart
//   (reg_*), because it is a global reduction rather than a stencil row.
struct HybView {
  const unsigned char* chunks;
  size_t chunk_bytes;
  int W;
  int rows;        // rows stored as chunks (excludes the regularisation row)
  const int* len;  // stored entries per row (may exceed W)
  int n_ovf;
  const int* ovf_rows;
  const int* ovf_ptr;
  const int* ovf_col;
  const double* ovf_val;
};

struct HybMatrix {
  int rows = 0, cols = 0, W = 0;
  size_t chunk_bytes = 0;
  bool diag_first = false;
  int64_t nnz = 0;  // stored entries incl. the regularisation row
  DevBuf<unsigned char> chunks;
  DevBuf<int> len;
  int n_ovf = 0;
  DevBuf<int> ovf_rows, ovf_ptr, ovf_col;
  DevBuf<double> ovf_val;
  // regularisation row (index reg_row == rows when present, -1 otherwise); diagonal excluded from reg_col/val
  int reg_row = -1;
  int reg_len = 0;
  double reg_diag = 0;
  DevBuf<int> reg_col;
  DevBuf<double> reg_val;
  HybView view() const {
    return HybView{chunks.p, chunk_bytes, W, rows, len.p, n_ovf, ovf_rows.p, ovf_ptr.p, ovf_col.p, ovf_val.p};
  }
  // algorithmic bytes of one pass over the matrix: 12 B per stored entry
  int64_t matrix_bytes() const { return nnz * 12; }
};

// Build the device HYB from a host CSR.  has_reg: the last row is the dense regularisation row.
void hyb_from_csr(HybMatrix& M, const HostCsr& A, bool diag_first, bool has_reg, cudaStream_t s);
// Rebuild the host CSR (Eigen order) from the device HYB.
void hyb_to_csr(const HybMatrix& M, HostCsr& A, cudaStream_t s);

struct Boundary {
  int type = 0;
  std::vector<int> pts;
  std::vector<double> vals;
};

constexpr int kMaxLevels = 24;
struct Timers {
  bool on = false;
  double ms[kMaxLevels][MMG_T_COUNT] = {};
  int64_t launches[kMaxLevels][MMG_T_COUNT] = {};
  int64_t bytes[kMaxLevels][MMG_T_COUNT] = {};
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending;   // (level*MMG_T_COUNT+class, events)
  std::vector<cudaEvent_t> pool;
  int64_t total_launches = 0;
};

struct Grid {
  int device = 0;
  int level = 0;    // position inside the owning solver (0 = coarsest); only used to bucket the timers
  void* owner = nullptr;   // the Solver that took ownership (Multigrid::addGrid, multigrid.cpp:10-16); guards double adds / frees
  int n = 0;        // laplaceMatSize_
  int A = 0;        // rows of laplaceMat_ (n, or n+1 with any Neumann boundary)
  bool neumann = false, implicit = false;
  bool exact = true;   // reference-order arithmetic (bit-faithful parity mode) vs reordered warp reductions (throughput mode)
  mmg_props props{};
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  Timers* timers = nullptr;   // shared with the owning solver (or own)
  Timers own_timers;

  // host mirrors of the small / integer state
  std::vector<double> hx, hy, hnx, hny;
  std::vector<int> bcflags;
  std::vector<Boundary> boundaries;
  std::vector<int> order;      // rcm permutation (new -> old)
  std::vector<double> diags;   // Grid::diags
  HostCsr nbc;                 // neumann_boundary_coeffs_ (only rows near Neumann boundaries are non-empty)
  bool have_laplacian = false;
  int lex_neu_ok = -1;         // cached: may a Neumann-type grid use the pipelined lexicographic kernel (-1 = not decided yet)

  // device state
  DevBuf<double> xs;                     // versioned vectors of the sweep-pipelined lexicographic kernel
  DevBuf<double> x, x_alt, b, r;         // values_, sweep scratch, source_, residual scratch (A entries)
  DevBuf<double> px, py;                 // points_
  DevBuf<unsigned char> rowflag;         // bcFlags_ per row as uint8 (A entries; regularisation row = 0)
  DevBuf<int> dir_pts, neu_pts;          // Dirichlet / Neumann node lists in boundary-list order
  DevBuf<double> dir_vals, neu_vals;
  HybMatrix Lap;                         // laplaceMat_
  DevBuf<double> partials;               // reduction scratch
  DevBuf<double> reg_prod;               // products of the regularisation row (reference-order dot product, two stages)
  DevBuf<int> abort_flag;
  // multicolour schedule
  bool have_colours = false;
  int n_colours = 0;
  std::vector<int> colour_ptr;           // n_colours+1
  DevBuf<int> colour_rows;               // rows grouped by colour, ascending inside a colour
  DevBuf<int> colour_ptr_dev;            // colour_ptr on the device (persistent multicolour kernel)
  std::vector<int> colour_host;          // per-row colour (-1 skipped)
  std::vector<int> colour_rows_host;     // host copy of colour_rows (multi-GPU sub-ranges)
  int mc_row0 = 0, mc_row1 = -1;         // rows the packed copy covers (a rank's block on a partitioned level; -1 = all)
  std::vector<int> mc_colour_ptr;        // colour offsets inside the packed copy
  DevBuf<int> mc_colour_ptr_dev;
  bool mc_packed = false;
  int mc_bnd_phase = -1;                 // Neumann-type grid: index (inside a sweep) of the boundary-evaluation / regularisation phase of the packed copy
  DevBuf<double> mc_reg_partial;         // per sweep and CTA: partial dot products of the regularisation row
  DevBuf<int> mc_ctl;                    // tile tickets + colour-barrier arrival counters of the TMA-fed sweep (zeroed per launch)
  DevBuf<unsigned char> mc_chunks;       // colour-major packed copy of Lap.chunks (Morton order inside a colour), fast multicolour sweep
  // block-lexicographic schedule
  int block_size = 4096;
  bool have_blocks = false;
  int n_blk_colours = 0;
  std::vector<int> blk_colour, blk_phase_ptr;
  DevBuf<int> blk_phase_blocks;          // block ids grouped by colour
  // FractionalStepGrid state (fractionalStepGrid.hpp:4-30): created by mmg_grid_fs_init
  struct FracStep {
    double dt = 0, mu = 1, rho = 1;
    DevBuf<double> vec[6];               // u, v, u_old, v_old, u_hat, v_hat (N entries each)
    DevBuf<double> t0, t1, t2;           // SpMV scratch
    HybMatrix Dx, Dy, Lap;               // derivXMat_, derivYMat_, uvLaplaceMat_
    bool have_ops = false;
    DevBuf<int> bnd_pts;                 // every boundary node, boundary-list order (set_ppe_source visits all boundaries)
    DevBuf<double> nx, ny;               // normalVecs_
  };
  FracStep* fs = nullptr;
  // per-level assembly state (device kNN lists etc.) lives in assembly.cu
  void* asm_state = nullptr;

  ~Grid();
  void sync() { MMG_CUDA(cudaStreamSynchronize(stream)); }
};

// halo exchange of contiguous index ranges of a full-length vector
struct ExchangePlan {
  struct Msg { int peer, offset, count; };
  std::vector<Msg> sends, recvs;
};
// how one level is spread over the ranks: contiguous row blocks in the reference order
struct LevelDist {
  bool partitioned = false;
  std::vector<int> bounds;                       // world+1 row offsets: rank r owns [bounds[r], bounds[r+1])
  ExchangePlan x_plan;                           // entries of values_ that my rows of laplaceMat_ read from other ranks
  std::vector<int> r_bounds;                     // split of the restriction's OUTPUT rows (coarse rows of level-1) among ranks
  ExchangePlan r_plan;                           // entries of this level's residual that my restriction rows read
  ExchangePlan p_plan;                           // entries of the coarser level's values_ that my prolongation rows read
  std::vector<std::pair<int, int>> colour_sub;   // per colour: (first, count) of my rows inside colour_rows
  // Peer-memory smoother: two sets of (iters+1) versioned vectors in an IPC-shared allocation; the
  // barrier-free sweep stores a boundary row's new value into the neighbour rank's copy as well (the value is the flag).
  bool peer_ready = false;
  int peer_iters = 0;                            // versions per set - 1
  size_t peer_stride = 0;
  DevBuf<double> peer_xs;                        // 2 x (peer_iters+1) x peer_stride
  int peer_parity = 0;                           // which set the next smoothing call uses (same sequence on every rank)
  int n_sends = 0;                               // <= 2 neighbours
  int send_lo[2] = {0, 0}, send_hi[2] = {0, 0};
  double* send_base[2] = {nullptr, nullptr};     // the neighbour's peer_xs (cudaIpcOpenMemHandle)
};
// by-value kernel argument of the peer-memory sweep
struct PeerSends {
  int n;
  int lo[2], hi[2];
  double* base[2];
};

struct Solver {
  int flavour = MMG_FLAVOUR_MULTIGRID;
  int rank = 0, world = 1;
  void* nccl_comm = nullptr;
  int64_t comm_msgs = 0, comm_bytes = 0;
  int part_threshold = 200000;                   // levels with fewer rows are replicated on every rank
  bool dist_ready = false;
  std::vector<LevelDist> dist;
  DevBuf<double> sums;                           // [num, den] of the distributed residual norm
  int smoother = MMG_SMOOTHER_LEXICOGRAPHIC;
  int arithmetic = MMG_ARITH_REFERENCE_ORDER;
  std::vector<Grid*> grids;                 // sorted ascending by (size, pointer), like multigrid.cpp:116-122
  std::vector<HybMatrix*> restrict_, prolong_;  // [i] as in the reference (restrict_[0]==nullptr, prolong_.back()==nullptr)
  DevBuf<double> hist;                      // residuals_ on device
  int hist_len = 0;
  std::vector<double> hist_host;
  cudaStream_t stream = nullptr;
  Timers timers;
  ~Solver();
};

// ---- operations (kernels.cu) -------------------------------------------------------------------
void op_residual(Grid& g, double* r_out);                        // r = b - A x, Dirichlet rows zeroed (device ptr, A entries)
void op_residual_norm(Grid& g, double* ratio_dev);               // *ratio_dev = |b-Ax|_1 / |b|_1
void op_sor(Grid& g, int smoother);                              // props.iters sweeps + bound_eval_neumann after each
void op_bound_eval_neumann(Grid& g);
void op_boundary_op(Grid& g, int coarse);
void op_modify_coeff_neumann(Grid& g, int coarse);
void op_fix_vector_bound_coarse(Grid& g, double* vec_dev);
void op_zero_values(Grid& g);
// coarse.b[0:Nc] = R * fine_residual[0:Nf]; then the masks of multigrid.cpp:82-86
void op_restrict(Grid& fine, Grid& coarse, const HybMatrix& R, const double* fine_res_dev);
// fine.x[0:Nf] += P * coarse.x[0:Nc] (Dirichlet rows masked when the fine grid is not Neumann) multigrid.cpp:102-106
void op_prolong_correct(Grid& fine, Grid& coarse, const HybMatrix& P);
void op_spmv(const HybMatrix& M, const double* x_dev, double* y_dev, Grid& ctx, int timer_class);
void build_colouring(Grid& g);
void ensure_mc_pack(Grid& g);
// mmg_stream.cu: TMA-fed (cp.async.bulk + mbarrier ring) multicolour sweep and SpMV-class operators; false = no instantiation
bool stream_sor_mc(Grid& g);
struct PeerSends;
bool stream_sor_mc_flow(Grid& g, double* xs, size_t stride, const PeerSends* peers);
bool stream_spmv(const HybMatrix& M, const double* x, const double* b, double* y, const unsigned char* rowflag, int op, int mask_d, int mask_n,
                 double* partial, int* nblocks_out, int device, cudaStream_t s, int row0, int nrows);
void peer_setup(Solver& s);          // mmg_comm.cu: IPC exchange of the versioned-vector allocations of the partitioned levels
void peer_teardown(Solver& s);
void peer_init_sets(Grid& g, LevelDist& D);   // mmg_kernels.cu: both sets start as sentinels on the swept rows
void build_block_colouring(Grid& g);
void compute_lex_levels(Grid& g, std::vector<int>& level, int& n_levels);

// ---- assembly (assembly.cu) ---------------------------------------------------------------------
void asm_release(Grid& g);
void asm_knn_points(Grid& g, int m, const double* qx, const double* qy, const int* qflag, int neumann, int k, int* out_host);
void asm_rcm_order_points(Grid& g);
void asm_build_deriv_normal_bound(Grid& g);
void asm_build_laplacian(Grid& g);
void asm_weights(Grid& g, int which, int m, const int* ids, double* w, int* nb);
void asm_point_interp_weights(Grid& g, int m, const double* px, const double* py, int polyDeg, double* w, int* nb);
void asm_build_interp(Grid& base, Grid& target, int polyDeg, HybMatrix& out);
void asm_build_fs_operators(Grid& g);
// fractional-step explicit operators (kernels.cu)
void fs_calc_hat(Grid& g, int component);   // MMG_FS_U, MMG_FS_V or MMG_FS_BOTH
void fs_set_ppe_source(Grid& g);
void fs_correct(Grid& g, int component);
double fs_residual(Grid& g);

// ---- timers ---------------------------------------------------------------------------------------
struct TimedScope {
  Timers* t; int cls; cudaStream_t s; cudaEvent_t e0 = nullptr, e1 = nullptr;
  TimedScope(Grid& g, int cls_, int64_t bytes, int launches = 1);
  ~TimedScope();
};
void timers_collect(Timers& t);
void debug_lex_trace(long long* out, int n);
// diagnostics (mmg_debug_last_kernel): the instantiation the last smoother (slot 0) / SpMV-class (slot 1) dispatch launched
std::string& last_kernel_slot(int slot);
inline void note_kernel_slot(int slot, const char* name, int a = -1, int b = -1, int c = -1) {
  char buf[96];
  if (a < 0) snprintf(buf, sizeof buf, "%s", name);
  else if (b < 0) snprintf(buf, sizeof buf, "%s<%d>", name, a);
  else if (c < 0) snprintf(buf, sizeof buf, "%s<%d,%d>", name, a, b);
  else snprintf(buf, sizeof buf, "%s<%d,%d,%d>", name, a, b, c);
  last_kernel_slot(slot) = buf;
}
inline void note_kernel(Grid&, const char* name, int a = -1, int b = -1, int c = -1) { note_kernel_slot(0, name, a, b, c); }

// ---- multi-GPU (comm.cu, kernels.cu) ---------------------------------------------------------------
void partition_bounds(int n, int world, int* bounds);
void comm_unique_id(char* out128);
void comm_init(Solver& s, int rank, int world, const char* id128);
void comm_destroy(Solver& s);
void plan_build(ExchangePlan& P, int rank, int world, const std::vector<std::pair<int, int>>& need, const std::vector<int>& bounds);
void plan_execute(Solver& s, const ExchangePlan& P, double* vec);
void allgather_blocks(Solver& s, double* vec, const std::vector<int>& bounds);
void allreduce_sum(Solver& s, double* dev, int count);
void dist_setup(Solver& s);
void dist_residual_norm(Solver& s, int level, double* ratio_dev);
void dist_residual(Solver& s, int level);
void dist_sor(Solver& s, int level);
void dist_restrict(Solver& s, int level);
void dist_prolong_correct(Solver& s, int level);

}  // namespace mmg

