// extern "C" layer of libmmg: argument checking, host<->device coherence at the points the
// reference's drivers touch the data, and the V-cycle schedule (Multigrid::vCycle) as a stream of
// kernel launches with no host synchronisation inside a cycle.
#include <algorithm>
#include <cstring>

#include <cmath>

#include "mmg_internal.hpp"

namespace mmg {

static thread_local std::string g_last_error;

Grid::~Grid() {
  delete fs;
  asm_release(*this);
  if (own_stream && stream) cudaStreamDestroy(stream);
}
Solver::~Solver() {
  comm_destroy(*this);
  for (Grid* g : grids) delete g;
  for (HybMatrix* m : restrict_) delete m;
  for (HybMatrix* m : prolong_) delete m;
  for (cudaEvent_t e : timers.pool) cudaEventDestroy(e);
  if (stream) cudaStreamDestroy(stream);
}

static void use_device(int device) { MMG_CUDA(cudaSetDevice(device)); }

// device copies of the boundary lists / row flags follow the host boundaries_ and bcFlags_
static void refresh_boundary_state(Grid& g) {
  std::vector<int> dp, np;
  std::vector<double> dv, nv;
  for (const Boundary& b : g.boundaries) {
    if (b.type == MMG_BC_DIRICHLET) { dp.insert(dp.end(), b.pts.begin(), b.pts.end()); dv.insert(dv.end(), b.vals.begin(), b.vals.end()); }
    else if (b.type == MMG_BC_NEUMANN) { np.insert(np.end(), b.pts.begin(), b.pts.end()); nv.insert(nv.end(), b.vals.begin(), b.vals.end()); }
  }
  dv.resize(dp.size(), 0.0);
  nv.resize(np.size(), 0.0);
  g.dir_pts.upload(dp, g.stream); g.dir_vals.upload(dv, g.stream);
  g.neu_pts.upload(np, g.stream); g.neu_vals.upload(nv, g.stream);
  std::vector<unsigned char> rf(g.A, 0);
  for (int i = 0; i < g.n; i++) rf[i] = (unsigned char)g.bcflags[i];
  g.rowflag.upload(rf, g.stream);
  g.sync();
}

static Grid* grid_create(int device, int n, const double* x, const double* y, const mmg_props* props, const double* source, int source_len,
                         int nb, const int* b_type, const int* b_ptr, const int* b_points, const double* b_values) {
  MMG_REQUIRE(n > 0 && x && y && props && source, MMG_ERR_ARG, "mmg_grid_create: null or empty input");
  int ndev = 0;
  MMG_CUDA(cudaGetDeviceCount(&ndev));
  MMG_REQUIRE(device >= 0 && device < ndev, MMG_ERR_CUDA, "mmg_grid_create: CUDA device " + std::to_string(device) + " not available (no CPU fallback)");
  use_device(device);
  Grid* g = new Grid();
  try {
    g->device = device;
    g->n = n;
    g->props = *props;
    MMG_CUDA(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
    g->own_stream = true;
    g->timers = &g->own_timers;
    g->hx.assign(x, x + n); g->hy.assign(y, y + n);
    g->hnx.assign(n, 0.0); g->hny.assign(n, 0.0);
    g->boundaries.resize(nb);
    for (int i = 0; i < nb; i++) {
      g->boundaries[i].type = b_type[i];
      g->boundaries[i].pts.assign(b_points + b_ptr[i], b_points + b_ptr[i + 1]);
      if (b_values) g->boundaries[i].vals.assign(b_values + b_ptr[i], b_values + b_ptr[i + 1]);
      else g->boundaries[i].vals.assign(b_ptr[i + 1] - b_ptr[i], 0.0);
      for (int p : g->boundaries[i].pts) MMG_REQUIRE(p >= 0 && p < n, MMG_ERR_ARG, "boundary point id out of range");
      if (b_type[i] == MMG_BC_NEUMANN) g->neumann = true;   // setNeumannFlag(), grid.cpp:52-60
    }
    g->A = g->neumann ? n + 1 : n;                          // grid.cpp:17
    MMG_REQUIRE(source_len == g->A, MMG_ERR_ARG, "source must have " + std::to_string(g->A) + " entries (n, or n+1 for a Neumann grid)");
    g->bcflags.assign(n, 0);
    g->diags.assign(g->A, 0.0);
    g->x.alloc(g->A); g->x.zero(g->stream);                 // values_->setZero(), grid.cpp:22-23
    g->x_alt.alloc(g->A); g->x_alt.zero(g->stream);
    g->r.alloc(g->A);
    g->b.upload(source, g->A, g->stream);
    g->px.upload(g->hx, g->stream); g->py.upload(g->hy, g->stream);
    g->abort_flag.alloc(1); g->abort_flag.zero(g->stream);
    refresh_boundary_state(*g);
  } catch (...) { delete g; throw; }
  return g;
}

static void check_abort(Grid& g) {
  int flag = 0;
  g.abort_flag.download(&flag, 1, g.stream);
  if (flag) {
    g.abort_flag.zero(g.stream);
    throw Error(MMG_ERR_TIMEOUT, "lexicographic sweep watchdog fired (dependency wait exceeded its limit)");
  }
}

static void set_laplacian_csr(Grid& g, int rows, const int* ptr, const int* idx, const double* val, const double* diags, const int* nb_ptr,
                              const int* nb_idx, const double* nb_val) {
  MMG_REQUIRE(rows == g.A, MMG_ERR_ARG, "laplaceMat_ must have " + std::to_string(g.A) + " rows");
  HostCsr A;
  A.rows = rows; A.cols = rows;
  A.ptr.assign(ptr, ptr + rows + 1);
  A.idx.assign(idx, idx + ptr[rows]);
  A.val.assign(val, val + ptr[rows]);
  for (int c : A.idx) MMG_REQUIRE(c >= 0 && c < rows, MMG_ERR_ARG, "column index out of range");
  // the parallel Neumann evaluation relies on boundary rows touching interior nodes only (grid.cpp:236,244)
  for (int i = 0; i < g.n; i++)
    if (g.bcflags[i] == MMG_BC_NEUMANN)
      for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++)
        MMG_REQUIRE(A.idx[k] == i || (A.idx[k] < g.n && g.bcflags[A.idx[k]] == 0), MMG_ERR_ARG,
                    "a Neumann row couples to another boundary node; bound_eval_neumann would be order dependent");
  hyb_from_csr(g.Lap, A, /*diag_first=*/true, /*has_reg=*/g.neumann, g.stream);
  if (diags) g.diags.assign(diags, diags + g.A);
  g.nbc = HostCsr();
  if (nb_ptr) {
    g.nbc.rows = g.nbc.cols = rows;
    g.nbc.ptr.assign(nb_ptr, nb_ptr + rows + 1);
    g.nbc.idx.assign(nb_idx, nb_idx + nb_ptr[rows]);
    g.nbc.val.assign(nb_val, nb_val + nb_ptr[rows]);
  }
  g.have_laplacian = true; g.lex_neu_ok = -1;
  g.have_colours = false;
  g.have_blocks = false;
}

// Grid::push_inhomog_to_rhs grid.cpp:664-685 — O(|boundary| * stencil) entries; done through a host round trip of
// the touched source entries only when the Laplacian came from the upload path (device path: assembly.cu).
static void push_inhomog_to_rhs(Grid& g) {
  if (!g.implicit) return;
  if (g.nbc.rows == 0 || g.nbc.idx.empty()) return;
  std::vector<double> src(g.A);
  g.b.download(src.data(), g.A, g.stream);
  const std::vector<double> copy = src;
  for (int i = 0; i < g.n; i++) {
    if (g.bcflags[i] != 0) continue;
    for (int j = g.nbc.ptr[i]; j < g.nbc.ptr[i + 1]; j++) {
      const double diag = g.diags[g.nbc.idx[j]];
      const double A_ij = g.nbc.val[j];
      src[i] -= A_ij * copy[g.nbc.idx[j]] / diag;
    }
  }
  g.b.upload(src.data(), g.A, g.stream);
  g.sync();
}

// ---- Multigrid::vCycle multigrid.cpp:62-110 / FracStepMultigrid.cpp:60-112 ---------------------------
static void vcycle(Solver& s) {
  const size_t L = s.grids.size();
  MMG_REQUIRE(L >= 1, MMG_ERR_STATE, "vCycle: no grids");
  Grid* cur = s.grids[L - 1];
  if (s.flavour == MMG_FLAVOUR_FRACSTEP && L == 1) { op_sor(*cur, s.smoother); return; }  // FracStepMultigrid.cpp:64-67
  for (size_t i = 1; i < L; i++) MMG_REQUIRE(s.restrict_[i] && s.prolong_[i - 1], MMG_ERR_STATE, "vCycle: buildMatrices() has not run");
  if ((size_t)s.hist_len >= s.hist.n) {
    DevBuf<double> bigger;
    bigger.alloc(std::max<size_t>(1024, s.hist.n * 2));
    if (s.hist_len) MMG_CUDA(cudaMemcpyAsync(bigger.p, s.hist.p, sizeof(double) * s.hist_len, cudaMemcpyDeviceToDevice, s.stream));
    MMG_CUDA(cudaStreamSynchronize(s.stream));
    s.hist = std::move(bigger);
  }
  const bool dist = s.world > 1;
  if (dist && !s.dist_ready) dist_setup(s);
  if (dist) dist_residual_norm(s, (int)L - 1, s.hist.p + s.hist_len);
  else op_residual_norm(*cur, s.hist.p + s.hist_len);   // residuals_.push_back(residual()), :66-67
  s.hist_len++;
  op_bound_eval_neumann(*cur);                      // :68
  for (size_t i = L - 1; i > 0; i--) {              // :71-88
    cur = s.grids[i];
    Grid* coarse = s.grids[i - 1];
    if (i != L - 1) op_zero_values(*cur);
    op_boundary_op(*cur, i == L - 1 ? MMG_FINE : MMG_COARSE);
    if (dist) { dist_sor(s, (int)i); dist_residual(s, (int)i); dist_restrict(s, (int)i); }
    else {
      op_sor(*cur, s.smoother);
      op_residual(*cur, cur->r.p);
      op_restrict(*cur, *coarse, *s.restrict_[i], cur->r.p);
    }
  }
  op_boundary_op(*cur, MMG_COARSE);                 // quirk kept: still grid 1 (or the only grid), :91
  cur = s.grids[0];
  op_zero_values(*cur);
  if (dist) { dist_sor(s, 0); dist_sor(s, 0); } else { op_sor(*cur, s.smoother); op_sor(*cur, s.smoother); }
  for (size_t i = 1; i < L; i++) {                  // :99-109
    cur = s.grids[i];
    if (dist) { dist_prolong_correct(s, (int)i); dist_sor(s, (int)i); }
    else {
      op_prolong_correct(*cur, *s.grids[i - 1], *s.prolong_[i - 1]);
      op_sor(*cur, s.smoother);
    }
  }
}

static void solver_check_abort(Solver& s) {
  for (Grid* g : s.grids) check_abort(*g);
}

static void fetch_history(Solver& s) {
  s.hist_host.resize(s.hist_len);
  if (s.hist_len) s.hist.download(s.hist_host.data(), s.hist_len, s.stream);
}

}  // namespace mmg

using namespace mmg;

#define API_BEGIN try {
#define API_END                                    \
    return MMG_OK;                                 \
  } catch (const mmg::Error& e) {                  \
    mmg::g_last_error = e.what();                  \
    return e.code;                                 \
  } catch (const std::exception& e) {              \
    mmg::g_last_error = e.what();                  \
    return MMG_ERR_STATE;                          \
  }
#define G(ptr) (*reinterpret_cast<mmg::Grid*>(ptr))
#define S(ptr) (*reinterpret_cast<mmg::Solver*>(ptr))
#define NEED(p) MMG_REQUIRE((p) != nullptr, MMG_ERR_ARG, std::string(__func__) + ": null argument " #p)

extern "C" {

const char* mmg_last_error(void) { return mmg::g_last_error.c_str(); }
const char* mmg_build_info(void) { return "libmmg sm_100a (compute_100a) CUDA " MMG_STR(__CUDACC_VER_MAJOR__) "." MMG_STR(__CUDACC_VER_MINOR__) " — no CPU fallback"; }

int mmg_device_count(int* count) {
  API_BEGIN
  NEED(count);
  *count = 0;
  MMG_CUDA(cudaGetDeviceCount(count));
  API_END
}

int mmg_grid_create(mmg_grid** out, int device, int n, const double* x, const double* y, const mmg_props* props, const double* source,
                    int source_len, int n_boundaries, const int* b_type, const int* b_ptr, const int* b_points, const double* b_values) {
  API_BEGIN
  NEED(out);
  MMG_REQUIRE(n_boundaries == 0 || (b_type && b_ptr && b_points), MMG_ERR_ARG, "mmg_grid_create: boundary arrays missing");
  *out = reinterpret_cast<mmg_grid*>(grid_create(device, n, x, y, props, source, source_len, n_boundaries, b_type, b_ptr, b_points, b_values));
  API_END
}
int mmg_grid_destroy(mmg_grid* g) {
  API_BEGIN
  if (g) {
    MMG_REQUIRE(G(g).owner == nullptr, MMG_ERR_STATE, "mmg_grid_destroy: the grid is owned by a solver, which frees it (multigrid.cpp:10-16)");
    use_device(G(g).device);
    delete &G(g);
  }
  API_END
}
int mmg_grid_set_implicit(mmg_grid* g, int flag) {
  API_BEGIN
  NEED(g);
  G(g).implicit = flag != 0;
  API_END
}
int mmg_grid_set_bc_flag(mmg_grid* g, int boundary, int type, const double* values, int n_values) {
  API_BEGIN
  NEED(g);
  Grid& gr = G(g);
  use_device(gr.device);
  MMG_REQUIRE(boundary >= 0 && boundary < (int)gr.boundaries.size(), MMG_ERR_ARG, "setBCFlag: no such boundary");
  MMG_REQUIRE(type == MMG_BC_DIRICHLET || type == MMG_BC_NEUMANN, MMG_ERR_ARG, "setBCFlag: type must be dirichlet(1) or neumann(2)");
  Boundary& b = gr.boundaries[boundary];
  MMG_REQUIRE((type == MMG_BC_NEUMANN) == gr.neumann || !gr.neumann || type == MMG_BC_DIRICHLET, MMG_ERR_ARG,
              "setBCFlag: a Neumann type must already be present at construction (Grid ctor sizes the system from it)");
  MMG_REQUIRE(type != MMG_BC_NEUMANN || gr.neumann, MMG_ERR_ARG, "setBCFlag: grid was constructed without a Neumann boundary");
  b.type = type;
  for (int p : b.pts) gr.bcflags[p] = type;
  if (values) { MMG_REQUIRE(n_values == (int)b.pts.size(), MMG_ERR_ARG, "setBCFlag: value count differs from the boundary's point count"); b.vals.assign(values, values + n_values); }
  refresh_boundary_state(gr);
  API_END
}
int mmg_grid_build_normal_vecs_square(mmg_grid* g) {
  API_BEGIN
  NEED(g);
  Grid& gr = G(g);
  MMG_REQUIRE(!gr.boundaries.empty(), MMG_ERR_STATE, "build_normal_vecs: no boundary");
  for (int p : gr.boundaries[0].pts) {  // grid.cpp:445-461: boundaries_[0] only, y tested first, inward normals
    const double x = gr.hx[p], y = gr.hy[p];
    if (y == 0) { gr.hnx[p] = 0; gr.hny[p] = 1; }
    else if (y == 1) { gr.hnx[p] = 0; gr.hny[p] = -1; }
    else if (x == 0) { gr.hnx[p] = 1; gr.hny[p] = 0; }
    else if (x == 1) { gr.hnx[p] = -1; gr.hny[p] = 0; }
  }
  API_END
}
// Grid::build_normal_vecs(filename, geomtype) grid.cpp:442-516: the square loop over boundaries_[0] runs for every geomtype;
// "square_with_circle" then gives boundaries_[1] (the hole) outward radial normals about (0.5, 0.5), "concentric_circles" gives
// boundaries_[0] (outer circle) inward and boundaries_[1] (inner circle) outward radial normals.  O(|boundary|) host work.
int mmg_grid_build_normal_vecs(mmg_grid* g, int geomtype) {
  API_BEGIN
  NEED(g);
  MMG_REQUIRE(geomtype >= MMG_GEOM_SQUARE && geomtype <= MMG_GEOM_CONCENTRIC_CIRCLES, MMG_ERR_ARG, "build_normal_vecs: unknown geomtype");
  if (int rc = mmg_grid_build_normal_vecs_square(g)) return rc;
  Grid& gr = G(g);
  auto radial = [&](const Boundary& bd, bool outward) {
    for (int p : bd.pts) {
      double x = gr.hx[p], y = gr.hy[p];
      x = x - 0.5; y = y - 0.5;
      const double norm = std::sqrt(x * x + y * y);
      x /= norm; y /= norm;
      gr.hnx[p] = outward ? x : -x; gr.hny[p] = outward ? y : -y;
    }
  };
  if (geomtype != MMG_GEOM_SQUARE) MMG_REQUIRE(gr.boundaries.size() >= 2, MMG_ERR_STATE, "build_normal_vecs: the circle geometries have two boundaries");
  if (geomtype == MMG_GEOM_SQUARE_WITH_CIRCLE) radial(gr.boundaries[1], true);
  else if (geomtype == MMG_GEOM_CONCENTRIC_CIRCLES) { radial(gr.boundaries[0], false); radial(gr.boundaries[1], true); }
  API_END
}
int mmg_grid_set_normal_vecs(mmg_grid* g, const double* nx, const double* ny) {
  API_BEGIN
  NEED(g); NEED(nx); NEED(ny);
  G(g).hnx.assign(nx, nx + G(g).n); G(g).hny.assign(ny, ny + G(g).n);
  API_END
}
int mmg_grid_rcm_order_points(mmg_grid* g) {
  API_BEGIN
  NEED(g);
  use_device(G(g).device);
  asm_rcm_order_points(G(g));
  refresh_boundary_state(G(g));
  API_END
}
int mmg_grid_build_deriv_normal_bound(mmg_grid* g) {
  API_BEGIN
  NEED(g);
  use_device(G(g).device);
  asm_build_deriv_normal_bound(G(g));
  API_END
}
int mmg_grid_build_laplacian(mmg_grid* g) {
  API_BEGIN
  NEED(g);
  use_device(G(g).device);
  asm_build_laplacian(G(g));
  API_END
}
int mmg_grid_modify_coeff_neumann(mmg_grid* g, int coarse) {
  API_BEGIN
  NEED(g);
  use_device(G(g).device);
  op_modify_coeff_neumann(G(g), coarse);
  API_END
}
int mmg_grid_push_inhomog_to_rhs(mmg_grid* g) {
  API_BEGIN
  NEED(g);
  use_device(G(g).device);
  push_inhomog_to_rhs(G(g));
  API_END
}
int mmg_grid_boundary_op(mmg_grid* g, int coarse) {
  API_BEGIN
  NEED(g);
  use_device(G(g).device);
  op_boundary_op(G(g), coarse);
  API_END
}
int mmg_grid_bound_eval_neumann(mmg_grid* g) {
  API_BEGIN
  NEED(g);
  use_device(G(g).device);
  op_bound_eval_neumann(G(g));
  API_END
}
int mmg_grid_set_props(mmg_grid* g, const mmg_props* props) {
  API_BEGIN
  NEED(g); NEED(props);
  MMG_REQUIRE(props->polyDeg == G(g).props.polyDeg && props->stencilSize == G(g).props.stencilSize && props->rbfExp == G(g).props.rbfExp, MMG_ERR_ARG,
              "set_props: only omega and iters may change once the grid exists (the operators are built from the rest)");
  MMG_REQUIRE(props->iters >= 0, MMG_ERR_ARG, "set_props: iters must be >= 0");
  G(g).props = *props;
  API_END
}
int mmg_grid_set_arithmetic(mmg_grid* g, int arithmetic) {
  API_BEGIN
  NEED(g);
  MMG_REQUIRE(arithmetic == MMG_ARITH_REFERENCE_ORDER || arithmetic == MMG_ARITH_FAST, MMG_ERR_ARG, "unknown arithmetic mode");
  G(g).exact = arithmetic == MMG_ARITH_REFERENCE_ORDER;
  API_END
}
int mmg_grid_sor(mmg_grid* g, int smoother) {
  API_BEGIN
  NEED(g);
  use_device(G(g).device);
  op_sor(G(g), smoother);
  G(g).sync();
  check_abort(G(g));
  API_END
}
int mmg_grid_residual(mmg_grid* g, double* out) {
  API_BEGIN
  NEED(g); NEED(out);
  Grid& gr = G(g);
  use_device(gr.device);
  op_residual(gr, gr.r.p);
  gr.r.download(out, gr.A, gr.stream);
  API_END
}
int mmg_grid_fix_vector_bound_coarse(mmg_grid* g, double* vec) {
  API_BEGIN
  NEED(g); NEED(vec);
  Grid& gr = G(g);
  use_device(gr.device);
  DevBuf<double> t;
  t.upload(vec, gr.A, gr.stream);
  op_fix_vector_bound_coarse(gr, t.p);
  t.download(vec, gr.A, gr.stream);
  API_END
}
int mmg_grid_knn(mmg_grid* g, int m, const double* qx, const double* qy, const int* q_bcflag, int neumann, int k, int* out) {
  API_BEGIN
  NEED(g); NEED(qx); NEED(qy); NEED(out);
  use_device(G(g).device);
  asm_knn_points(G(g), m, qx, qy, q_bcflag, neumann, k, out);
  API_END
}
int mmg_grid_weights(mmg_grid* g, int which, int m, const int* ids, double* w, int* nb) {
  API_BEGIN
  NEED(g); NEED(ids); NEED(w); NEED(nb);
  use_device(G(g).device);
  asm_weights(G(g), which, m, ids, w, nb);
  API_END
}
int mmg_grid_point_interp_weights(mmg_grid* g, int m, const double* px, const double* py, int polyDeg, double* w, int* nb) {
  API_BEGIN
  NEED(g); NEED(px); NEED(py); NEED(w); NEED(nb);
  use_device(G(g).device);
  asm_point_interp_weights(G(g), m, px, py, polyDeg, w, nb);
  API_END
}
int mmg_grid_sizes(mmg_grid* g, int* n, int* a_size, int* neumann_flag) {
  API_BEGIN
  NEED(g);
  if (n) *n = G(g).n;
  if (a_size) *a_size = G(g).A;
  if (neumann_flag) *neumann_flag = G(g).neumann;
  API_END
}
int mmg_grid_get_values(mmg_grid* g, double* out) {
  API_BEGIN
  NEED(g); NEED(out);
  use_device(G(g).device);
  G(g).x.download(out, G(g).A, G(g).stream);
  API_END
}
int mmg_grid_set_values(mmg_grid* g, const double* in) {
  API_BEGIN
  NEED(g); NEED(in);
  use_device(G(g).device);
  MMG_CUDA(cudaMemcpyAsync(G(g).x.p, in, sizeof(double) * G(g).A, cudaMemcpyHostToDevice, G(g).stream));
  G(g).sync();
  API_END
}
// ranged variants: a rank of a partitioned problem moves only its row block (+ halo) of values_ / source_ across PCIe
int mmg_grid_get_values_range(mmg_grid* g, int offset, int count, double* out) {
  API_BEGIN
  NEED(g); NEED(out);
  MMG_REQUIRE(offset >= 0 && count >= 0 && offset + count <= G(g).A, MMG_ERR_ARG, "get_values_range: range outside values_");
  use_device(G(g).device);
  if (count) MMG_CUDA(cudaMemcpyAsync(out, G(g).x.p + offset, sizeof(double) * count, cudaMemcpyDeviceToHost, G(g).stream));
  G(g).sync();
  API_END
}
int mmg_grid_set_values_range(mmg_grid* g, int offset, int count, const double* in) {
  API_BEGIN
  NEED(g); NEED(in);
  MMG_REQUIRE(offset >= 0 && count >= 0 && offset + count <= G(g).A, MMG_ERR_ARG, "set_values_range: range outside values_");
  use_device(G(g).device);
  if (count) MMG_CUDA(cudaMemcpyAsync(G(g).x.p + offset, in, sizeof(double) * count, cudaMemcpyHostToDevice, G(g).stream));
  G(g).sync();
  API_END
}
int mmg_grid_set_source_range(mmg_grid* g, int offset, int count, const double* in) {
  API_BEGIN
  NEED(g); NEED(in);
  MMG_REQUIRE(offset >= 0 && count >= 0 && offset + count <= G(g).A, MMG_ERR_ARG, "set_source_range: range outside source_");
  use_device(G(g).device);
  if (count) MMG_CUDA(cudaMemcpyAsync(G(g).b.p + offset, in, sizeof(double) * count, cudaMemcpyHostToDevice, G(g).stream));
  G(g).sync();
  API_END
}
int mmg_grid_get_source(mmg_grid* g, double* out) {
  API_BEGIN
  NEED(g); NEED(out);
  use_device(G(g).device);
  G(g).b.download(out, G(g).A, G(g).stream);
  API_END
}
int mmg_grid_set_source(mmg_grid* g, const double* in) {
  API_BEGIN
  NEED(g); NEED(in);
  use_device(G(g).device);
  MMG_CUDA(cudaMemcpyAsync(G(g).b.p, in, sizeof(double) * G(g).A, cudaMemcpyHostToDevice, G(g).stream));
  G(g).sync();
  API_END
}
int mmg_grid_get_points(mmg_grid* g, double* x, double* y) {
  API_BEGIN
  NEED(g); NEED(x); NEED(y);
  std::memcpy(x, G(g).hx.data(), sizeof(double) * G(g).n);
  std::memcpy(y, G(g).hy.data(), sizeof(double) * G(g).n);
  API_END
}
int mmg_grid_get_bcflags(mmg_grid* g, int* flags) {
  API_BEGIN
  NEED(g); NEED(flags);
  std::memcpy(flags, G(g).bcflags.data(), sizeof(int) * G(g).n);
  API_END
}
int mmg_grid_get_normals(mmg_grid* g, double* nx, double* ny) {
  API_BEGIN
  NEED(g); NEED(nx); NEED(ny);
  std::memcpy(nx, G(g).hnx.data(), sizeof(double) * G(g).n);
  std::memcpy(ny, G(g).hny.data(), sizeof(double) * G(g).n);
  API_END
}
int mmg_grid_get_diags(mmg_grid* g, double* out) {
  API_BEGIN
  NEED(g); NEED(out);
  std::memcpy(out, G(g).diags.data(), sizeof(double) * G(g).A);
  API_END
}
int mmg_grid_get_perm(mmg_grid* g, int* order) {
  API_BEGIN
  NEED(g); NEED(order);
  MMG_REQUIRE((int)G(g).order.size() == G(g).n, MMG_ERR_STATE, "rcm_order_points has not run on this grid");
  std::memcpy(order, G(g).order.data(), sizeof(int) * G(g).n);
  API_END
}
int mmg_grid_get_boundary(mmg_grid* g, int boundary, int* type, int* count, int* points, double* values) {
  API_BEGIN
  NEED(g);
  MMG_REQUIRE(boundary >= 0 && boundary < (int)G(g).boundaries.size(), MMG_ERR_ARG, "no such boundary");
  const Boundary& b = G(g).boundaries[boundary];
  if (type) *type = b.type;
  if (count) *count = (int)b.pts.size();
  if (points) std::memcpy(points, b.pts.data(), sizeof(int) * b.pts.size());
  if (values) std::memcpy(values, b.vals.data(), sizeof(double) * b.vals.size());
  API_END
}
static void grid_csr(Grid& gr, int which, HostCsr& A) {
  if (which == MMG_MAT_DERIVX || which == MMG_MAT_DERIVY || which == MMG_MAT_UVLAPLACE) {
    MMG_REQUIRE(gr.fs && gr.fs->have_ops, MMG_ERR_STATE, "the fractional-step operators have not been built or uploaded");
    hyb_to_csr(which == MMG_MAT_DERIVX ? gr.fs->Dx : which == MMG_MAT_DERIVY ? gr.fs->Dy : gr.fs->Lap, A, gr.stream);
    return;
  }
  if (which == MMG_MAT_LAPLACE) {
    MMG_REQUIRE(gr.have_laplacian, MMG_ERR_STATE, "laplaceMat_ has not been built or uploaded");
    hyb_to_csr(gr.Lap, A, gr.stream);
  } else if (which == MMG_MAT_NEUMANN_COEFFS) {
    A = gr.nbc;
    if (A.rows == 0) { A.rows = A.cols = gr.A; A.ptr.assign(gr.A + 1, 0); }
  } else {
    throw Error(MMG_ERR_ARG, "matrix selector not available on a grid");
  }
}
int mmg_grid_csr_nnz(mmg_grid* g, int which, int64_t* nnz) {
  API_BEGIN
  NEED(g); NEED(nnz);
  use_device(G(g).device);
  if (which == MMG_MAT_LAPLACE) { MMG_REQUIRE(G(g).have_laplacian, MMG_ERR_STATE, "laplaceMat_ has not been built or uploaded"); *nnz = G(g).Lap.nnz; }
  else { HostCsr A; grid_csr(G(g), which, A); *nnz = A.nnz(); }
  API_END
}
int mmg_grid_get_csr(mmg_grid* g, int which, int* ptr, int* idx, double* val) {
  API_BEGIN
  NEED(g); NEED(ptr); NEED(idx); NEED(val);
  use_device(G(g).device);
  HostCsr A;
  grid_csr(G(g), which, A);
  std::memcpy(ptr, A.ptr.data(), sizeof(int) * A.ptr.size());
  std::memcpy(idx, A.idx.data(), sizeof(int) * A.idx.size());
  std::memcpy(val, A.val.data(), sizeof(double) * A.val.size());
  API_END
}
// y = M x on the device for one of the grid's stencil matrices (host vectors in and out): the reference's drivers multiply
// derivXMat_ / derivYMat_ / uvLaplaceMat_ by host vectors in check_derivs (FractionalStepSim.cpp:80-113); for laplaceMat_ the
// dense regularisation row is included.  x has n entries for the derivative operators, A for laplaceMat_.
int mmg_grid_apply_matrix(mmg_grid* g, int which, const double* x, double* y) {
  API_BEGIN
  NEED(g); NEED(x); NEED(y);
  Grid& gr = G(g);
  use_device(gr.device);
  const HybMatrix* M = nullptr;
  if (which == MMG_MAT_LAPLACE) { MMG_REQUIRE(gr.have_laplacian, MMG_ERR_STATE, "laplaceMat_ has not been built"); M = &gr.Lap; }
  else {
    MMG_REQUIRE(gr.fs && gr.fs->have_ops, MMG_ERR_STATE, "apply_matrix: the fractional-step operators have not been built");
    M = which == MMG_MAT_DERIVX ? &gr.fs->Dx : which == MMG_MAT_DERIVY ? &gr.fs->Dy : which == MMG_MAT_UVLAPLACE ? &gr.fs->Lap : nullptr;
  }
  MMG_REQUIRE(M != nullptr, MMG_ERR_ARG, "apply_matrix: which must be MMG_MAT_LAPLACE / DERIVX / DERIVY / UVLAPLACE");
  const int rows = M->rows + (M->reg_row >= 0 ? 1 : 0);
  DevBuf<double> dx, dy;
  dx.upload(x, (size_t)M->cols, gr.stream);
  dy.alloc((size_t)rows);
  op_spmv(*M, dx.p, dy.p, gr, MMG_T_OTHER);
  if (M->reg_row >= 0) {   // dense last row: in-order on the host side of the call is not needed -- a debug product, tree sums are fine
    std::vector<int> rc = M->reg_col.to_host(gr.stream);
    std::vector<double> rv = M->reg_val.to_host(gr.stream);
    double s = M->reg_diag * x[M->reg_row];
    for (size_t k = 0; k < rc.size(); k++) s += rv[k] * x[rc[k]];
    dy.download(y, (size_t)M->rows, gr.stream);
    y[M->reg_row] = s;
  } else {
    dy.download(y, (size_t)rows, gr.stream);
  }
  API_END
}
int mmg_grid_set_laplacian_csr(mmg_grid* g, int rows, const int* ptr, const int* idx, const double* val, const double* diags, const int* nb_ptr,
                               const int* nb_idx, const double* nb_val) {
  API_BEGIN
  NEED(g); NEED(ptr); NEED(idx); NEED(val);
  use_device(G(g).device);
  set_laplacian_csr(G(g), rows, ptr, idx, val, diags, nb_ptr, nb_idx, nb_val);
  API_END
}
int mmg_grid_get_colouring(mmg_grid* g, int* n_colours, int* colour) {
  API_BEGIN
  NEED(g); NEED(n_colours); NEED(colour);
  Grid& gr = G(g);
  use_device(gr.device);
  MMG_REQUIRE(gr.have_laplacian, MMG_ERR_STATE, "laplaceMat_ has not been built or uploaded");
  if (!gr.have_colours) build_colouring(gr);
  *n_colours = gr.n_colours;
  std::memcpy(colour, gr.colour_host.data(), sizeof(int) * gr.A);
  API_END
}
int mmg_grid_get_colour_counts(mmg_grid* g, int* n_colours, int* counts, int cap) {
  API_BEGIN
  NEED(g); NEED(n_colours);
  Grid& gr = G(g);
  use_device(gr.device);
  MMG_REQUIRE(gr.have_laplacian, MMG_ERR_STATE, "laplaceMat_ has not been built or uploaded");
  if (!gr.have_colours) build_colouring(gr);
  *n_colours = gr.n_colours;
  if (counts) for (int c = 0; c < gr.n_colours && c < cap; c++) counts[c] = gr.colour_ptr[c + 1] - gr.colour_ptr[c];
  API_END
}
int mmg_grid_set_block_size(mmg_grid* g, int rows_per_block) {
  API_BEGIN
  NEED(g);
  MMG_REQUIRE(rows_per_block >= 32, MMG_ERR_ARG, "block size must be at least 32 rows");
  if (G(g).block_size != rows_per_block) { G(g).block_size = rows_per_block; G(g).have_blocks = false; }
  API_END
}
int mmg_grid_get_block_colouring(mmg_grid* g, int* n_blocks, int* n_colours, int* colour, int cap) {
  API_BEGIN
  NEED(g); NEED(n_blocks); NEED(n_colours);
  Grid& gr = G(g);
  use_device(gr.device);
  MMG_REQUIRE(gr.have_laplacian, MMG_ERR_STATE, "laplaceMat_ has not been built or uploaded");
  if (!gr.have_blocks) build_block_colouring(gr);
  *n_blocks = (int)gr.blk_colour.size();
  *n_colours = gr.n_blk_colours;
  if (colour) for (int i = 0; i < *n_blocks && i < cap; i++) colour[i] = gr.blk_colour[i];
  API_END
}
int mmg_grid_get_lex_levels(mmg_grid* g, int* n_levels, int* level) {
  API_BEGIN
  NEED(g); NEED(n_levels); NEED(level);
  Grid& gr = G(g);
  use_device(gr.device);
  MMG_REQUIRE(gr.have_laplacian, MMG_ERR_STATE, "laplaceMat_ has not been built or uploaded");
  std::vector<int> lv;
  compute_lex_levels(gr, lv, *n_levels);
  std::memcpy(level, lv.data(), sizeof(int) * gr.A);
  API_END
}

// ------------------------------------------------------------------------------------------------ fractional step
static void fs_refresh_boundary(Grid& g) {
  std::vector<int> pts;
  for (const Boundary& b : g.boundaries) pts.insert(pts.end(), b.pts.begin(), b.pts.end());
  g.fs->bnd_pts.upload(pts, g.stream);
  g.fs->nx.upload(g.hnx, g.stream); g.fs->ny.upload(g.hny, g.stream);
  g.sync();
}
int mmg_grid_fs_init(mmg_grid* g, double dt, double mu, double rho) {
  API_BEGIN
  NEED(g);
  Grid& gr = G(g);
  use_device(gr.device);
  if (!gr.fs) {
    gr.fs = new Grid::FracStep();
    for (int i = 0; i < 6; i++) { gr.fs->vec[i].alloc(gr.n); gr.fs->vec[i].zero(gr.stream); }   // fractionalStepGrid.cpp:5-16
    gr.fs->t0.alloc(gr.n); gr.fs->t1.alloc(gr.n); gr.fs->t2.alloc(gr.n);
  }
  gr.fs->dt = dt; gr.fs->mu = mu; gr.fs->rho = rho;
  fs_refresh_boundary(gr);
  API_END
}
int mmg_grid_fs_build_operators(mmg_grid* g) {
  API_BEGIN
  NEED(g);
  use_device(G(g).device);
  asm_build_fs_operators(G(g));
  fs_refresh_boundary(G(g));
  API_END
}
int mmg_grid_fs_set_operator_csr(mmg_grid* g, int which, int rows, const int* ptr, const int* idx, const double* val) {
  API_BEGIN
  NEED(g); NEED(ptr); NEED(idx); NEED(val);
  Grid& gr = G(g);
  use_device(gr.device);
  MMG_REQUIRE(gr.fs != nullptr, MMG_ERR_STATE, "mmg_grid_fs_init has not been called");
  MMG_REQUIRE(which == MMG_MAT_DERIVX || which == MMG_MAT_DERIVY || which == MMG_MAT_UVLAPLACE, MMG_ERR_ARG, "which must be MMG_MAT_DERIVX, _DERIVY or _UVLAPLACE");
  MMG_REQUIRE(rows == gr.n, MMG_ERR_ARG, "the fractional-step operators are N x N");
  HostCsr A;
  A.rows = rows; A.cols = rows;
  A.ptr.assign(ptr, ptr + rows + 1); A.idx.assign(idx, idx + ptr[rows]); A.val.assign(val, val + ptr[rows]);
  for (int c : A.idx) MMG_REQUIRE(c >= 0 && c < rows, MMG_ERR_ARG, "column index out of range");
  HybMatrix& M = which == MMG_MAT_DERIVX ? gr.fs->Dx : which == MMG_MAT_DERIVY ? gr.fs->Dy : gr.fs->Lap;
  hyb_from_csr(M, A, false, false, gr.stream);
  gr.fs->have_ops = gr.fs->Dx.rows == gr.n && gr.fs->Dy.rows == gr.n && gr.fs->Lap.rows == gr.n;
  fs_refresh_boundary(gr);
  API_END
}
static DevBuf<double>& fs_vec(Grid& gr, int which) {
  MMG_REQUIRE(gr.fs != nullptr, MMG_ERR_STATE, "mmg_grid_fs_init has not been called");
  MMG_REQUIRE(which >= 0 && which < 6, MMG_ERR_ARG, "vector selector must be MMG_FS_U .. MMG_FS_V_HAT");
  return gr.fs->vec[which];
}
int mmg_grid_fs_get_vec(mmg_grid* g, int which, double* out) {
  API_BEGIN
  NEED(g); NEED(out);
  use_device(G(g).device);
  fs_vec(G(g), which).download(out, G(g).n, G(g).stream);
  API_END
}
int mmg_grid_fs_set_vec(mmg_grid* g, int which, const double* in) {
  API_BEGIN
  NEED(g); NEED(in);
  use_device(G(g).device);
  MMG_CUDA(cudaMemcpyAsync(fs_vec(G(g), which).p, in, sizeof(double) * G(g).n, cudaMemcpyHostToDevice, G(g).stream));
  G(g).sync();
  API_END
}
int mmg_grid_fs_scatter(mmg_grid* g, int which, int count, const int* idx, const double* vals) {
  API_BEGIN
  NEED(g); NEED(idx); NEED(vals);
  Grid& gr = G(g);
  use_device(gr.device);
  DevBuf<double>& v = fs_vec(gr, which);
  for (int i = 0; i < count; i++) MMG_REQUIRE(idx[i] >= 0 && idx[i] < gr.n, MMG_ERR_ARG, "scatter index out of range");
  for (int i = 0; i < count; i++) MMG_CUDA(cudaMemcpyAsync(v.p + idx[i], vals + i, sizeof(double), cudaMemcpyHostToDevice, gr.stream));
  gr.sync();
  API_END
}
__global__ void k_fs_scatter_uv(int count, const int* __restrict__ idx, const double* __restrict__ uv, const double* __restrict__ vv,
                                double* u, double* v, double* uo, double* vo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int c = idx[i];
  u[c] = uv[i]; uo[c] = uv[i];
  v[c] = vv[i]; vo[c] = vv[i];
}
static void fs_scatter_pairs(Grid& gr, const std::vector<int>& idx, const std::vector<double>& uval, const std::vector<double>& vval) {
  const int count = (int)idx.size();
  if (count == 0) return;
  DevBuf<int> di; DevBuf<double> du, dv;
  di.alloc(count); du.alloc(count); dv.alloc(count);
  di.upload(idx.data(), count, gr.stream); du.upload(uval.data(), count, gr.stream); dv.upload(vval.data(), count, gr.stream);
  k_fs_scatter_uv<<<(count + 255) / 256, 256, 0, gr.stream>>>(count, di.p, du.p, dv.p, gr.fs->vec[MMG_FS_U].p, gr.fs->vec[MMG_FS_V].p,
                                                              gr.fs->vec[MMG_FS_U_OLD].p, gr.fs->vec[MMG_FS_V_OLD].p);
  MMG_CUDA(cudaGetLastError());
  gr.sync();
}
// fractionalStepGrid.cpp:41-59: the Kovasznay velocities on every boundary node, evaluated on the host with the same libm
// calls the reference makes (std::exp / std::cos / std::sin on doubles), written into u, v, u_old and v_old.
int mmg_grid_fs_set_uv_bound(mmg_grid* g) {
  API_BEGIN
  NEED(g);
  Grid& gr = G(g);
  MMG_REQUIRE(gr.fs != nullptr, MMG_ERR_STATE, "mmg_grid_fs_init has not been called");
  use_device(gr.device);
  const double pi = 3.141592653589793238462643383279502884;
  const double re = gr.fs->rho / gr.fs->mu;
  const double lambda = 0.5 * re - std::sqrt(0.25 * re * re + 4 * pi * pi);
  std::vector<int> idx;
  std::vector<double> uval, vval;
  for (const Boundary& b : gr.boundaries)
    for (int c : b.pts) {
      const double x = gr.hx[c], y = gr.hy[c];
      idx.push_back(c);
      uval.push_back(1 - std::exp(lambda * x) * std::cos(2 * pi * y));
      vval.push_back(lambda / (2 * pi) * std::exp(lambda * x) * std::sin(2 * pi * y));
    }
  fs_scatter_pairs(gr, idx, uval, vval);
  API_END
}
int mmg_grid_fs_calc_hat(mmg_grid* g, int component) {
  API_BEGIN
  NEED(g);
  MMG_REQUIRE(G(g).fs != nullptr, MMG_ERR_STATE, "mmg_grid_fs_init has not been called");
  use_device(G(g).device);
  MMG_REQUIRE(component == MMG_FS_BOTH || component == MMG_FS_U || component == MMG_FS_V, MMG_ERR_ARG, "component must be MMG_FS_U, MMG_FS_V or MMG_FS_BOTH");
  fs_calc_hat(G(g), component);
  API_END
}
int mmg_grid_fs_set_ppe_source(mmg_grid* g) {
  API_BEGIN
  NEED(g);
  MMG_REQUIRE(G(g).fs != nullptr, MMG_ERR_STATE, "mmg_grid_fs_init has not been called");
  use_device(G(g).device);
  fs_set_ppe_source(G(g));
  API_END
}
int mmg_grid_fs_correct(mmg_grid* g, int component) {
  API_BEGIN
  NEED(g);
  MMG_REQUIRE(G(g).fs != nullptr, MMG_ERR_STATE, "mmg_grid_fs_init has not been called");
  use_device(G(g).device);
  MMG_REQUIRE(component == MMG_FS_BOTH || component == MMG_FS_U || component == MMG_FS_V, MMG_ERR_ARG, "component must be MMG_FS_U, MMG_FS_V or MMG_FS_BOTH");
  fs_correct(G(g), component);
  API_END
}
int mmg_grid_fs_residual(mmg_grid* g, double* out) {
  API_BEGIN
  NEED(g); NEED(out);
  MMG_REQUIRE(G(g).fs != nullptr, MMG_ERR_STATE, "mmg_grid_fs_init has not been called");
  use_device(G(g).device);
  *out = fs_residual(G(g));
  API_END
}

// ------------------------------------------------------------------------------------------------ solver
int mmg_solver_create(mmg_solver** out, int flavour) {
  API_BEGIN
  NEED(out);
  MMG_REQUIRE(flavour == MMG_FLAVOUR_MULTIGRID || flavour == MMG_FLAVOUR_FRACSTEP, MMG_ERR_ARG, "unknown solver flavour");
  Solver* s = new Solver();
  s->flavour = flavour;
  *out = reinterpret_cast<mmg_solver*>(s);
  API_END
}
int mmg_solver_destroy(mmg_solver* s) {
  API_BEGIN
  if (s) {
    if (!S(s).grids.empty()) use_device(S(s).grids[0]->device);
    delete &S(s);
  }
  API_END
}
int mmg_solver_add_grid(mmg_solver* s, mmg_grid* g) {
  API_BEGIN
  NEED(s); NEED(g);
  Solver& so = S(s);
  Grid* gr = &G(g);
  MMG_REQUIRE(gr->owner == nullptr, MMG_ERR_STATE, "addGrid: the grid already belongs to a solver (Multigrid owns its grids, multigrid.cpp:10-16)");
  gr->owner = &so;
  use_device(gr->device);
  if (!so.stream) MMG_CUDA(cudaStreamCreateWithFlags(&so.stream, cudaStreamNonBlocking));
  gr->sync();
  if (gr->own_stream) { cudaStreamDestroy(gr->stream); gr->own_stream = false; }
  gr->stream = so.stream;               // one stream for the whole cycle
  gr->exact = so.arithmetic == MMG_ARITH_REFERENCE_ORDER;
  gr->timers = &so.timers;
  so.grids.push_back(gr);
  std::sort(so.grids.begin(), so.grids.end(), [](Grid* a, Grid* b) { return a->n != b->n ? a->n < b->n : a < b; });  // multigrid.cpp:116-122
  for (size_t i = 0; i < so.grids.size(); i++) so.grids[i]->level = (int)i;
  for (HybMatrix* m : so.restrict_) delete m;
  for (HybMatrix* m : so.prolong_) delete m;
  so.restrict_.assign(so.grids.size(), nullptr);
  so.prolong_.assign(so.grids.size(), nullptr);
  API_END
}
int mmg_solver_num_grids(mmg_solver* s, int* n) {
  API_BEGIN
  NEED(s); NEED(n);
  *n = (int)S(s).grids.size();
  API_END
}
int mmg_solver_grid(mmg_solver* s, int level, mmg_grid** g) {
  API_BEGIN
  NEED(s); NEED(g);
  MMG_REQUIRE(level >= 0 && level < (int)S(s).grids.size(), MMG_ERR_ARG, "no such level");
  *g = reinterpret_cast<mmg_grid*>(S(s).grids[level]);
  API_END
}
int mmg_solver_finish_build(mmg_solver* s) {
  API_BEGIN
  NEED(s);
  Solver& so = S(s);
  for (size_t i = 0; i + 1 < so.grids.size(); i++) { use_device(so.grids[i]->device); op_modify_coeff_neumann(*so.grids[i], MMG_COARSE); }  // multigrid.cpp:54-59
  API_END
}
int mmg_solver_build_matrices(mmg_solver* s) {
  API_BEGIN
  NEED(s);
  Solver& so = S(s);
  const size_t L = so.grids.size();
  MMG_REQUIRE(L >= 1, MMG_ERR_STATE, "buildMatrices: no grids");
  use_device(so.grids[0]->device);
  const int finepoly = so.grids[L - 1]->props.polyDeg;
  for (size_t i = 0; i + 1 < L; i++) {   // buildProlongMatrices multigrid.cpp:34-40: base = level i, target = level i+1
    delete so.prolong_[i];
    so.prolong_[i] = new HybMatrix();
    const int poly = so.flavour == MMG_FLAVOUR_FRACSTEP ? so.grids[i]->props.polyDeg : finepoly;  // FracStepMultigrid.cpp:23 vs multigrid.cpp:22,25
    asm_build_interp(*so.grids[i], *so.grids[i + 1], poly, *so.prolong_[i]);
  }
  for (size_t i = 1; i < L; i++) {       // buildRestrictionMatrices multigrid.cpp:41-47: base = level i, target = level i-1
    delete so.restrict_[i];
    so.restrict_[i] = new HybMatrix();
    const int poly = so.flavour == MMG_FLAVOUR_FRACSTEP ? so.grids[i]->props.polyDeg : finepoly;
    asm_build_interp(*so.grids[i], *so.grids[i - 1], poly, *so.restrict_[i]);
  }
  for (size_t i = 0; i + 1 < L; i++) op_modify_coeff_neumann(*so.grids[i], MMG_COARSE);
  API_END
}
int mmg_solver_set_interp_csr(mmg_solver* s, int which, int level, int rows, int cols, const int* ptr, const int* idx, const double* val) {
  API_BEGIN
  NEED(s); NEED(ptr); NEED(idx); NEED(val);
  Solver& so = S(s);
  MMG_REQUIRE(level >= 0 && level < (int)so.grids.size(), MMG_ERR_ARG, "no such level");
  MMG_REQUIRE(which == MMG_MAT_RESTRICT || which == MMG_MAT_PROLONG, MMG_ERR_ARG, "which must be MMG_MAT_RESTRICT or MMG_MAT_PROLONG");
  use_device(so.grids[level]->device);
  HostCsr A;
  A.rows = rows; A.cols = cols;
  A.ptr.assign(ptr, ptr + rows + 1);
  A.idx.assign(idx, idx + ptr[rows]);
  A.val.assign(val, val + ptr[rows]);
  for (int c : A.idx) MMG_REQUIRE(c >= 0 && c < cols, MMG_ERR_ARG, "column index out of range");
  std::vector<HybMatrix*>& dst = which == MMG_MAT_RESTRICT ? so.restrict_ : so.prolong_;
  delete dst[level];
  dst[level] = new HybMatrix();
  hyb_from_csr(*dst[level], A, false, false, so.stream);
  API_END
}
static HybMatrix* interp_of(Solver& so, int which, int level) {
  MMG_REQUIRE(level >= 0 && level < (int)so.grids.size(), MMG_ERR_ARG, "no such level");
  MMG_REQUIRE(which == MMG_MAT_RESTRICT || which == MMG_MAT_PROLONG, MMG_ERR_ARG, "which must be MMG_MAT_RESTRICT or MMG_MAT_PROLONG");
  HybMatrix* m = which == MMG_MAT_RESTRICT ? so.restrict_[level] : so.prolong_[level];
  MMG_REQUIRE(m != nullptr, MMG_ERR_STATE, "that interpolation matrix does not exist (NULL in the reference too) or is not built");
  return m;
}
int mmg_solver_interp_nnz(mmg_solver* s, int which, int level, int* rows, int* cols, int64_t* nnz) {
  API_BEGIN
  NEED(s);
  HybMatrix* m = interp_of(S(s), which, level);
  if (rows) *rows = m->rows;
  if (cols) *cols = m->cols;
  if (nnz) *nnz = m->nnz;
  API_END
}
int mmg_solver_get_interp_csr(mmg_solver* s, int which, int level, int* ptr, int* idx, double* val) {
  API_BEGIN
  NEED(s); NEED(ptr); NEED(idx); NEED(val);
  Solver& so = S(s);
  HybMatrix* m = interp_of(so, which, level);
  use_device(so.grids[level]->device);
  HostCsr A;
  hyb_to_csr(*m, A, so.stream);
  std::memcpy(ptr, A.ptr.data(), sizeof(int) * A.ptr.size());
  std::memcpy(idx, A.idx.data(), sizeof(int) * A.idx.size());
  std::memcpy(val, A.val.data(), sizeof(double) * A.val.size());
  API_END
}
int mmg_solver_set_smoother(mmg_solver* s, int smoother) {
  API_BEGIN
  NEED(s);
  MMG_REQUIRE(smoother == MMG_SMOOTHER_LEXICOGRAPHIC || smoother == MMG_SMOOTHER_MULTICOLOUR || smoother == MMG_SMOOTHER_BLOCK_LEXICOGRAPHIC, MMG_ERR_ARG,
              "unknown smoother");
  S(s).smoother = smoother;
  API_END
}
int mmg_solver_restrict(mmg_solver* s, int level) {
  API_BEGIN
  NEED(s);
  Solver& so = S(s);
  MMG_REQUIRE(level >= 1 && level < (int)so.grids.size() && so.restrict_[level], MMG_ERR_ARG, "restrict: level must be >= 1 with a built restriction matrix");
  Grid& fine = *so.grids[level];
  use_device(fine.device);
  op_residual(fine, fine.r.p);
  op_restrict(fine, *so.grids[level - 1], *so.restrict_[level], fine.r.p);
  MMG_CUDA(cudaStreamSynchronize(so.stream));
  API_END
}
int mmg_solver_prolong_correct(mmg_solver* s, int level) {
  API_BEGIN
  NEED(s);
  Solver& so = S(s);
  MMG_REQUIRE(level >= 1 && level < (int)so.grids.size() && so.prolong_[level - 1], MMG_ERR_ARG, "prolong: level must be >= 1 with a built prolongation matrix");
  use_device(so.grids[level]->device);
  op_prolong_correct(*so.grids[level], *so.grids[level - 1], *so.prolong_[level - 1]);
  MMG_CUDA(cudaStreamSynchronize(so.stream));
  API_END
}
int mmg_solver_coarse_solve(mmg_solver* s) {
  API_BEGIN
  NEED(s);
  Solver& so = S(s);
  MMG_REQUIRE(!so.grids.empty(), MMG_ERR_STATE, "coarse_solve: no grids");
  Grid& c = *so.grids[0];
  use_device(c.device);
  op_zero_values(c);
  op_sor(c, so.smoother);
  op_sor(c, so.smoother);
  MMG_CUDA(cudaStreamSynchronize(so.stream));
  solver_check_abort(so);
  API_END
}
int mmg_solver_set_block_size(mmg_solver* s, int rows_per_block) {
  API_BEGIN
  NEED(s);
  MMG_REQUIRE(rows_per_block >= 32, MMG_ERR_ARG, "block size must be at least 32 rows");
  for (Grid* g : S(s).grids)
    if (g->block_size != rows_per_block) { g->block_size = rows_per_block; g->have_blocks = false; }
  API_END
}
int mmg_solver_set_omega(mmg_solver* s, double omega) {
  API_BEGIN
  NEED(s);
  for (Grid* g : S(s).grids) g->props.omega = omega;
  API_END
}
int mmg_solver_set_arithmetic(mmg_solver* s, int arithmetic) {
  API_BEGIN
  NEED(s);
  MMG_REQUIRE(arithmetic == MMG_ARITH_REFERENCE_ORDER || arithmetic == MMG_ARITH_FAST, MMG_ERR_ARG, "unknown arithmetic mode");
  S(s).arithmetic = arithmetic;
  for (Grid* g : S(s).grids) g->exact = arithmetic == MMG_ARITH_REFERENCE_ORDER;
  API_END
}
int mmg_solver_vcycle(mmg_solver* s, int n_cycles) {
  API_BEGIN
  NEED(s);
  Solver& so = S(s);
  MMG_REQUIRE(!so.grids.empty(), MMG_ERR_STATE, "vCycle: no grids");
  use_device(so.grids[0]->device);
  for (int i = 0; i < n_cycles; i++) vcycle(so);
  MMG_CUDA(cudaStreamSynchronize(so.stream));
  solver_check_abort(so);
  API_END
}
int mmg_solver_time_vcycles(mmg_solver* s, int n_cycles, double* ms) {
  API_BEGIN
  NEED(s); NEED(ms);
  Solver& so = S(s);
  MMG_REQUIRE(!so.grids.empty(), MMG_ERR_STATE, "vCycle: no grids");
  use_device(so.grids[0]->device);
  cudaEvent_t e0, e1;
  MMG_CUDA(cudaEventCreate(&e0)); MMG_CUDA(cudaEventCreate(&e1));
  MMG_CUDA(cudaStreamSynchronize(so.stream));
  MMG_CUDA(cudaEventRecord(e0, so.stream));
  for (int i = 0; i < n_cycles; i++) vcycle(so);
  MMG_CUDA(cudaEventRecord(e1, so.stream));
  MMG_CUDA(cudaEventSynchronize(e1));
  float t = 0;
  MMG_CUDA(cudaEventElapsedTime(&t, e0, e1));
  *ms = t;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  solver_check_abort(so);
  API_END
}
int mmg_solver_residual(mmg_solver* s, double* out) {
  API_BEGIN
  NEED(s); NEED(out);
  Solver& so = S(s);
  MMG_REQUIRE(!so.grids.empty(), MMG_ERR_STATE, "residual: no grids");
  Grid& fine = *so.grids.back();
  use_device(fine.device);
  DevBuf<double> ratio;
  ratio.alloc(1);
  if (so.world > 1) { if (!so.dist_ready) dist_setup(so); dist_residual_norm(so, (int)so.grids.size() - 1, ratio.p); }
  else op_residual_norm(fine, ratio.p);
  ratio.download(out, 1, so.stream);
  API_END
}
int mmg_solver_history_len(mmg_solver* s, int* n) {
  API_BEGIN
  NEED(s); NEED(n);
  *n = S(s).hist_len;
  API_END
}
int mmg_solver_get_history(mmg_solver* s, double* out, int cap) {
  API_BEGIN
  NEED(s); NEED(out);
  Solver& so = S(s);
  if (!so.grids.empty()) use_device(so.grids[0]->device);
  fetch_history(so);
  for (int i = 0; i < so.hist_len && i < cap; i++) out[i] = so.hist_host[i];
  API_END
}
int mmg_solver_solve(mmg_solver* s, double tol, int max_cycles, int extra_bound_eval, int* cycles_done, double* final_residual) {
  API_BEGIN
  NEED(s);
  Solver& so = S(s);
  MMG_REQUIRE(!so.grids.empty(), MMG_ERR_STATE, "solve: no grids");
  Grid& fine = *so.grids.back();
  use_device(fine.device);
  DevBuf<double> ratio;
  ratio.alloc(1);
  int n = 0;
  double r = 0;
  while (true) {                       // while (mg.residual() >= tol) { mg.vCycle(); [bound_eval_neumann();] }
    if (so.world > 1) { if (!so.dist_ready) dist_setup(so); dist_residual_norm(so, (int)so.grids.size() - 1, ratio.p); }
    else op_residual_norm(fine, ratio.p);
    ratio.download(&r, 1, so.stream);
    if (!(r >= tol) || n >= max_cycles) break;
    vcycle(so);
    if (extra_bound_eval) op_bound_eval_neumann(fine);
    n++;
  }
  MMG_CUDA(cudaStreamSynchronize(so.stream));
  solver_check_abort(so);
  if (cycles_done) *cycles_done = n;
  if (final_residual) *final_residual = r;
  API_END
}
// multi-GPU: after a partitioned vcycle / solve a rank's values_ are current on its row block and halo ranges only; this
// collective makes every partitioned level's values_ complete on every rank (grouped ncclBroadcast of the row blocks)
// rows [own_lo, own_hi) of `level` belong to this rank; its kernels read values_ in [need_lo, need_hi) (block + halo)
int mmg_solver_owned_range(mmg_solver* s, int level, int* own_lo, int* own_hi, int* need_lo, int* need_hi) {
  API_BEGIN
  NEED(s); NEED(own_lo); NEED(own_hi); NEED(need_lo); NEED(need_hi);
  Solver& so = S(s);
  MMG_REQUIRE(level >= 0 && level < (int)so.grids.size(), MMG_ERR_ARG, "owned_range: no such level");
  Grid& g = *so.grids[level];
  *own_lo = 0; *own_hi = g.A; *need_lo = 0; *need_hi = g.A;
  if (so.world > 1) {
    if (!so.dist_ready) { use_device(g.device); dist_setup(so); }
    const LevelDist& D = so.dist[level];
    if (D.partitioned) {
      *own_lo = D.bounds[so.rank]; *own_hi = D.bounds[so.rank + 1];
      *need_lo = *own_lo; *need_hi = *own_hi;
      for (const ExchangePlan::Msg& m : D.x_plan.recvs) { *need_lo = std::min(*need_lo, m.offset); *need_hi = std::max(*need_hi, m.offset + m.count); }
    }
  }
  API_END
}
int mmg_solver_gather_values(mmg_solver* s) {
  API_BEGIN
  NEED(s);
  Solver& so = S(s);
  if (so.world > 1 && so.dist_ready) {
    use_device(so.grids[0]->device);
    for (size_t l = 0; l < so.grids.size(); l++)
      if (so.dist[l].partitioned) allgather_blocks(so, so.grids[l]->x.p, so.dist[l].bounds);
  }
  API_END
}
int mmg_solver_sync(mmg_solver* s) {
  API_BEGIN
  NEED(s);
  if (S(s).stream) MMG_CUDA(cudaStreamSynchronize(S(s).stream));
  API_END
}
int mmg_solver_enable_timers(mmg_solver* s, int on) {
  API_BEGIN
  NEED(s);
  S(s).timers.on = on != 0;
  API_END
}
int mmg_solver_get_timers(mmg_solver* s, int level, double* ms, int64_t* launches, int64_t* bytes) {
  API_BEGIN
  NEED(s);
  Solver& so = S(s);
  MMG_REQUIRE(level >= -1 && level < kMaxLevels, MMG_ERR_ARG, "timers: level out of range (-1 = all levels)");
  if (so.stream) MMG_CUDA(cudaStreamSynchronize(so.stream));
  timers_collect(so.timers);
  for (int i = 0; i < MMG_T_COUNT; i++) {
    double m = 0; int64_t l = 0, b = 0;
    for (int lv = 0; lv < kMaxLevels; lv++)
      if (level < 0 || lv == level) { m += so.timers.ms[lv][i]; l += so.timers.launches[lv][i]; b += so.timers.bytes[lv][i]; }
    if (ms) ms[i] = m;
    if (launches) launches[i] = l;
    if (bytes) bytes[i] = b;
  }
  API_END
}
int mmg_solver_reset_timers(mmg_solver* s) {
  API_BEGIN
  NEED(s);
  Solver& so = S(s);
  if (so.stream) MMG_CUDA(cudaStreamSynchronize(so.stream));
  timers_collect(so.timers);
  for (int lv = 0; lv < kMaxLevels; lv++)
    for (int i = 0; i < MMG_T_COUNT; i++) { so.timers.ms[lv][i] = 0; so.timers.launches[lv][i] = 0; so.timers.bytes[lv][i] = 0; }
  API_END
}
// diagnostics: which kernel instantiation the last smoother (slot 0) / SpMV-class (slot 1) call of this thread launched
int mmg_debug_last_kernel(int slot, char* out, int cap) {
  API_BEGIN
  NEED(out);
  MMG_REQUIRE(cap > 0 && (slot == 0 || slot == 1), MMG_ERR_ARG, "last_kernel: slot is 0 (smoother) or 1 (SpMV class)");
  const std::string& n = last_kernel_slot(slot);
  snprintf(out, (size_t)cap, "%s", n.c_str());
  API_END
}
// diagnostics: clock64 stamps of one CTA of the chunked lexicographic kernel (MMG_LEX_TRACE=1)
int mmg_debug_lex_trace(long long* out, int n) {
  API_BEGIN
  mmg::debug_lex_trace(out, n);
  API_END
}
// ---------------------------------------------------------------------------------------------- multi-GPU
int mmg_partition_bounds(int n, int world, int* bounds) {
  API_BEGIN
  NEED(bounds);
  MMG_REQUIRE(n >= 0 && world >= 1, MMG_ERR_ARG, "partition_bounds: bad sizes");
  partition_bounds(n, world, bounds);
  API_END
}
int mmg_comm_unique_id(char* out128) {
  API_BEGIN
  NEED(out128);
  comm_unique_id(out128);
  API_END
}
int mmg_solver_init_comm(mmg_solver* s, int rank, int world, const char* id128) {
  API_BEGIN
  NEED(s);
  MMG_REQUIRE(world == 1 || id128 != nullptr, MMG_ERR_ARG, "init_comm: the NCCL unique id is missing");
  use_device(S(s).grids.empty() ? 0 : S(s).grids[0]->device);
  comm_init(S(s), rank, world, id128);
  S(s).dist_ready = false;
  API_END
}
int mmg_solver_set_partition_threshold(mmg_solver* s, int rows) {
  API_BEGIN
  NEED(s);
  S(s).part_threshold = rows;
  S(s).dist_ready = false;
  API_END
}
int mmg_solver_comm_stats(mmg_solver* s, int64_t* messages, int64_t* bytes_sent, int* partitioned_levels) {
  API_BEGIN
  NEED(s);
  if (messages) *messages = S(s).comm_msgs;
  if (bytes_sent) *bytes_sent = S(s).comm_bytes;
  if (partitioned_levels) { int n = 0; for (const LevelDist& d : S(s).dist) n += d.partitioned; *partitioned_levels = n; }
  API_END
}
// diagnostics: the exchange plan a rank derives from everybody's need intervals; pure host logic
int mmg_debug_exchange_plan(int rank, int world, const int* need, const int* bounds, int* n_send, int* sends, int* n_recv, int* recvs) {
  API_BEGIN
  std::vector<std::pair<int, int>> nd(world);
  for (int r = 0; r < world; r++) nd[r] = {need[2 * r], need[2 * r + 1]};
  std::vector<int> b(bounds, bounds + world + 1);
  ExchangePlan P;
  plan_build(P, rank, world, nd, b);
  *n_send = (int)P.sends.size(); *n_recv = (int)P.recvs.size();
  for (size_t i = 0; i < P.sends.size(); i++) { sends[3 * i] = P.sends[i].peer; sends[3 * i + 1] = P.sends[i].offset; sends[3 * i + 2] = P.sends[i].count; }
  for (size_t i = 0; i < P.recvs.size(); i++) { recvs[3 * i] = P.recvs[i].peer; recvs[3 * i + 1] = P.recvs[i].offset; recvs[3 * i + 2] = P.recvs[i].count; }
  API_END
}
int mmg_solver_launch_count(mmg_solver* s, int64_t* launches) {
  API_BEGIN
  NEED(s); NEED(launches);
  *launches = S(s).timers.total_launches;
  API_END
}

}  // extern "C"
