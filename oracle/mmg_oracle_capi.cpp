// TEST INFRASTRUCTURE ONLY — flat C interface over the CPU oracle for ctypes (tests/, bench.py
// cpu_baseline / --impl reference, __graft_entry__.smoke()).  Not part of the product.
#include <chrono>
#include <cstring>
#include <stdexcept>
#include <string>

#include "mmg_oracle.hpp"

using namespace orc;

namespace {
thread_local std::string g_err;
struct Handle {
  Multigrid mg;
};
Grid* lv(void* h, int l) { return static_cast<Handle*>(h)->mg.grids_.at(l).second; }
Csr* pick(void* h, int l, int which) {
  Multigrid& mg = static_cast<Handle*>(h)->mg;
  Grid* g = mg.grids_.at(l).second;
  switch (which) {
    case 0: return &g->laplaceMat_;
    case 1: return &g->neumann_boundary_coeffs_;
    case 2: return &mg.restrictionMatrices_.at(l);
    case 3: return &mg.prolongMatrices_.at(l);
    case 4: return &static_cast<FractionalStepGrid*>(g)->derivXMat_;
    case 5: return &static_cast<FractionalStepGrid*>(g)->derivYMat_;
    case 6: return &static_cast<FractionalStepGrid*>(g)->uvLaplaceMat_;
  }
  return nullptr;
}
std::vector<double>* pickv(void* h, int l, int which) {
  Grid* g = lv(h, l);
  switch (which) {
    case 0: return &g->values_;
    case 1: return &g->source_;
    case 2: return &g->diags;
    case 3: return &static_cast<FractionalStepGrid*>(g)->u;
    case 4: return &static_cast<FractionalStepGrid*>(g)->v;
    case 5: return &static_cast<FractionalStepGrid*>(g)->u_hat;
    case 6: return &static_cast<FractionalStepGrid*>(g)->v_hat;
  }
  return nullptr;
}
}  // namespace

#define ORC_TRY try {
#define ORC_CATCH(ret)                 \
  }                                    \
  catch (const std::exception& e) {    \
    g_err = e.what();                  \
    return ret;                        \
  }

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

void* orc_mg_new(int fracstep) {
  Handle* h = new Handle();
  h->mg.fracstep = fracstep != 0;
  return h;
}
void orc_mg_free(void* h) { delete static_cast<Handle*>(h); }

// kind: 0 Dirichlet square, 1 Neumann square, 2 fractional-step PPE, 3 mixed square
int orc_mg_add_level(void* h, int kind, int n, const double* x, const double* y, int polyDeg, int iters, double omega, int rbfExp, int k1, int k2,
                     int fine, int knn_mode, double dt, double mu, double rho) {
  ORC_TRY
  std::vector<Pt> pts(n);
  for (int i = 0; i < n; i++) pts[i] = Pt{x[i], y[i], 0.0};
  GridProperties p;
  p.rbfExp = rbfExp; p.polyDeg = polyDeg; p.omega = omega; p.iters = iters;
  p.stencilSize = (int)(2.5 * (polyDeg + 1) * (polyDeg + 2) / 2);  // testing_functions.cpp:378
  const std::string coarse = fine ? "fine" : "coarse";
  const KnnMode mode = (knn_mode & 1) ? KNN_CELLS : KNN_BRUTE;
  Grid* g = nullptr;
  const int geom = knn_mode >> 4;                    // bits 4..: geometry (0 square, 1 square_with_circle, 2 concentric_circles)
  if (kind == 0) g = genGridDirichlet(pts, p, k1, k2, mode, geom);
  else if (kind == 1) g = genGridNeumann(pts, p, k1, k2, coarse, mode, geom);
  else if (kind == 2) g = genFractionalStepGrid(pts, p, dt, mu, rho, 1e-10, coarse, mode);
  else if (kind == 3) g = genGridMixedSquare(pts, p, k1, k2, coarse, mode);
  else throw std::runtime_error("unknown level kind");
  static_cast<Handle*>(h)->mg.addGrid(g);
  return 0;
  ORC_CATCH(-1)
}

// A level whose state (reordered points, boundary lists, flags, source) was built elsewhere -- the tests use it to mirror a
// DEVICE-built hierarchy into the oracle at sizes where the oracle's own kNN + LU set-up would take minutes: the operators
// then arrive through orc_lv_set_csr (identical matrices on both sides).  No reordering, no assembly.
// btype[b] in {1 dirichlet, 2 neumann}; bptr has nb+1 entries into bpts / bvals.
int orc_mg_add_level_raw(void* h, int fracstep_grid, int n, const double* x, const double* y, int polyDeg, int stencil, int iters, double omega,
                         int rbfExp, int implicit, int nb, const int* btype, const int* bptr, const int* bpts, const double* bvals,
                         const double* source, int source_len) {
  ORC_TRY
  std::vector<Pt> pts(n);
  for (int i = 0; i < n; i++) pts[i] = Pt{x[i], y[i], 0.0};
  GridProperties p;
  p.rbfExp = rbfExp; p.polyDeg = polyDeg; p.omega = omega; p.iters = iters; p.stencilSize = stencil;
  std::vector<Boundary> bnds(nb);
  for (int b = 0; b < nb; b++) {
    bnds[b].type = btype[b];
    bnds[b].bcPoints.assign(bpts + bptr[b], bpts + bptr[b + 1]);
    bnds[b].values.assign(bvals + bptr[b], bvals + bptr[b + 1]);
  }
  std::vector<double> src(source, source + source_len);
  Grid* g = fracstep_grid ? new FractionalStepGrid(pts, bnds, p, src) : new Grid(pts, bnds, p, src);
  g->implicitFlag_ = implicit != 0;
  for (int b = 0; b < nb; b++) g->setBCFlag(b, btype[b] == 1 ? "dirichlet" : "neumann", bnds[b].values);
  if ((int)g->source_.size() != g->laplaceMat_.rows) throw std::runtime_error("source length does not match the level");
  g->order_.resize(n);
  for (int i = 0; i < n; i++) g->order_[i] = i;
  static_cast<Handle*>(h)->mg.addGrid(g);
  return 0;
  ORC_CATCH(-1)
}
// size the P / R slots without building them (they are then filled by orc_lv_set_csr)
void orc_mg_alloc_interp(void* h) {
  Multigrid& mg = static_cast<Handle*>(h)->mg;
  mg.prolongMatrices_.assign(mg.grids_.size(), Csr());
  mg.restrictionMatrices_.assign(mg.grids_.size(), Csr());
}

int orc_mg_build(void* h) {
  ORC_TRY
  static_cast<Handle*>(h)->mg.buildMatrices();
  return 0;
  ORC_CATCH(-1)
}
int orc_mg_nlevels(void* h) { return (int)static_cast<Handle*>(h)->mg.grids_.size(); }
void orc_mg_set_multicolour(void* h, int on) { static_cast<Handle*>(h)->mg.multicolour = on != 0; }
// smoother: 0 lexicographic (reference), 1 multicolour, 2 block-lexicographic with the given block size
void orc_mg_set_smoother(void* h, int smoother, int block_size) {
  Multigrid& mg = static_cast<Handle*>(h)->mg;
  mg.multicolour = smoother == 1;
  mg.blocklex = smoother == 2;
  if (smoother == 2)
    for (auto& g : mg.grids_) { g.second->block_size_ = block_size; g.second->block_colour_.clear(); }
}
void orc_mg_set_omega(void* h, double omega) {   // properties_.omega of every level (gridclasses.hpp:12)
  for (auto& g : static_cast<Handle*>(h)->mg.grids_) g.second->properties_.omega = omega;
}
void orc_lv_sor_blocklex(void* h, int l, int block_size) {
  Grid* g = lv(h, l);
  if (g->block_size_ != block_size) { g->block_size_ = block_size; g->block_colour_.clear(); }
  g->sor_blocklex(g->laplaceMat_, g->values_, g->source_);
}
int orc_lv_block_colouring(void* h, int l, int block_size, int* colour, int cap) {
  Grid* g = lv(h, l);
  g->block_size_ = block_size;
  g->build_block_colouring();
  for (int i = 0; i < (int)g->block_colour_.size() && i < cap; i++) colour[i] = g->block_colour_[i];
  return g->n_block_colours_;
}
int orc_mg_vcycle(void* h, int n) {
  ORC_TRY
  for (int i = 0; i < n; i++) static_cast<Handle*>(h)->mg.vCycle();
  return 0;
  ORC_CATCH(-1)
}
// wall-clock seconds around exactly the loop the reference times (testing_functions.cpp:340-344)
double orc_mg_time_vcycles(void* h, int n) {
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < n; i++) static_cast<Handle*>(h)->mg.vCycle();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
// while (mg.residual() >= tol) mg.vCycle();  (FractionalStepSim.cpp:139-142 loop shape); returns cycles, seconds in *secs
int orc_mg_solve(void* h, double tol, int max_cycles, double* secs) {
  Multigrid& mg = static_cast<Handle*>(h)->mg;
  auto t0 = std::chrono::steady_clock::now();
  int n = 0;
  while (n < max_cycles && mg.residual() >= tol) { mg.vCycle(); n++; }
  if (secs) *secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return n;
}
double orc_mg_residual(void* h) { return static_cast<Handle*>(h)->mg.residual(); }
int orc_mg_history(void* h, double* out, int cap) {
  const auto& r = static_cast<Handle*>(h)->mg.residuals_;
  for (int i = 0; i < (int)r.size() && i < cap; i++) out[i] = r[i];
  return (int)r.size();
}

int orc_lv_n(void* h, int l) { return lv(h, l)->laplaceMatSize_; }
int orc_lv_A(void* h, int l) { return lv(h, l)->laplaceMat_.rows; }
int orc_lv_neumann(void* h, int l) { return lv(h, l)->neumannFlag_; }
int orc_lv_implicit(void* h, int l) { return lv(h, l)->implicitFlag_; }
void orc_lv_props(void* h, int l, int* polyDeg, int* stencil, int* iters, int* rbfExp, double* omega) {
  const GridProperties& p = lv(h, l)->properties_;
  *polyDeg = p.polyDeg; *stencil = p.stencilSize; *iters = p.iters; *rbfExp = p.rbfExp; *omega = p.omega;
}
void orc_lv_get_points(void* h, int l, double* x, double* y) {
  Grid* g = lv(h, l);
  for (int i = 0; i < g->laplaceMatSize_; i++) { x[i] = g->points_[i].x; y[i] = g->points_[i].y; }
}
void orc_lv_get_normals(void* h, int l, double* nx, double* ny) {
  Grid* g = lv(h, l);
  for (int i = 0; i < g->laplaceMatSize_; i++) { nx[i] = g->normalVecs_[i].x; ny[i] = g->normalVecs_[i].y; }
}
void orc_lv_get_perm(void* h, int l, int* order) {
  Grid* g = lv(h, l);
  std::memcpy(order, g->order_.data(), sizeof(int) * g->order_.size());
}
void orc_lv_get_bcflags(void* h, int l, int* f) {
  Grid* g = lv(h, l);
  std::memcpy(f, g->bcFlags_.data(), sizeof(int) * g->bcFlags_.size());
}
int orc_lv_nboundaries(void* h, int l) { return (int)lv(h, l)->boundaries_.size(); }
int orc_lv_boundary_size(void* h, int l, int b) { return (int)lv(h, l)->boundaries_.at(b).bcPoints.size(); }
int orc_lv_get_boundary(void* h, int l, int b, int* pts, double* vals) {
  const Boundary& bd = lv(h, l)->boundaries_.at(b);
  for (size_t j = 0; j < bd.bcPoints.size(); j++) { pts[j] = bd.bcPoints[j]; vals[j] = bd.values[j]; }
  return bd.type;
}

long orc_lv_csr_nnz(void* h, int l, int which) { return (long)pick(h, l, which)->idx.size(); }
void orc_lv_csr_shape(void* h, int l, int which, int* rows, int* cols) {
  Csr* A = pick(h, l, which);
  *rows = A->rows; *cols = A->cols;
}
void orc_lv_get_csr(void* h, int l, int which, int* ptr, int* idx, double* val) {
  Csr* A = pick(h, l, which);
  std::memcpy(ptr, A->ptr.data(), sizeof(int) * A->ptr.size());
  std::memcpy(idx, A->idx.data(), sizeof(int) * A->idx.size());
  std::memcpy(val, A->val.data(), sizeof(double) * A->val.size());
}
void orc_lv_set_csr(void* h, int l, int which, int rows, int cols, const int* ptr, const int* idx, const double* val) {
  Csr* A = pick(h, l, which);
  A->rows = rows; A->cols = cols;
  A->ptr.assign(ptr, ptr + rows + 1);
  A->idx.assign(idx, idx + ptr[rows]);
  A->val.assign(val, val + ptr[rows]);
  if (which == 0) lv(h, l)->colour_.clear();
}
int orc_lv_vec_size(void* h, int l, int which) { return (int)pickv(h, l, which)->size(); }
void orc_lv_get_vec(void* h, int l, int which, double* out) {
  auto* v = pickv(h, l, which);
  std::memcpy(out, v->data(), sizeof(double) * v->size());
}
void orc_lv_set_vec(void* h, int l, int which, const double* in) {
  auto* v = pickv(h, l, which);
  std::memcpy(v->data(), in, sizeof(double) * v->size());
}

// per-operator entry points (mirror of the reference's Grid methods)
void orc_lv_sor(void* h, int l) { Grid* g = lv(h, l); g->sor(g->laplaceMat_, g->values_, g->source_); }
void orc_lv_sor_multicolour(void* h, int l) { Grid* g = lv(h, l); g->sor_multicolour(g->laplaceMat_, g->values_, g->source_); }
void orc_lv_residual(void* h, int l, double* out) {
  const std::vector<double> r = lv(h, l)->residual();
  std::memcpy(out, r.data(), sizeof(double) * r.size());
}
void orc_lv_bound_eval_neumann(void* h, int l) { lv(h, l)->bound_eval_neumann(); }
void orc_lv_boundary_op(void* h, int l, int coarse) { lv(h, l)->boundaryOp(coarse ? "coarse" : "fine"); }
void orc_lv_modify_coeff_neumann(void* h, int l, int coarse) { lv(h, l)->modify_coeff_neumann(coarse ? "coarse" : "fine"); }
void orc_lv_push_inhomog_to_rhs(void* h, int l) { lv(h, l)->push_inhomog_to_rhs(); }
void orc_lv_fix_vector_bound_coarse(void* h, int l, double* v) {
  Grid* g = lv(h, l);
  std::vector<double> t(v, v + g->laplaceMat_.rows);
  g->fix_vector_bound_coarse(t);
  std::memcpy(v, t.data(), sizeof(double) * t.size());
}
// y = M x for any stored matrix (Eigen product order)
void orc_lv_spmv(void* h, int l, int which, const double* x, double* y) { spmv(*pick(h, l, which), x, y); }

int orc_lv_knn(void* h, int l, double x, double y, int neumann, int pointBCFlag, int k, int mode, int* out) {
  ORC_TRY
  Grid* g = lv(h, l);
  const KnnMode keep = g->knn_mode;
  g->knn_mode = mode ? KNN_CELLS : KNN_BRUTE;
  std::vector<int> nn;
  try { nn = g->kNearestNeighbors(Pt{x, y, 0}, neumann != 0, pointBCFlag != 0, k); } catch (...) { g->knn_mode = keep; throw; }
  g->knn_mode = keep;
  std::memcpy(out, nn.data(), sizeof(int) * k);
  return 0;
  ORC_CATCH(-1)
}
// which: 0 laplace, 1 d/dx, 2 d/dy (point id); w has stencil+polyTerms entries, nb has stencil entries
int orc_lv_weights(void* h, int l, int which, int pointID, double* w, int* nb) {
  ORC_TRY
  Grid* g = lv(h, l);
  auto r = which == 0 ? g->laplaceWeights(pointID) : which == 1 ? g->derivx_weights(pointID) : g->derivy_weights(pointID);
  std::memcpy(w, r.first.data(), sizeof(double) * r.first.size());
  std::memcpy(nb, r.second.data(), sizeof(int) * r.second.size());
  return 0;
  ORC_CATCH(-1)
}
int orc_lv_interp_weights(void* h, int l, double x, double y, int polyDeg, double* w, int* nb) {
  ORC_TRY
  auto r = lv(h, l)->pointInterpWeights(Pt{x, y, 0}, polyDeg);
  std::memcpy(w, r.first.data(), sizeof(double) * r.first.size());
  std::memcpy(nb, r.second.data(), sizeof(int) * r.second.size());
  return 0;
  ORC_CATCH(-1)
}
// dense local system of grid.cpp:263-299 (column-major S x S), neighbour ids, scaled points (x,y pairs, n+2 of them)
int orc_lv_coeff_matrix(void* h, int l, double x, double y, int neumann, int pointBCFlag, int polyDeg, double* M, int* nb, double* sp) {
  ORC_TRY
  std::vector<double> Mx; std::vector<int> n; std::vector<Pt> s;
  lv(h, l)->buildCoeffMatrix(Pt{x, y, 0}, neumann != 0, pointBCFlag != 0, polyDeg, Mx, n, s);
  std::memcpy(M, Mx.data(), sizeof(double) * Mx.size());
  std::memcpy(nb, n.data(), sizeof(int) * n.size());
  for (size_t i = 0; i < s.size(); i++) { sp[2 * i] = s[i].x; sp[2 * i + 1] = s[i].y; }
  return 0;
  ORC_CATCH(-1)
}
int orc_lv_colouring(void* h, int l, int* colour) {
  Grid* g = lv(h, l);
  g->build_colouring();
  std::memcpy(colour, g->colour_.data(), sizeof(int) * g->colour_.size());
  return g->n_colours_;
}
void orc_lv_lex_levels(void* h, int l, int* lev) {
  const std::vector<int> v = lv(h, l)->lex_levels();
  std::memcpy(lev, v.data(), sizeof(int) * v.size());
}

// fractional-step operators (fractionalStepGrid.cpp:41-154); level must have been added with kind 2
void orc_fs_set_params(void* h, int l, double dt, double mu, double rho) {
  auto* g = static_cast<FractionalStepGrid*>(lv(h, l));
  g->dt = dt; g->mu = mu; g->rho = rho;
}
void orc_fs_step_pre(void* h, int l) {  // FractionalStepSim.cpp:131-137
  auto* g = static_cast<FractionalStepGrid*>(lv(h, l));
  g->u_old = g->u; g->v_old = g->v;
  g->set_uv_bound(); g->calc_u_hat(); g->calc_v_hat(); g->set_ppe_source(); g->push_inhomog_to_rhs();
}
double orc_fs_step_post(void* h, int l) {  // FractionalStepSim.cpp:144-147
  auto* g = static_cast<FractionalStepGrid*>(lv(h, l));
  g->correct_u(); g->correct_v(); g->set_uv_bound();
  return g->fs_residual();
}

// ---- free-standing pieces for unit tests ----
double orc_distance(double ax, double ay, double bx, double by) { return distance(Pt{ax, ay, 0}, Pt{bx, by, 0}); }
void orc_fullpivlu_solve(int n, const double* A_colmajor, const double* b, double* x) {
  std::vector<double> A(A_colmajor, A_colmajor + (size_t)n * n), bb(b, b + n), xx;
  fullpivlu_solve(A, n, bb, xx);
  std::memcpy(x, xx.data(), sizeof(double) * n);
}
long orc_csr_from_triplets(int rows, int cols, long nt, const int* r, const int* c, const double* v, int* ptr, int* idx, double* val) {
  std::vector<Trip> t(nt);
  for (long i = 0; i < nt; i++) t[i] = Trip{r[i], c[i], v[i]};
  Csr A = csr_from_triplets(rows, cols, t);
  std::memcpy(ptr, A.ptr.data(), sizeof(int) * A.ptr.size());
  std::memcpy(idx, A.idx.data(), sizeof(int) * A.idx.size());
  std::memcpy(val, A.val.data(), sizeof(double) * A.val.size());
  return (long)A.idx.size();
}
void orc_bfs_order(int n, const int* adj_ptr, const int* adj, int* order_out, int* count_out) {
  std::vector<std::vector<int>> a(n);
  for (int i = 0; i < n; i++) a[i].assign(adj + adj_ptr[i], adj + adj_ptr[i + 1]);
  std::vector<int> order(n);
  reverse_cuthill_mckee_ordering(a, order);
  *count_out = (int)order.size();
  std::memcpy(order_out, order.data(), sizeof(int) * order.size());
}
}
