// TEST INFRASTRUCTURE ONLY — drives the reference's OWN classes (compiled from
// /root/reference/MeshlessPoisson against the Eigen-subset shim) the way run_mg_sim
// (testing_functions.cpp:328-350) and run_fracstep_param (FractionalStepSim.cpp:114-147) do, and
// exposes their public members so tests can compare them with the oracle restatement.
#include <chrono>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>

#include "FractionalStepSim.hpp"
#include "testing_functions.hpp"

namespace {
struct Ref {
  Multigrid* mg = nullptr;
  FractionalStepMultigrid* fmg = nullptr;
  Grid* grid(int l) { return mg ? mg->grids_.at(l).second : (Grid*)fmg->grids_.at(l).second; }
  int nlev() { return mg ? (int)mg->grids_.size() : (int)fmg->grids_.size(); }
};
struct Quiet {  // Multigrid::vCycle prints every cycle (multigrid.cpp:69)
  std::ostringstream sink; std::streambuf* old;
  Quiet() { old = std::cout.rdbuf(sink.rdbuf()); }
  ~Quiet() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {
// kind 0 Dirichlet / 1 Neumann (Multigrid), 2 PPE (FractionalStepMultigrid); files are Gmsh $Nodes files
void* ref_new_geom(int kind, const char* geomtype, int nfiles, const char** files, const int* polyDeg, int k1, int k2, double dt, double mu, double rho);
void* ref_new(int kind, int nfiles, const char** files, const int* polyDeg, int k1, int k2, double dt, double mu, double rho) {
  return ref_new_geom(kind, "square", nfiles, files, polyDeg, k1, k2, dt, mu, rho);
}
void* ref_new_geom(int kind, const char* geomtype, int nfiles, const char** files, const int* polyDeg, int k1, int k2, double dt, double mu, double rho) {
  Quiet q;
  Ref* r = new Ref();
  if (kind == 2) r->fmg = new FractionalStepMultigrid(); else r->mg = new Multigrid();
  for (int i = 0; i < nfiles; i++) {
    GridProperties p;   // gen_mg_param, testing_functions.cpp:372-380
    p.iters = 5; p.polyDeg = polyDeg[i]; p.omega = 1.4; p.rbfExp = 3;
    p.stencilSize = (int)(2.5 * (p.polyDeg + 1) * (p.polyDeg + 2) / 2);
    const std::string coarse = (i == nfiles - 1) ? "fine" : "coarse";
    if (kind == 0) r->mg->addGrid(genGmshGridDirichlet(geomtype, files[i], p, "msh", k1, k2));
    else if (kind == 1) r->mg->addGrid(genGmshGridNeumann(geomtype, files[i], p, "msh", k1, k2, coarse));
    else r->fmg->addGrid(genFractionalStepGrid(files[i], p, dt, mu, rho, 1e-10, coarse));
  }
  if (r->mg) r->mg->buildMatrices(); else r->fmg->buildMatrices();
  return r;
}
// ---- the reference's own Multigrid over operators assembled elsewhere (Dirichlet levels).  The reference's set-up is a brute-force
// kNN, O(N^2) per level and O(N_fine x N_coarse) per interpolation matrix, which rules it out beyond ~100k nodes; its V-cycle
// (multigrid.cpp:62-110, grid.cpp:104-151) has no such limit.  So bench.py's reference arm assembles with the oracle's threaded
// set-up (pinned bit-identical to the reference's, test_oracle_is_bit_identical_to_the_reference_sources) and hands the matrices
// to unmodified reference objects: every cycle that is timed runs the reference's own code.
void* ref_new_raw() {
  Ref* r = new Ref();
  r->mg = new Multigrid();
  return r;
}
void ref_free(void* h) {
  Ref* r = (Ref*)h;
  if (r->mg) {   // ~Multigrid deletes prolongMatrices_.at(i) / restrictionMatrices_.at(i) for every grid (multigrid.cpp:10-16)
    r->mg->prolongMatrices_.resize(r->mg->grids_.size(), nullptr);
    r->mg->restrictionMatrices_.resize(r->mg->grids_.size(), nullptr);
    delete r->mg;
  }
  delete r->fmg;
  delete r;
}
// one Dirichlet level: points, ONE boundary (ids + values), properties, source_, laplaceMat_ as CSR with ascending columns
void ref_add_level_raw(void* h, int n, const double* x, const double* y, int polyDeg, int iters, double omega, int rbfExp, int nb, const int* bpts,
                       const double* bvals, const double* source, const int* ptr, const int* idx, const double* val) {
  Ref* r = (Ref*)h;
  std::vector<Point> pts(n);
  for (int i = 0; i < n; i++) pts[i] = Point(x[i], y[i], 0.0);
  Boundary b;
  b.type = 1;
  b.bcPoints.assign(bpts, bpts + nb);
  b.values.assign(bvals, bvals + nb);
  GridProperties p;
  p.rbfExp = rbfExp; p.polyDeg = polyDeg; p.iters = iters; p.omega = omega; p.laplaceMatSize = n;
  p.stencilSize = (int)(2.5 * (polyDeg + 1) * (polyDeg + 2) / 2);
  Eigen::VectorXd src(n);
  for (int i = 0; i < n; i++) src(i) = source[i];
  Grid* g = new Grid(pts, std::vector<Boundary>{b}, p, src);
  g->implicitFlag_ = false;
  g->setBCFlag(0, std::string("dirichlet"), b.values);       // genGmshGridDirichlet, testing_functions.cpp:150-152
  g->laplaceMat_->adoptCompressed(ptr, idx, val, ptr[n]);    // in place of rcm_order_points() + build_laplacian()
  r->mg->addGrid(g);
}
// restrictionMatrices_[level] (which 0) / prolongMatrices_[level] (which 1): column-major like the reference's (multigrid.h:8-9)
void ref_set_interp_raw(void* h, int which, int level, int rows, int cols, const int* colptr, const int* rowidx, const double* val) {
  Ref* r = (Ref*)h;
  auto& vec = which == 0 ? r->mg->restrictionMatrices_ : r->mg->prolongMatrices_;
  vec.resize(r->mg->grids_.size(), nullptr);
  auto* M = new Eigen::SparseMatrix<double>(rows, cols);
  M->adoptCompressed(colptr, rowidx, val, colptr[cols]);
  delete vec.at(level);
  vec.at(level) = M;
}
void ref_vcycle(void* h, int n) {
  Quiet q;
  Ref* r = (Ref*)h;
  for (int i = 0; i < n; i++) { if (r->mg) r->mg->vCycle(); else r->fmg->vCycle(); }
}
double ref_time_vcycles(void* h, int n) {
  Quiet q;
  Ref* r = (Ref*)h;
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < n; i++) { if (r->mg) r->mg->vCycle(); else r->fmg->vCycle(); }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
double ref_residual(void* h) { Ref* r = (Ref*)h; return r->mg ? r->mg->residual() : r->fmg->residual(); }
int ref_history(void* h, double* out, int cap) {
  Ref* r = (Ref*)h;
  const std::vector<double>& v = r->mg ? r->mg->residuals_ : r->fmg->residuals_;
  for (int i = 0; i < (int)v.size() && i < cap; i++) out[i] = v[i];
  return (int)v.size();
}
int ref_nlevels(void* h) { return ((Ref*)h)->nlev(); }
int ref_lv_n(void* h, int l) { return ((Ref*)h)->grid(l)->laplaceMatSize_; }
int ref_lv_A(void* h, int l) { return (int)((Ref*)h)->grid(l)->laplaceMat_->rows(); }
long ref_lv_nnz(void* h, int l) { return (long)((Ref*)h)->grid(l)->laplaceMat_->nonZeros(); }
void ref_lv_csr(void* h, int l, int* ptr, int* idx, double* val) {
  auto* A = ((Ref*)h)->grid(l)->laplaceMat_;
  std::memcpy(ptr, A->outerIndexPtr(), sizeof(int) * (A->rows() + 1));
  std::memcpy(idx, A->innerIndexPtr(), sizeof(int) * A->nonZeros());
  std::memcpy(val, A->valuePtr(), sizeof(double) * A->nonZeros());
}
void ref_lv_points(void* h, int l, double* x, double* y) {
  Grid* g = ((Ref*)h)->grid(l);
  for (int i = 0; i < g->laplaceMatSize_; i++) { x[i] = std::get<0>(g->points_[i]); y[i] = std::get<1>(g->points_[i]); }
}
void ref_lv_vec(void* h, int l, int which, double* out) {  // 0 values_, 1 source_
  Grid* g = ((Ref*)h)->grid(l);
  const Eigen::VectorXd& v = which == 0 ? *g->values_ : g->source_;
  for (Eigen::Index i = 0; i < v.rows(); i++) out[i] = v.coeff(i);
}
void ref_lv_bcflags(void* h, int l, int* f) {
  Grid* g = ((Ref*)h)->grid(l);
  for (size_t i = 0; i < g->bcFlags_.size(); i++) f[i] = g->bcFlags_[i];
}
// interp matrices are column-major in the reference (multigrid.h:8-9); exported as stored
long ref_interp_nnz(void* h, int which, int l) {
  Ref* r = (Ref*)h;
  auto* M = which == 0 ? (r->mg ? r->mg->restrictionMatrices_.at(l) : r->fmg->restrictionMatrices_.at(l))
                       : (r->mg ? r->mg->prolongMatrices_.at(l) : r->fmg->prolongMatrices_.at(l));
  return M ? (long)M->nonZeros() : -1;
}
void ref_interp_csc(void* h, int which, int l, int* rows, int* cols, int* ptr, int* idx, double* val) {
  Ref* r = (Ref*)h;
  auto* M = which == 0 ? (r->mg ? r->mg->restrictionMatrices_.at(l) : r->fmg->restrictionMatrices_.at(l))
                       : (r->mg ? r->mg->prolongMatrices_.at(l) : r->fmg->prolongMatrices_.at(l));
  *rows = (int)M->rows(); *cols = (int)M->cols();
  std::memcpy(ptr, M->outerIndexPtr(), sizeof(int) * (M->cols() + 1));
  std::memcpy(idx, M->innerIndexPtr(), sizeof(int) * M->nonZeros());
  std::memcpy(val, M->valuePtr(), sizeof(double) * M->nonZeros());
}
// fractional-step time step around the PPE solve (FractionalStepSim.cpp:131-147)
void ref_fs_step_pre(void* h) {
  FractionalStepGrid* g = ((Ref*)h)->fmg->grids_.back().second;
  *(g->u_old) = *(g->u); *(g->v_old) = *(g->v);
  g->set_uv_bound(); g->calc_u_hat(); g->calc_v_hat(); g->set_ppe_source(); g->push_inhomog_to_rhs();
}
double ref_fs_step_post(void* h) {
  FractionalStepGrid* g = ((Ref*)h)->fmg->grids_.back().second;
  g->correct_u(); g->correct_v(); g->set_uv_bound();
  return g->fs_residual();
}
void ref_fine_bound_eval(void* h) {   // finestGrid->bound_eval_neumann() of the PPE loop (FractionalStepSim.cpp:141)
  Ref* r = (Ref*)h;
  r->grid(r->nlev() - 1)->bound_eval_neumann();
}
void ref_fs_vec(void* h, int which, double* out) {  // 0 u, 1 v, 2 u_hat, 3 v_hat
  FractionalStepGrid* g = ((Ref*)h)->fmg->grids_.back().second;
  const Eigen::VectorXd& v = which == 0 ? *g->u : which == 1 ? *g->v : which == 2 ? *g->u_hat : *g->v_hat;
  for (Eigen::Index i = 0; i < v.rows(); i++) out[i] = v.coeff(i);
}
}
