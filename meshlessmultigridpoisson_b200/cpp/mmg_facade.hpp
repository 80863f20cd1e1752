// C++ facade: the reference's Grid / Multigrid / FractionalStepMultigrid classes re-created over the
// C-ABI of libmmg (include/mmg.h), so the reference's drivers (testing_functions.cpp:328-350,
// FractionalStepSim.cpp:114-200) keep their call sites.  Same method names, same argument meaning, same
// mode strings ("fine"/"coarse", "dirichlet"/"neumann"); errors surface as std::runtime_error because
// a CUDA failure has no CPU fallback to fall back to.
//
// The reference exposes raw public members (values_, source_, residuals_) that its drivers read and
// write directly.  Here they are HostVector mirrors: reads pull from the device when the device copy is
// newer, writes are pushed before the next device operation.  With real Eigen on the include path,
// HostVector converts to and from Eigen::VectorXd.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <fstream>
#include <iostream>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "../../include/mmg.h"

#if defined(__has_include)
#if __has_include(<Eigen/Dense>) && !defined(MMG_FACADE_NO_EIGEN)
#include <Eigen/Dense>
#define MMG_FACADE_HAVE_EIGEN 1
#endif
#endif

namespace mmgf {

typedef std::tuple<double, double, double> Point;   // grid.h:16

inline void check(int rc, const char* what) {
  if (rc != MMG_OK) throw std::runtime_error(std::string(what) + ": " + mmg_last_error());
}

// gridclasses.hpp:6-20
struct GridProperties { int rbfExp = 3, polyDeg = 3, laplaceMatSize = 0, stencilSize = 25; double omega = 1.4; int iters = 5; };
struct Boundary { int type = 0; std::vector<int> bcPoints; std::vector<double> values; };

// host mirror of a device vector (values_ / source_)
class HostVector {
 public:
  typedef int (*Getter)(mmg_grid*, double*);
  typedef int (*Setter)(mmg_grid*, const double*);
  HostVector() {}
  void bind(mmg_grid* g, int n, Getter get, Setter set) { g_ = g; h_.assign(n, 0.0); get_ = get; set_ = set; device_newer_ = true; }
  int rows() const { return (int)h_.size(); }
  int size() const { return (int)h_.size(); }
  double coeff(int i) const { pull(); return h_[i]; }
  double operator()(int i) const { return coeff(i); }
  double& coeffRef(int i) { pull(); host_newer_ = true; return h_[i]; }
  double& operator()(int i) { return coeffRef(i); }
  void setZero() { std::fill(h_.begin(), h_.end(), 0.0); device_newer_ = false; host_newer_ = true; }
  const std::vector<double>& host() const { pull(); return h_; }
  HostVector& operator=(const std::vector<double>& v) { h_ = v; device_newer_ = false; host_newer_ = true; return *this; }
  // *u_old = *u copies the values, never the binding
  HostVector& operator=(const HostVector& o) { if (this != &o) { if (g_) *this = o.host(); else { h_ = o.host(); } } return *this; }
  HostVector(const HostVector&) = delete;
  double lpNorm1() const { pull(); double s = 0; for (double t : h_) s += t < 0 ? -t : t; return s; }
  template <int P> double lpNorm() const { static_assert(P == 1, "lpNorm<1> only (multigrid.cpp:114)"); return lpNorm1(); }
  double maxCoeff() const { pull(); return *std::max_element(h_.begin(), h_.end()); }
  double minCoeff() const { pull(); return *std::min_element(h_.begin(), h_.end()); }
#ifdef MMG_FACADE_HAVE_EIGEN
  operator Eigen::VectorXd() const { return head(rows()); }
  Eigen::VectorXd head(int n) const {                  // values_->head(laplaceMatSize_), FractionalStepSim.cpp:105-111
    pull();
    Eigen::VectorXd v(n);
    for (int i = 0; i < n; i++) v(i) = h_[i];
    return v;
  }
  HostVector& operator=(const Eigen::VectorXd& v) { h_.resize(v.rows()); for (int i = 0; i < (int)v.rows(); i++) h_[i] = v(i); device_newer_ = false; host_newer_ = true; return *this; }
#endif
  // facade internals
  void push() { if (host_newer_) { check(set_(g_, h_.data()), "upload"); host_newer_ = false; } }
  void invalidate() { device_newer_ = true; }
 private:
  void pull() const {
    if (device_newer_ && !host_newer_) { check(get_(g_, const_cast<double*>(h_.data())), "download"); device_newer_ = false; }
  }
  mmg_grid* g_ = nullptr;
  mutable std::vector<double> h_;
  Getter get_ = nullptr;
  Setter set_ = nullptr;
  mutable bool device_newer_ = false;
  bool host_newer_ = false;
};

template <class GridT> class BasicMultigrid;

#ifdef MMG_FACADE_HAVE_EIGEN
typedef Eigen::VectorXd DenseVector;                   // what Grid::residual() returns in the reference (grid.h:47)
inline std::vector<double> to_std(const Eigen::VectorXd& v) { std::vector<double> o(v.rows()); for (int i = 0; i < (int)v.rows(); i++) o[i] = v(i); return o; }
inline Eigen::VectorXd operator+(const HostVector& a, const Eigen::VectorXd& b) { return Eigen::VectorXd(a) + b; }
inline Eigen::VectorXd operator+(const Eigen::VectorXd& a, const HostVector& b) { return a + Eigen::VectorXd(b); }
inline Eigen::VectorXd operator-(const HostVector& a, const Eigen::VectorXd& b) { return Eigen::VectorXd(a) - b; }
inline Eigen::VectorXd operator-(const Eigen::VectorXd& a, const HostVector& b) { return a - Eigen::VectorXd(b); }
// laplaceMat_ / derivXMat_ / derivYMat_ / uvLaplaceMat_ of the reference are host sparse matrices the drivers multiply by host
// vectors; here they are handles to the device operators and the product runs on the GPU (mmg_grid_apply_matrix)
struct DeviceMatrix {
  mmg_grid* g = nullptr;
  int which = 0, rows = 0, cols = 0;
  Eigen::VectorXd times(const Eigen::VectorXd& x, double scale) const {
    std::vector<double> in = to_std(x), out(rows);
    in.resize(cols, 0.0);
    check(mmg_grid_apply_matrix(g, which, in.data(), out.data()), "sparse * dense");
    Eigen::VectorXd y(rows);
    for (int i = 0; i < rows; i++) y(i) = scale * out[i];
    return y;
  }
};
struct ScaledDeviceMatrix { const DeviceMatrix* m; double s; };
inline Eigen::VectorXd operator*(const DeviceMatrix& m, const Eigen::VectorXd& x) { return m.times(x, 1.0); }
inline Eigen::VectorXd operator*(const DeviceMatrix& m, const HostVector& x) { return m.times(Eigen::VectorXd(x), 1.0); }
inline ScaledDeviceMatrix operator*(double s, const DeviceMatrix& m) { return ScaledDeviceMatrix{&m, s}; }
inline Eigen::VectorXd operator*(const ScaledDeviceMatrix& m, const Eigen::VectorXd& x) { return m.m->times(x, m.s); }
inline Eigen::VectorXd operator*(const ScaledDeviceMatrix& m, const HostVector& x) { return m.m->times(Eigen::VectorXd(x), m.s); }
#else
typedef std::vector<double> DenseVector;
struct DeviceMatrix { mmg_grid* g = nullptr; int which = 0, rows = 0, cols = 0; };
#endif

// grid.h:20-79
class Grid {
 public:
  HostVector values_holder_;
  HostVector* values_ = &values_holder_;   // the reference keeps a pointer (grid.h:23): (*grid->values_)(i) still compiles
  HostVector source_;
  DeviceMatrix laplaceMat_holder_;
  DeviceMatrix* laplaceMat_ = &laplaceMat_holder_;   // grid.h:25; passed back to sor() by testGmshSingleGrid (testing_functions.cpp:440)
  std::vector<Point> points_;
  std::vector<Boundary> boundaries_;
  GridProperties properties_;
  int laplaceMatSize_ = 0;
  bool neumannFlag_ = false;
  bool implicitFlag_ = false;

  Grid(std::vector<Point> points, std::vector<Boundary> boundaries, GridProperties properties, const std::vector<double>& source, int device = 0)
      : points_(std::move(points)), boundaries_(std::move(boundaries)), properties_(properties) {
    const int n = (int)points_.size();
    std::vector<double> x(n), y(n);
    for (int i = 0; i < n; i++) { x[i] = std::get<0>(points_[i]); y[i] = std::get<1>(points_[i]); }
    std::vector<int> type, ptr(1, 0), pts;
    std::vector<double> vals;
    for (const Boundary& b : boundaries_) {
      type.push_back(b.type);
      pts.insert(pts.end(), b.bcPoints.begin(), b.bcPoints.end());
      std::vector<double> v = b.values;
      v.resize(b.bcPoints.size(), 0.0);
      vals.insert(vals.end(), v.begin(), v.end());
      ptr.push_back((int)pts.size());
      if (b.type == MMG_BC_NEUMANN) neumannFlag_ = true;
    }
    mmg_props p{properties.rbfExp, properties.polyDeg, properties.stencilSize, properties.iters, properties.omega};
    check(mmg_grid_create(&h_, device, n, x.data(), y.data(), &p, source.data(), (int)source.size(), (int)boundaries_.size(), type.data(), ptr.data(),
                          pts.data(), vals.data()), "Grid::Grid");
    laplaceMatSize_ = n;
    const int A = neumannFlag_ ? n + 1 : n;
    values_holder_.bind(h_, A, mmg_grid_get_values, mmg_grid_set_values);
    source_.bind(h_, A, mmg_grid_get_source, mmg_grid_set_source);
    laplaceMat_holder_ = DeviceMatrix{h_, MMG_MAT_LAPLACE, A, A};
  }
#ifdef MMG_FACADE_HAVE_EIGEN
  // the reference's signature: Grid(points, boundaries, properties, Eigen::VectorXd source) grid.h:41
  Grid(std::vector<Point> points, std::vector<Boundary> boundaries, GridProperties properties, const Eigen::VectorXd& source, int device = 0)
      : Grid(std::move(points), std::move(boundaries), properties, to_std(source), device) {}
#endif
  ~Grid() { if (h_ && owned_) mmg_grid_destroy(h_); }
  Grid(const Grid&) = delete;
  Grid& operator=(const Grid&) = delete;

  void setBCFlag(int boundary, std::string type, std::vector<double> boundValue) {
    const int t = type.compare("dirichlet") == 0 ? MMG_BC_DIRICHLET : MMG_BC_NEUMANN;   // grid.cpp:35
    check(mmg_grid_set_implicit(h_, implicitFlag_), "implicitFlag_");
    check(mmg_grid_set_bc_flag(h_, boundary, t, boundValue.data(), (int)boundValue.size()), "Grid::setBCFlag");
    boundaries_.at(boundary).type = t;
    boundaries_.at(boundary).values = boundValue;
  }
  void build_normal_vecs(const char* /*filename*/, std::string geomtype) {
    const int geom = geomtype == "square" ? MMG_GEOM_SQUARE : geomtype == "square_with_circle" ? MMG_GEOM_SQUARE_WITH_CIRCLE
                     : geomtype == "concentric_circles" ? MMG_GEOM_CONCENTRIC_CIRCLES : -1;
    if (geom < 0) return;                             // the reference ignores unknown geomtype strings (grid.cpp:442-516)
    check(mmg_grid_build_normal_vecs(h_, geom), "Grid::build_normal_vecs");
  }
  void rcm_order_points() { sync_flags(); push(); check(mmg_grid_rcm_order_points(h_), "Grid::rcm_order_points"); refresh(); }
  void build_deriv_normal_bound() { sync_flags(); check(mmg_grid_build_deriv_normal_bound(h_), "Grid::build_deriv_normal_bound"); }
  void build_laplacian() { sync_flags(); check(mmg_grid_build_laplacian(h_), "Grid::build_laplacian"); }
  void modify_coeff_neumann(std::string coarse) { push(); check(mmg_grid_modify_coeff_neumann(h_, coarse == "coarse"), "Grid::modify_coeff_neumann"); source_.invalidate(); }
  void push_inhomog_to_rhs() { sync_flags(); push(); check(mmg_grid_push_inhomog_to_rhs(h_), "Grid::push_inhomog_to_rhs"); source_.invalidate(); }
  void boundaryOp(std::string coarse) { push(); check(mmg_grid_boundary_op(h_, coarse == "coarse"), "Grid::boundaryOp"); values_->invalidate(); }
  void bound_eval_neumann() { push(); check(mmg_grid_bound_eval_neumann(h_), "Grid::bound_eval_neumann"); values_->invalidate(); }
  // Grid::sor(laplaceMat_, values_, &source_): the reference always passes its own members (multigrid.cpp:79,93-94,108)
  void sor() { push(); check(mmg_grid_sor(h_, MMG_SMOOTHER_LEXICOGRAPHIC), "Grid::sor"); values_->invalidate(); }
  void sor(DeviceMatrix* matrix, HostVector* values, HostVector* rhs) {          // grid.h:44
    if (matrix != laplaceMat_ || values != values_ || rhs != &source_) throw std::runtime_error("Grid::sor: the device smoother works on the grid's own members");
    sor();
  }
  DenseVector residual() {
    push();
    std::vector<double> r(values_->rows());
    check(mmg_grid_residual(h_, r.data()), "Grid::residual");
#ifdef MMG_FACADE_HAVE_EIGEN
    Eigen::VectorXd out((int)r.size());
    for (int i = 0; i < (int)r.size(); i++) out(i) = r[i];
    return out;
#else
    return r;
#endif
  }
  void fix_vector_bound_coarse(std::vector<double>* vec) { check(mmg_grid_fix_vector_bound_coarse(h_, vec->data()), "Grid::fix_vector_bound_coarse"); }
  int getSize() const { return laplaceMatSize_; }
  int getStencilSize() const { return properties_.stencilSize; }
  int getPolyDeg() const { return properties_.polyDeg; }
  mmg_grid* handle() { return h_; }

  // ---- per-point queries of grid.h:60-72.  The bulk assembly never goes through them (one launch builds every row); they are
  // here for callers that inspect single stencils.  Weights come back as (weights, neighbour ids) like in the reference, the
  // weight vector as DenseVector (Eigen::VectorXd with Eigen, std::vector<double> without).
  std::vector<Point> pointIDs_to_vector(const std::vector<int>& pointIDs) const {          // grid.cpp:206-212
    std::vector<Point> pts;
    for (size_t i = 0; i < pointIDs.size(); i++) pts.push_back(points_.at(pointIDs.at(i)));
    return pts;
  }
  std::vector<int> kNearestNeighbors(Point point, bool neumannFlag, bool pointBCFlag, int k) {   // grid.cpp:216-260
    sync_flags();
    const double x = std::get<0>(point), y = std::get<1>(point);
    const int flag = pointBCFlag ? 1 : 0;
    std::vector<int> out(k);
    check(mmg_grid_knn(h_, 1, &x, &y, &flag, neumannFlag ? 1 : 0, k, out.data()), "Grid::kNearestNeighbors");
    return out;
  }
  std::vector<int> kNearestNeighbors(int pointNumber, bool neumannFlag, int k) {            // grid.cpp:213-215
    return kNearestNeighbors(points_.at(pointNumber), neumannFlag, bcFlags().at(pointNumber) != 0, k);
  }
  std::pair<DenseVector, std::vector<int>> laplaceWeights(int pointID) { return weights_of(MMG_MAT_LAPLACE, pointID, "Grid::laplaceWeights"); }   // grid.cpp:381-424
  std::pair<DenseVector, std::vector<int>> derivx_weights(int pointID) { return weights_of(MMG_MAT_DERIVX, pointID, "Grid::derivx_weights"); }    // grid.cpp:304-342
  std::pair<DenseVector, std::vector<int>> derivy_weights(int pointID) { return weights_of(MMG_MAT_DERIVY, pointID, "Grid::derivy_weights"); }    // grid.cpp:343-380
  std::pair<DenseVector, std::vector<int>> pointInterpWeights(Point point, int polyDeg) {   // grid.cpp:687-712
    sync_flags();
    const int n = (int)(2.5 * (polyDeg + 1) * (polyDeg + 2) / 2);                            // grid.cpp:266-267
    const double x = std::get<0>(point), y = std::get<1>(point);
    std::vector<double> w(n);
    std::vector<int> nb(n);
    check(mmg_grid_point_interp_weights(h_, 1, &x, &y, polyDeg, w.data(), nb.data()), "Grid::pointInterpWeights");
    return std::make_pair(to_dense(w), nb);
  }
  // public data members of grid.h:23-38 that live on the device: fetched on request
  std::vector<int> bcFlags() {                                                              // bcFlags_
    std::vector<int> f(laplaceMatSize_);
    check(mmg_grid_get_bcflags(h_, f.data()), "bcFlags_");
    return f;
  }
  std::vector<Point> normalVecs() {                                                         // normalVecs_ (after build_normal_vecs)
    std::vector<double> nx(laplaceMatSize_), ny(laplaceMatSize_);
    check(mmg_grid_get_normals(h_, nx.data(), ny.data()), "normalVecs_");
    std::vector<Point> out(laplaceMatSize_);
    for (int i = 0; i < laplaceMatSize_; i++) out[i] = Point(nx[i], ny[i], 0.0);
    return out;
  }
  DenseVector diags() {                                                                     // diags (after build_laplacian)
    std::vector<double> d(values_->rows());
    check(mmg_grid_get_diags(h_, d.data()), "diags");
    return to_dense(d);
  }
  void setNeumannFlag() {                                                                   // grid.cpp:52-60
    neumannFlag_ = false;
    for (const Boundary& b : boundaries_) if (b.type == MMG_BC_NEUMANN) neumannFlag_ = true;
  }
  void print_bc_values() {                                                                  // grid.cpp:165-171
    for (const Boundary& b : boundaries_)
      for (size_t j = 0; j < b.bcPoints.size(); j++) std::cout << "bc value: " << b.values.at(j) << std::endl;
  }
  void print_bc_values(const DenseVector& vec) {                                            // grid.cpp:156-164: Dirichlet boundaries only
    for (const Boundary& b : boundaries_)
      if (b.type == MMG_BC_DIRICHLET)
        for (size_t j = 0; j < b.bcPoints.size(); j++) std::cout << "bc value: " << entry(vec, b.bcPoints.at(j)) << std::endl;
  }

 protected:
  void push_all() { sync_flags(); push(); }

 private:
  template <class GridT> friend class BasicMultigrid;
  void sync_flags() { check(mmg_grid_set_implicit(h_, implicitFlag_), "implicitFlag_"); }
  void push() { values_->push(); source_.push(); }
#ifdef MMG_FACADE_HAVE_EIGEN
  static double entry(const Eigen::VectorXd& v, int i) { return v(i); }
#else
  static double entry(const std::vector<double>& v, int i) { return v.at(i); }
#endif
  static DenseVector to_dense(const std::vector<double>& v) {
#ifdef MMG_FACADE_HAVE_EIGEN
    Eigen::VectorXd out((int)v.size());
    for (int i = 0; i < (int)v.size(); i++) out(i) = v[i];
    return out;
#else
    return v;
#endif
  }
  std::pair<DenseVector, std::vector<int>> weights_of(int which, int pointID, const char* what) {
    sync_flags();
    const int n = properties_.stencilSize;
    std::vector<double> w(n);
    std::vector<int> nb(n);
    check(mmg_grid_weights(h_, which, 1, &pointID, w.data(), nb.data()), what);
    return std::make_pair(to_dense(w), nb);
  }
  void refresh() {   // after a reordering the host-side copies of points_/boundaries_ follow the device
    const int n = laplaceMatSize_;
    std::vector<double> x(n), y(n);
    check(mmg_grid_get_points(h_, x.data(), y.data()), "points_");
    for (int i = 0; i < n; i++) points_[i] = Point(x[i], y[i], 0.0);
    for (size_t b = 0; b < boundaries_.size(); b++) {
      int type = 0, count = 0;
      check(mmg_grid_get_boundary(h_, (int)b, &type, &count, nullptr, nullptr), "boundaries_");
      boundaries_[b].bcPoints.resize(count); boundaries_[b].values.resize(count);
      check(mmg_grid_get_boundary(h_, (int)b, &type, &count, boundaries_[b].bcPoints.data(), boundaries_[b].values.data()), "boundaries_");
    }
    source_.invalidate();
  }
  mmg_grid* h_ = nullptr;
  bool owned_ = true;
};

// fractionalStepGrid.hpp:4-30.  The six velocity vectors live on the device; u, v, ... are HostVector mirrors bound to
// them, so driver statements like (*grid->u)(i) or grid->u_old = ... keep compiling.  dt / mu / rho are plain members the
// drivers assign after construction (FractionalStepSim.cpp:26-29); they are pushed with every call that reads them.
class FractionalStepGrid : public Grid {
 public:
  double dt = 0, ppe_conv_res = 0, rho = 1, mu = 1, lambda = 0;
  std::string flowType = "kovasznay";
  HostVector u_h_, v_h_, u_old_h_, v_old_h_, u_hat_h_, v_hat_h_;
  HostVector *u = &u_h_, *v = &v_h_, *u_old = &u_old_h_, *v_old = &v_old_h_, *u_hat = &u_hat_h_, *v_hat = &v_hat_h_;
  DeviceMatrix dx_h_, dy_h_, lap_h_;
  DeviceMatrix *derivXMat_ = &dx_h_, *derivYMat_ = &dy_h_, *uvLaplaceMat_ = &lap_h_;   // fractionalStepGrid.hpp:16-18

  FractionalStepGrid(std::vector<Point> points, std::vector<Boundary> boundaries, GridProperties properties, const std::vector<double>& source, int device = 0)
      : Grid(std::move(points), std::move(boundaries), properties, source, device) {
    check(mmg_grid_fs_init(handle(), dt, mu, rho), "FractionalStepGrid::FractionalStepGrid");
    const int n = laplaceMatSize_;
    u_h_.bind(handle(), n, get<MMG_FS_U>, set<MMG_FS_U>);             v_h_.bind(handle(), n, get<MMG_FS_V>, set<MMG_FS_V>);
    u_old_h_.bind(handle(), n, get<MMG_FS_U_OLD>, set<MMG_FS_U_OLD>); v_old_h_.bind(handle(), n, get<MMG_FS_V_OLD>, set<MMG_FS_V_OLD>);
    u_hat_h_.bind(handle(), n, get<MMG_FS_U_HAT>, set<MMG_FS_U_HAT>); v_hat_h_.bind(handle(), n, get<MMG_FS_V_HAT>, set<MMG_FS_V_HAT>);
    dx_h_ = DeviceMatrix{handle(), MMG_MAT_DERIVX, n, n}; dy_h_ = DeviceMatrix{handle(), MMG_MAT_DERIVY, n, n}; lap_h_ = DeviceMatrix{handle(), MMG_MAT_UVLAPLACE, n, n};
  }
#ifdef MMG_FACADE_HAVE_EIGEN
  FractionalStepGrid(std::vector<Point> points, std::vector<Boundary> boundaries, GridProperties properties, const Eigen::VectorXd& source, int device = 0)
      : FractionalStepGrid(std::move(points), std::move(boundaries), properties, to_std(source), device) {}
#endif
  void prescribe_soln() {                          // fractionalStepGrid.cpp:26-40: the exact Kovasznay fields (debug helper of check_derivs)
    const double PI = 3.141592653589793238462643383279502884;
    const double re = rho / mu;
    lambda = 0.5 * re - std::sqrt(0.25 * re * re + 4 * PI * PI);
    for (int i = 0; i < laplaceMatSize_; i++) {
      const double x = std::get<0>(points_[i]), y = std::get<1>(points_[i]);
      u->coeffRef(i) = 1 - std::exp(lambda * x) * std::cos(2 * PI * y);
      v->coeffRef(i) = lambda / (2 * PI) * std::exp(lambda * x) * std::sin(2 * PI * y);
      u_old->coeffRef(i) = 1 - std::exp(lambda * x) * std::cos(2 * PI * y);
      v_old->coeffRef(i) = lambda / (2 * PI) * std::exp(lambda * x) * std::sin(2 * PI * y);
      values_->coeffRef(i) = 0.5 * std::exp(2 * lambda * x);
    }
    values_->coeffRef(laplaceMatSize_) = 0;
    sync_fs();
  }
  void set_uv_bound() {
    if (flowType != "kovasznay") return;          // the reference does nothing for any other flow type (fractionalStepGrid.cpp:45)
    sync_fs(); check(mmg_grid_fs_set_uv_bound(handle()), "set_uv_bound");
    u_h_.invalidate(); v_h_.invalidate(); u_old_h_.invalidate(); v_old_h_.invalidate();
  }
  // the three builders share one kNN pass on the device; the first call builds all of them
  void build_derivX_mat() { build_ops(); }
  void build_derivY_mat() { build_ops(); }
  void build_uv_laplace_mat() { build_ops(); }
  void calc_u_hat() { sync_fs(); check(mmg_grid_fs_calc_hat(handle(), MMG_FS_U), "calc_u_hat"); u_hat_h_.invalidate(); }
  void calc_v_hat() { sync_fs(); check(mmg_grid_fs_calc_hat(handle(), MMG_FS_V), "calc_v_hat"); v_hat_h_.invalidate(); }
  void set_ppe_source() { sync_fs(); check(mmg_grid_fs_set_ppe_source(handle()), "set_ppe_source"); source_.invalidate(); }
  void correct_u() { sync_fs(); check(mmg_grid_fs_correct(handle(), MMG_FS_U), "correct_u"); u_h_.invalidate(); }
  void correct_v() { sync_fs(); check(mmg_grid_fs_correct(handle(), MMG_FS_V), "correct_v"); v_h_.invalidate(); }
  double fs_residual() {
    sync_fs();
    double r = 0;
    check(mmg_grid_fs_residual(handle(), &r), "fs_residual");
    return r;
  }

 private:
  bool built_ = false;
  template <int W> static int get(mmg_grid* g, double* out) { return mmg_grid_fs_get_vec(g, W, out); }
  template <int W> static int set(mmg_grid* g, const double* in) { return mmg_grid_fs_set_vec(g, W, in); }
  void build_ops() {
    if (built_) return;
    sync_fs(); check(mmg_grid_fs_build_operators(handle()), "build_deriv*_mat"); built_ = true;
  }
  void sync_fs() {
    push_all();
    check(mmg_grid_fs_init(handle(), dt, mu, rho), "dt/mu/rho");
    u_h_.push(); v_h_.push(); u_old_h_.push(); v_old_h_.push(); u_hat_h_.push(); v_hat_h_.push();
  }
};

// multigrid.h:4-23; FractionalStepMultigrid (FracStepMultigrid.hpp:4-25) is the same class over FractionalStepGrid* with
// the twin's flavour, so both are one template here.
template <class GridT>
class BasicMultigrid {
 public:
  std::vector<std::pair<int, GridT*>> grids_;
  std::vector<double> residuals_;

  explicit BasicMultigrid(int flavour = MMG_FLAVOUR_MULTIGRID) { check(mmg_solver_create(&h_, flavour), "Multigrid::Multigrid"); }
  BasicMultigrid(const BasicMultigrid&) = delete;
  BasicMultigrid& operator=(const BasicMultigrid&) = delete;
  ~BasicMultigrid() {
    for (auto& g : grids_) delete g.second;      // takes ownership like multigrid.cpp:10-16 (device grids go with the solver)
    mmg_solver_destroy(h_);
  }
  void addGrid(GridT* grid) {
    grid->push();
    check(mmg_solver_add_grid(h_, grid->h_), "Multigrid::addGrid");
    grid->owned_ = false;
    grids_.push_back(std::pair<int, GridT*>(grid->getSize(), grid));
    std::sort(grids_.begin(), grids_.end());     // multigrid.cpp:116-122
  }
  void buildMatrices() {
    for (auto& g : grids_) g.second->push();
    check(mmg_solver_build_matrices(h_), "Multigrid::buildMatrices");
    for (auto& g : grids_) g.second->source_.invalidate();
  }
  void vCycle() {
    for (auto& g : grids_) g.second->push();
    check(mmg_solver_vcycle(h_, 1), "Multigrid::vCycle");
    for (auto& g : grids_) { g.second->values_->invalidate(); g.second->source_.invalidate(); }
    int n = 0;
    check(mmg_solver_history_len(h_, &n), "residuals_");
    residuals_.resize(n);
    if (n) check(mmg_solver_get_history(h_, residuals_.data(), n), "residuals_");
  }
  double residual() {
    for (auto& g : grids_) g.second->push();
    double r = 0;
    check(mmg_solver_residual(h_, &r), "Multigrid::residual");
    return r;
  }
  void setSmoother(int smoother) { check(mmg_solver_set_smoother(h_, smoother), "setSmoother"); }
  mmg_solver* handle() { return h_; }
  void sortGridsBySize() { std::sort(grids_.begin(), grids_.end()); }                       // multigrid.cpp:116-122
  // restrictionMatrices_[level] / prolongMatrices_[level] (multigrid.h:8-9) live on the device; a host copy in CSR on request.
  // Levels as in the reference: restriction `level` maps grid level -> level-1 (level >= 1), prolongation `level` maps
  // grid level -> level+1 (level < numGrids-1).
  struct InterpCsr { int rows = 0, cols = 0; std::vector<int> ptr, idx; std::vector<double> val; };
  InterpCsr restrictionMatrix(int level) { return interp(MMG_MAT_RESTRICT, level); }
  InterpCsr prolongMatrix(int level) { return interp(MMG_MAT_PROLONG, level); }

 private:
  InterpCsr interp(int which, int level) {
    InterpCsr m;
    int64_t nnz = 0;
    check(mmg_solver_interp_nnz(h_, which, level, &m.rows, &m.cols, &nnz), "restrictionMatrices_/prolongMatrices_");
    m.ptr.resize(m.rows + 1); m.idx.resize((size_t)nnz); m.val.resize((size_t)nnz);
    check(mmg_solver_get_interp_csr(h_, which, level, m.ptr.data(), m.idx.data(), m.val.data()), "restrictionMatrices_/prolongMatrices_");
    return m;
  }
  mmg_solver* h_ = nullptr;
};

typedef BasicMultigrid<Grid> Multigrid;

class FractionalStepMultigrid : public BasicMultigrid<FractionalStepGrid> {
 public:
  FractionalStepMultigrid() : BasicMultigrid<FractionalStepGrid>(MMG_FLAVOUR_FRACSTEP) {}
  void solveLoop() {}   // empty in the reference too (FracStepMultigrid.cpp:113-115)
};

}  // namespace mmgf

// Text writers live in their own namespace so that argument-dependent lookup does not find them when the reference's unmodified
// drivers (which define the same function names) are compiled against this header (cpp/dropin/).
namespace mmgf_io {
using mmgf::Grid;
using mmgf::BasicMultigrid;
// ---- the reference's text writers, same file names and number format (std::ofstream default: 6 significant digits, one
// value per line), so the author's plotting scripts read the files unchanged
// Gmsh v2 ASCII $Nodes block -> points, what pointsFromMshFile does in the reference (fileReadingFunctions.cpp:6-32); unlike the
// reference a missing file or a malformed block is an exception, not an endless loop
inline std::vector<mmgf::Point> pointsFromMshFile(const char* fname) {
  std::vector<mmgf::Point> points;
  FILE* f = std::fopen(fname, "r");
  if (!f) throw std::runtime_error(std::string("cannot open ") + fname);
  char tok[64];
  bool found = false;
  while (std::fscanf(f, "%63s ", tok) == 1) if (std::strcmp(tok, "$Nodes") == 0) { found = true; break; }
  int nv = 0;
  if (!found || std::fscanf(f, "%i ", &nv) != 1) { std::fclose(f); throw std::runtime_error(std::string("no $Nodes block in ") + fname); }
  for (int iv = 0; iv < nv; iv++) {
    int id;
    double x, y, z;
    if (std::fscanf(f, "%i %lf %lf %lf ", &id, &x, &y, &z) != 4) { std::fclose(f); throw std::runtime_error(std::string("bad node line in ") + fname); }
    points.push_back(mmgf::Point(x, y, z));
  }
  std::fclose(f);
  return points;
}
inline void writeVectorToTxt(const std::vector<double>& vec, const char* filename) {      // fileReadingFunctions.cpp:70-79
  std::ofstream file;
  file.open(filename);
  for (size_t i = 0; i < vec.size(); i++) file << vec[i] << "\n";
  file.close();
}
inline void write_temp_contour(Grid* testGrid, const std::string& directory, const std::string& extension) {   // testing_functions.cpp:285-307
  std::vector<double> xv, yv, temp;
  for (size_t i = 0; i < testGrid->points_.size(); i++) { xv.push_back(std::get<0>(testGrid->points_[i])); yv.push_back(std::get<1>(testGrid->points_[i])); }
  writeVectorToTxt(xv, (directory + "x_" + extension + ".txt").c_str());
  writeVectorToTxt(yv, (directory + "y_" + extension + ".txt").c_str());
  for (int i = 0; i < testGrid->values_->rows() - 1; i++) temp.push_back(testGrid->values_->coeff(i));      // rows()-1, like the reference
  writeVectorToTxt(temp, (directory + "temp_" + extension + ".txt").c_str());
}
template <class GridT>
inline void write_mg_resid(BasicMultigrid<GridT>& mg, const std::string& directory, const std::string& extension) {   // testing_functions.cpp:308-312
  writeVectorToTxt(mg.residuals_, (directory + "resid_" + extension + ".txt").c_str());
}

}  // namespace mmgf_io

