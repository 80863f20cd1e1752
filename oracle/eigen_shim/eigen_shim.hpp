// ============================================================================
// TEST INFRASTRUCTURE ONLY — minimal stand-in for the subset of Eigen 3.4 that
// the reference (MeshlessPoisson/*.cpp) uses, so that the reference's OWN
// sources can be compiled where they lie into oracle/_ref/libref.so and run
// beside the oracle restatement.  Eigen itself is not vendored by the
// reference and is absent from this image.
//
// Containers are eager (no expression templates).  The arithmetic that real
// Eigen would perform — FullPivLU, setFromTriplets, sparse*dense — is NOT
// re-derived here: shim_impl.cpp forwards to the oracle's restatement of
// Eigen 3.4.0 semantics (mmg_oracle.cpp), so libref.so validates every line
// of the reference's own algorithmic code against the oracle, while sharing
// the oracle's reading of Eigen.  Never used by the product.
// ============================================================================
#pragma once
#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <vector>

// Real Eigen defines this with an L suffix; the reference was built with MSVC,
// where long double == double, so the faithful value is the double literal.
#define EIGEN_PI 3.141592653589793238462643383279502884197169399375105820974944592307816406

namespace Eigen {
typedef std::ptrdiff_t Index;
enum { ColMajor = 0, RowMajor = 1 };

struct SeqRange { Index first, last; };
inline SeqRange seq(Index a, Index b) { return SeqRange{a, b}; }

class VectorXd;
class VecSegment {  // writable view returned by v(seq(a,b)) / v.head(n)
 public:
  VecSegment(double* p, Index n) : p_(p), n_(n) {}
  VecSegment& operator=(const VectorXd& o);
  VecSegment& operator+=(const VectorXd& o);
  Index rows() const { return n_; }
  const double* data() const { return p_; }
 private:
  double* p_;
  Index n_;
};

class VectorXd {
 public:
  VectorXd() {}
  explicit VectorXd(Index n) : d_(n) {}
  VectorXd(const VecSegment& s) : d_(s.data(), s.data() + s.rows()) {}
  static VectorXd Zero(Index n) { VectorXd v(n); v.setZero(); return v; }
  Index rows() const { return (Index)d_.size(); }
  Index size() const { return (Index)d_.size(); }
  void setZero() { for (double& t : d_) t = 0; }
  double& operator()(Index i) { return d_[i]; }
  double operator()(Index i) const { return d_[i]; }
  double& coeffRef(Index i) { return d_[i]; }
  double coeff(Index i) const { return d_[i]; }
  VecSegment operator()(SeqRange r) { return VecSegment(d_.data() + r.first, r.last - r.first + 1); }
  VecSegment head(Index n) { return VecSegment(d_.data(), n); }
  double* data() { return d_.data(); }
  const double* data() const { return d_.data(); }
  VectorXd& operator*=(double s) { for (double& t : d_) t *= s; return *this; }
  VectorXd& operator+=(const VectorXd& o) { for (size_t i = 0; i < d_.size(); i++) d_[i] += o.d_[i]; return *this; }
  template <int P> double lpNorm() const { static_assert(P == 1, "shim: lpNorm<1> only"); double s = 0; for (double t : d_) s += std::fabs(t); return s; }
  double norm() const { double s = 0; for (double t : d_) s += t * t; return std::sqrt(s); }
  double maxCoeff() const { double m = d_[0]; for (double t : d_) if (t > m) m = t; return m; }
  double minCoeff() const { double m = d_[0]; for (double t : d_) if (t < m) m = t; return m; }
 private:
  std::vector<double> d_;
};
inline VectorXd operator-(const VectorXd& a, const VectorXd& b) { VectorXd r(a.rows()); for (Index i = 0; i < a.rows(); i++) r(i) = a(i) - b(i); return r; }
inline VectorXd operator+(const VectorXd& a, const VectorXd& b) { VectorXd r(a.rows()); for (Index i = 0; i < a.rows(); i++) r(i) = a(i) + b(i); return r; }
inline VectorXd operator*(double s, const VectorXd& a) { VectorXd r(a.rows()); for (Index i = 0; i < a.rows(); i++) r(i) = s * a(i); return r; }
inline VecSegment& VecSegment::operator=(const VectorXd& o) { for (Index i = 0; i < n_; i++) p_[i] = o(i); return *this; }
inline VecSegment& VecSegment::operator+=(const VectorXd& o) { for (Index i = 0; i < n_; i++) p_[i] += o(i); return *this; }

class MatrixXd;
template <class M> class FullPivLU {
 public:
  explicit FullPivLU(const M& m) : m_(m) {}
  VectorXd solve(const VectorXd& b) const;   // shim_impl.cpp -> orc::fullpivlu_solve
 private:
  M m_;
};
template <class M> class PartialPivLU {
 public:
  explicit PartialPivLU(const M&) {}
  double rcond() const { throw std::runtime_error("eigen_shim: PartialPivLU::rcond (diagnostic, grid.cpp:153) is not provided"); }
};
class MatrixXd {  // column-major, like Eigen's default
 public:
  MatrixXd() : r_(0), c_(0) {}
  MatrixXd(Index r, Index c) : r_(r), c_(c), d_((size_t)r * c) {}
  static MatrixXd Zero(Index r, Index c) { MatrixXd m(r, c); for (double& t : m.d_) t = 0; return m; }
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  double& operator()(Index i, Index j) { return d_[i + (size_t)j * r_]; }
  double operator()(Index i, Index j) const { return d_[i + (size_t)j * r_]; }
  const std::vector<double>& storage() const { return d_; }
  FullPivLU<MatrixXd> fullPivLu() const { return FullPivLU<MatrixXd>(*this); }
  PartialPivLU<MatrixXd> partialPivLu() const { return PartialPivLU<MatrixXd>(*this); }
 private:
  Index r_, c_;
  std::vector<double> d_;
};

template <class T> class Triplet {
 public:
  Triplet() : r_(0), c_(0), v_(0) {}
  Triplet(int r, int c, const T& v) : r_(r), c_(c), v_(v) {}
  int row() const { return r_; }
  int col() const { return c_; }
  const T& value() const { return v_; }
 private:
  int r_, c_;
  T v_;
};

// Compressed storage along the outer dimension (rows if RowMajor, columns otherwise).
void shim_set_from_triplets(int outer, int inner, const std::vector<int>& o, const std::vector<int>& i, const std::vector<double>& v,
                            std::vector<int>& ptr, std::vector<int>& idx, std::vector<double>& val);
void shim_rowmajor_times(int rows, const int* ptr, const int* idx, const double* val, const double* x, double* y);
void shim_colmajor_times(int rows, int cols, const int* ptr, const int* idx, const double* val, const double* x, double* y);

template <class T, int Opt = ColMajor, class I = int> class SparseMatrix {
 public:
  SparseMatrix() : r_(0), c_(0), ptr_(1, 0) {}
  SparseMatrix(Index r, Index c) : r_(r), c_(c), ptr_((Opt == RowMajor ? r : c) + 1, 0) {}
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Index nonZeros() const { return (Index)idx_.size(); }
  void setZero() { idx_.clear(); val_.clear(); for (int& p : ptr_) p = 0; }
  void makeCompressed() {}
  template <class It> void setFromTriplets(It b, It e) {
    std::vector<int> o, i; std::vector<double> v;
    for (It t = b; t != e; ++t) { o.push_back(Opt == RowMajor ? t->row() : t->col()); i.push_back(Opt == RowMajor ? t->col() : t->row()); v.push_back(t->value()); }
    shim_set_from_triplets((int)(Opt == RowMajor ? r_ : c_), (int)(Opt == RowMajor ? c_ : r_), o, i, v, ptr_, idx_, val_);
  }
  T* valuePtr() { return val_.data(); }
  const T* valuePtr() const { return val_.data(); }
  const int* innerIndexPtr() const { return idx_.data(); }
  const int* outerIndexPtr() const { return ptr_.data(); }
  T coeff(Index i, Index j) const {
    const Index o = Opt == RowMajor ? i : j, in = Opt == RowMajor ? j : i;
    for (int k = ptr_[o]; k < ptr_[o + 1]; k++) if (idx_[k] == in) return val_[k];
    return T(0);
  }
  MatrixXd toDense() const {
    MatrixXd m = MatrixXd::Zero(r_, c_);
    const Index no = Opt == RowMajor ? r_ : c_;
    for (Index o = 0; o < no; o++) for (int k = ptr_[o]; k < ptr_[o + 1]; k++) { if (Opt == RowMajor) m(o, idx_[k]) = val_[k]; else m(idx_[k], o) = val_[k]; }
    return m;
  }
  VectorXd times(const double* x) const {
    VectorXd y(r_);
    if (Opt == RowMajor) shim_rowmajor_times((int)r_, ptr_.data(), idx_.data(), val_.data(), x, y.data());
    else shim_colmajor_times((int)r_, (int)c_, ptr_.data(), idx_.data(), val_.data(), x, y.data());
    return y;
  }
  SparseMatrix scaled(double s) const { SparseMatrix m(*this); for (T& t : m.val_) t = s * t; return m; }
  // NOT Eigen API (with real Eigen: Eigen::Map<const SparseMatrix>): take over compressed arrays as they are -- outer pointers
  // (rows for RowMajor, columns for ColMajor), inner indices ascending within an outer slice, values.  Only ref_capi.cpp's raw
  // injection uses it, to put operators assembled by the oracle's threaded set-up under the reference's own V-cycle.
  void adoptCompressed(const int* outer, const int* inner, const T* val, Index nnz) {
    ptr_.assign(outer, outer + (Opt == RowMajor ? r_ : c_) + 1);
    idx_.assign(inner, inner + nnz);
    val_.assign(val, val + nnz);
  }
 private:
  Index r_, c_;
  std::vector<int> ptr_, idx_;
  std::vector<T> val_;
};
template <class T, int O, class I> VectorXd operator*(const SparseMatrix<T, O, I>& A, const VectorXd& x) { return A.times(x.data()); }
template <class T, int O, class I> VectorXd operator*(const SparseMatrix<T, O, I>& A, const VecSegment& x) { return A.times(x.data()); }
template <class T, int O, class I> SparseMatrix<T, O, I> operator*(double s, const SparseMatrix<T, O, I>& A) { return A.scaled(s); }
}  // namespace Eigen
