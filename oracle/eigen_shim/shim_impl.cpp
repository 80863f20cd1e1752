// TEST INFRASTRUCTURE ONLY — arithmetic behind the Eigen-subset shim; forwards to the oracle's
// restatement of Eigen 3.4.0 semantics so the reference build and the oracle share one reading.
#include "eigen_shim.hpp"
#include "../mmg_oracle.hpp"

namespace Eigen {
template <> VectorXd FullPivLU<MatrixXd>::solve(const VectorXd& b) const {
  std::vector<double> A = m_.storage(), bb(b.data(), b.data() + b.rows()), x;
  orc::fullpivlu_solve(A, (int)m_.rows(), bb, x);
  VectorXd r((Index)x.size());
  for (size_t i = 0; i < x.size(); i++) r((Index)i) = x[i];
  return r;
}
void shim_set_from_triplets(int outer, int inner, const std::vector<int>& o, const std::vector<int>& i, const std::vector<double>& v,
                            std::vector<int>& ptr, std::vector<int>& idx, std::vector<double>& val) {
  std::vector<orc::Trip> t(o.size());
  for (size_t k = 0; k < o.size(); k++) t[k] = orc::Trip{o[k], i[k], v[k]};
  orc::Csr A = orc::csr_from_triplets(outer, inner, t);
  ptr = A.ptr; idx = A.idx; val = A.val;
}
void shim_rowmajor_times(int rows, const int* ptr, const int* idx, const double* val, const double* x, double* y) {
  // Eigen 3.4.0 sparse_time_dense_product_impl<RowMajor>: tmp=0; tmp += a*x; res += 1*tmp
  for (int r = 0; r < rows; r++) {
    double tmp = 0;
    for (int k = ptr[r]; k < ptr[r + 1]; k++) tmp += val[k] * x[idx[k]];
    y[r] = tmp;
  }
}
void shim_colmajor_times(int rows, int cols, const int* ptr, const int* idx, const double* val, const double* x, double* y) {
  // Eigen 3.4.0 sparse_time_dense_product_impl<ColMajor>: res=0; for each column j: res(i) += a_ij * (1*x_j)
  for (int r = 0; r < rows; r++) y[r] = 0;
  for (int c = 0; c < cols; c++) {
    const double xc = x[c];
    for (int k = ptr[c]; k < ptr[c + 1]; k++) y[idx[k]] += val[k] * xc;
  }
}
}  // namespace Eigen
