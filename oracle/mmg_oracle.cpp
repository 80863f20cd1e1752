// TEST INFRASTRUCTURE ONLY — see mmg_oracle.hpp.  CPU restatement of the reference
// solve path; every function cites the reference lines it follows
// (paths relative to /root/reference/MeshlessPoisson/).
#include "mmg_oracle.hpp"

#include <algorithm>
#include <cmath>
#include <limits>
#include <queue>
#include <stdexcept>

namespace orc {

static const double kPi = 3.141592653589793238462643383279;  // testing_functions.hpp:9, grid.cpp:2

// ---------------------------------------------------------------------------------
// general_computation_functions.cpp
// ---------------------------------------------------------------------------------
double distance(const Pt& a, const Pt& b) {  // :4-6  (gcc folds pow(t,2) to t*t; checked in tests)
  return std::sqrt(std::pow(a.x - b.x, 2) + std::pow(a.y - b.y, 2));
}

std::vector<Pt> shifting_scaling(const std::vector<Pt>& pts, const Pt& eval) {  // :82-107 (+ minMaxCoord :8-35)
  double minX = pts[0].x, maxX = pts[0].x, minY = pts[0].y, maxY = pts[0].y;
  for (const Pt& p : pts) {
    if (p.x > maxX) maxX = p.x; else if (p.x < minX) minX = p.x;
    if (p.y > maxY) maxY = p.y; else if (p.y < minY) minY = p.y;
  }
  const double scale = std::max(maxX - minX, maxY - minY);
  std::vector<Pt> out;
  out.reserve(pts.size() + 2);
  for (const Pt& p : pts) out.push_back(Pt{(p.x - minX) / scale, (p.y - minY) / scale, 0});
  out.push_back(Pt{scale, scale, scale});
  out.push_back(Pt{(eval.x - minX) / scale, (eval.y - minY) / scale, 0});
  return out;
}

void reverse_cuthill_mckee_ordering(const std::vector<std::vector<int>>& adjacency, std::vector<int>& order) {  // :108-134
  std::vector<char> seen(order.size(), 0);
  std::queue<int> q;
  std::vector<int> visit;
  seen[0] = 1;
  q.push(0);
  while (!q.empty()) {
    const int cur = q.front();
    visit.push_back(cur);
    q.pop();
    for (int nb : adjacency[cur])
      if (!seen[nb]) { seen[nb] = 1; q.push(nb); }
  }
  order = visit;
  std::reverse(order.begin(), order.end());
}

// ---------------------------------------------------------------------------------
// Eigen 3.4.0 semantics, restated
// ---------------------------------------------------------------------------------
Csr csr_from_triplets(int rows, int cols, const std::vector<Trip>& t) {
  // SparseMatrix::setFromTriplets: bucket by outer index keeping triplet order, sum duplicates
  // into the first occurrence in that order, inner indices ascending, explicit zeros kept.
  Csr A;
  A.rows = rows; A.cols = cols;
  std::vector<int> cnt(rows + 1, 0);
  for (const Trip& e : t) cnt[e.r + 1]++;
  for (int r = 0; r < rows; r++) cnt[r + 1] += cnt[r];
  std::vector<int> pos(cnt.begin(), cnt.end() - 1);
  std::vector<std::pair<int, double>> tmp(t.size());
  for (const Trip& e : t) tmp[pos[e.r]++] = {e.c, e.v};
  A.ptr.assign(rows + 1, 0);
  for (int r = 0; r < rows; r++) {
    auto b = tmp.begin() + cnt[r], e = tmp.begin() + cnt[r + 1];
    std::stable_sort(b, e, [](const std::pair<int, double>& p, const std::pair<int, double>& q) { return p.first < q.first; });
    for (auto it = b; it != e;) {
      double s = it->second;
      auto jt = it + 1;
      for (; jt != e && jt->first == it->first; ++jt) s = s + jt->second;
      A.idx.push_back(it->first);
      A.val.push_back(s);
      it = jt;
    }
    A.ptr[r + 1] = (int)A.idx.size();
  }
  return A;
}

void spmv(const Csr& A, const double* x, double* y) {
  for (int i = 0; i < A.rows; i++) {
    double tmp = 0;
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++) tmp += A.val[k] * x[A.idx[k]];
    y[i] = tmp;
  }
}

void fullpivlu_solve(std::vector<double>& A, int n, const std::vector<double>& b, std::vector<double>& x) {
  // FullPivLU::computeInPlace — pivot = first strict maximum of |a| over the trailing block in a
  // column-major scan; swap row, swap column, divide the column by the pivot, rank-1 update.
  std::vector<int> rowT(n), colT(n);
  int nonzero = n;
  double maxpivot = 0;
  for (int k = 0; k < n; k++) {
    int br = k, bc = k;
    double bv = std::fabs(A[k + (size_t)k * n]);
    for (int c = k; c < n; c++)
      for (int r = k; r < n; r++) {
        const double v = std::fabs(A[r + (size_t)c * n]);
        if (v > bv) { bv = v; br = r; bc = c; }
      }
    if (bv == 0) {
      nonzero = k;
      for (int i = k; i < n; i++) rowT[i] = colT[i] = i;
      break;
    }
    if (bv > maxpivot) maxpivot = bv;
    rowT[k] = br; colT[k] = bc;
    if (br != k) for (int c = 0; c < n; c++) std::swap(A[k + (size_t)c * n], A[br + (size_t)c * n]);
    if (bc != k) for (int r = 0; r < n; r++) std::swap(A[r + (size_t)k * n], A[r + (size_t)bc * n]);
    if (k < n - 1) {
      const double piv = A[k + (size_t)k * n];
      for (int r = k + 1; r < n; r++) A[r + (size_t)k * n] /= piv;
      for (int c = k + 1; c < n; c++) {
        const double u = A[k + (size_t)c * n];
        for (int r = k + 1; r < n; r++) A[r + (size_t)c * n] -= A[r + (size_t)k * n] * u;
      }
    }
  }
  // FullPivLU::_solve_impl with rank() at threshold eps*size*|maxpivot|
  const double thr = std::fabs(maxpivot) * (std::numeric_limits<double>::epsilon() * n);
  int rank = 0;
  for (int i = 0; i < nonzero; i++) rank += (std::fabs(A[i + (size_t)i * n]) > thr);
  x.assign(n, 0.0);
  if (rank == 0) return;
  std::vector<double> c(b);
  for (int k = 0; k < n; k++) std::swap(c[k], c[rowT[k]]);  // c = P b
  for (int j = 0; j < n; j++) {                              // unit-lower solve
    const double cj = c[j];
    for (int i = j + 1; i < n; i++) c[i] -= A[i + (size_t)j * n] * cj;
  }
  for (int j = rank - 1; j >= 0; j--) {                      // upper solve on the rank x rank corner
    c[j] /= A[j + (size_t)j * n];
    const double cj = c[j];
    for (int i = 0; i < j; i++) c[i] -= A[i + (size_t)j * n] * cj;
  }
  for (int i = 0; i < rank; i++) x[i] = c[i];
  for (int k = n - 1; k >= 0; k--) std::swap(x[k], x[colT[k]]);  // x = Q y
}

// ---------------------------------------------------------------------------------
// grid.cpp
// ---------------------------------------------------------------------------------
Grid::Grid(std::vector<Pt> points, std::vector<Boundary> boundaries, GridProperties props, std::vector<double> source) {  // :5-27
  const int numPoint = (int)points.size();
  points_ = std::move(points);
  boundaries_ = std::move(boundaries);
  properties_ = props;
  source_ = std::move(source);
  neumannFlag_ = false;
  setNeumannFlag();
  laplaceMatSize_ = numPoint;
  const int A_size = neumannFlag_ ? numPoint + 1 : numPoint;
  bcFlags_.assign(numPoint, 0);
  normalVecs_.assign(numPoint, Pt{0, 0, 0});
  laplaceMat_.rows = laplaceMat_.cols = A_size;
  laplaceMat_.ptr.assign(A_size + 1, 0);
  values_.assign(A_size, 0.0);
  neumann_boundary_coeffs_.rows = neumann_boundary_coeffs_.cols = A_size;
  neumann_boundary_coeffs_.ptr.assign(A_size + 1, 0);
  diags.assign(A_size, 0.0);  // reference leaves this uninitialised; entries it reads are always written first
}

void Grid::setBCFlag(int bNum, const std::string& type, const std::vector<double>& vals) {  // :33-40
  Boundary& bound = boundaries_.at(bNum);
  bound.type = type.compare("dirichlet") == 0 ? 1 : 2;
  for (size_t i = 0; i < bound.bcPoints.size(); i++) bcFlags_[bound.bcPoints[i]] = bound.type;
  bound.values = vals;
}

void Grid::setNeumannFlag() {  // :52-60
  for (const Boundary& b : boundaries_)
    if (b.type == 2) { neumannFlag_ = true; return; }
  neumannFlag_ = false;
}

void Grid::boundaryOp(const std::string& coarse) {  // :42-51
  const bool isCoarse = coarse.compare("coarse") == 0;
  for (const Boundary& b : boundaries_)
    if (b.type == 1)
      for (size_t j = 0; j < b.bcPoints.size(); j++) values_[b.bcPoints.at(j)] = isCoarse ? 0 : b.values.at(j);
}

void Grid::modify_coeff_neumann(const std::string& coarse) {  // :62-72
  const bool isCoarse = coarse.compare("coarse") == 0;
  for (const Boundary& b : boundaries_)
    if (b.type == 2)
      for (size_t j = 0; j < b.bcPoints.size(); j++) source_[b.bcPoints.at(j)] = isCoarse ? 0 : b.values.at(j);
  source_[source_.size() - 1] = 0;
}

void Grid::bound_eval_neumann() {  // :73-103
  const Csr& A = laplaceMat_;
  for (const Boundary& b : boundaries_) {
    if (b.type != 2) continue;
    for (size_t j = 0; j < b.bcPoints.size(); j++) {
      const int curr = b.bcPoints[j];
      double diag = 0;
      double boundValue = source_[curr];
      for (int k = A.ptr[curr]; k < A.ptr[curr + 1]; k++) {
        if (A.idx[k] == curr) { diag = A.val[k]; continue; }
        boundValue -= values_[A.idx[k]] * A.val[k];
      }
      boundValue /= diag;
      values_[curr] = boundValue;
    }
  }
}

void Grid::sor(const Csr& A, std::vector<double>& values, const std::vector<double>& rhs) {  // :104-146
  for (int it = 0; it < properties_.iters; it++) {
    for (int i = 0; i < A.rows; i++) {
      if (!(neumannFlag_ && i == A.rows - 1) && bcFlags_[i] != 0) continue;
      double x_i = 0, diagCoeff = 0;
      for (int j = A.ptr[i]; j < A.ptr[i + 1]; j++) {
        if (A.idx[j] == i) { diagCoeff = A.val[j]; continue; }
        x_i -= A.val[j] * values[A.idx[j]];
      }
      x_i += rhs[i];
      x_i *= properties_.omega / diagCoeff;
      x_i += (1 - properties_.omega) * values[i];
      values[i] = x_i;
    }
    bound_eval_neumann();
  }
}

std::vector<double> Grid::residual() {  // :147-151
  std::vector<double> Ax(laplaceMat_.rows);
  spmv(laplaceMat_, values_.data(), Ax.data());
  std::vector<double> res(laplaceMat_.rows);
  for (int i = 0; i < laplaceMat_.rows; i++) res[i] = source_[i] - Ax[i];
  fix_vector_bound_coarse(res);
  return res;
}

void Grid::fix_vector_bound_coarse(std::vector<double>& v) {  // :197-205
  for (const Boundary& b : boundaries_)
    if (b.type == 1)
      for (size_t j = 0; j < b.bcPoints.size(); j++) v[b.bcPoints.at(j)] = 0;
}

std::vector<int> Grid::kNearestNeighbors(int pointID, bool neumann, int k) {  // :213-215
  return kNearestNeighbors(points_[pointID], neumann, bcFlags_[pointID] != 0, k);
}

std::vector<int> Grid::kNearestNeighbors(const Pt& ref, bool neumann, bool pointBCFlag, int k) {  // :216-260
  if (knn_mode == KNN_CELLS) return knn_cells(ref, neumann, pointBCFlag, k);
  const int N = laplaceMatSize_;
  std::vector<std::pair<double, int>> dist(N);
  int samePoint = -1;
  for (int i = 0; i < N; i++) {
    dist[i] = {distance(ref, points_[i]), i};
    if (dist[i].first == 0) samePoint = i;
  }
  auto excluded = [&](int i) { return pointBCFlag && neumann && bcFlags_[i] != 0; };
  int lastInit = k;
  std::vector<std::pair<double, int>> heap;
  for (int i = 0; i < lastInit; i++) {
    if (i >= N) throw std::runtime_error("kNearestNeighbors: fewer admissible points than stencil size");
    if (i != samePoint && excluded(i)) { lastInit++; continue; }
    heap.push_back(dist[i]);
  }
  std::make_heap(heap.begin(), heap.end());
  for (int i = lastInit; i < N; i++) {
    if (i == samePoint || (!excluded(i) && dist[i] < heap.front())) {
      std::pop_heap(heap.begin(), heap.end());
      heap.pop_back();
      heap.push_back(dist[i]);
      std::push_heap(heap.begin(), heap.end());
    }
  }
  std::sort_heap(heap.begin(), heap.end());
  std::vector<int> nn(k);
  for (int i = 0; i < k; i++) nn[i] = heap[i].second;
  return nn;
}

// Oracle-only accelerator: exact cell-grid search that returns the same list as the brute-force
// selection above (same (distance,index) key, same admission rule); proven equal in tests.
void Grid::build_cells() {
  const int N = laplaceMatSize_;
  double minX = points_[0].x, maxX = minX, minY = points_[0].y, maxY = minY;
  for (const Pt& p : points_) {
    minX = std::min(minX, p.x); maxX = std::max(maxX, p.x);
    minY = std::min(minY, p.y); maxY = std::max(maxY, p.y);
  }
  CellIndex& C = cells_;
  C.x0 = minX; C.y0 = minY;
  const double area = std::max((maxX - minX) * (maxY - minY), 1e-300);
  C.cs = 2.0 * std::sqrt(area / N);
  C.nx = (int)std::floor((maxX - minX) / C.cs) + 1;
  C.ny = (int)std::floor((maxY - minY) / C.cs) + 1;
  C.start.assign((size_t)C.nx * C.ny + 1, 0);
  std::vector<int> cell(N);
  for (int i = 0; i < N; i++) {
    int cx = std::min(C.nx - 1, std::max(0, (int)std::floor((points_[i].x - C.x0) / C.cs)));
    int cy = std::min(C.ny - 1, std::max(0, (int)std::floor((points_[i].y - C.y0) / C.cs)));
    cell[i] = cy * C.nx + cx;
    C.start[cell[i] + 1]++;
  }
  for (size_t c = 0; c < (size_t)C.nx * C.ny; c++) C.start[c + 1] += C.start[c];
  C.ids.assign(N, 0);
  std::vector<int> pos(C.start.begin(), C.start.end() - 1);
  for (int i = 0; i < N; i++) C.ids[pos[cell[i]]++] = i;
  C.valid = true;
}

std::vector<int> Grid::knn_cells(const Pt& ref, bool neumann, bool pointBCFlag, int k) {
  if (!cells_.valid) {
#pragma omp critical(orc_cells)
    if (!cells_.valid) build_cells();
  }
  const CellIndex& C = cells_;
  const int cx = std::min(C.nx - 1, std::max(0, (int)std::floor((ref.x - C.x0) / C.cs)));
  const int cy = std::min(C.ny - 1, std::max(0, (int)std::floor((ref.y - C.y0) / C.cs)));
  // samePoint = last index at distance exactly 0 (identical coordinates share the home cell)
  int samePoint = -1;
  for (int q = C.start[cy * C.nx + cx]; q < C.start[cy * C.nx + cx + 1]; q++) {
    const int i = C.ids[q];
    if (distance(ref, points_[i]) == 0) samePoint = std::max(samePoint, i);
  }
  auto excluded = [&](int i) { return pointBCFlag && neumann && bcFlags_[i] != 0; };
  std::vector<std::pair<double, int>> heap;
  heap.reserve(k + 1);
  auto visit = [&](int ccx, int ccy) {
    for (int q = C.start[ccy * C.nx + ccx]; q < C.start[ccy * C.nx + ccx + 1]; q++) {
      const int i = C.ids[q];
      if (i != samePoint && excluded(i)) continue;
      const std::pair<double, int> key(distance(ref, points_[i]), i);
      if ((int)heap.size() < k) {
        heap.push_back(key);
        std::push_heap(heap.begin(), heap.end());
      } else if (i == samePoint || key < heap.front()) {
        std::pop_heap(heap.begin(), heap.end());
        heap.back() = key;
        std::push_heap(heap.begin(), heap.end());
      }
    }
  };
  const int rmax = std::max(C.nx, C.ny);
  for (int r = 0; r <= rmax; r++) {
    const int xlo = cx - r, xhi = cx + r, ylo = cy - r, yhi = cy + r;
    for (int yy = std::max(ylo, 0); yy <= std::min(yhi, C.ny - 1); yy++)
      for (int xx = std::max(xlo, 0); xx <= std::min(xhi, C.nx - 1); xx++)
        if (r == 0 || yy == ylo || yy == yhi || xx == xlo || xx == xhi) visit(xx, yy);
    // every unvisited point lies outside the visited square of cells
    const double inf = std::numeric_limits<double>::infinity();
    const double sx0 = (xlo <= 0) ? inf : ref.x - (C.x0 + xlo * C.cs);
    const double sx1 = (xhi >= C.nx - 1) ? inf : (C.x0 + (xhi + 1) * C.cs) - ref.x;
    const double sy0 = (ylo <= 0) ? inf : ref.y - (C.y0 + ylo * C.cs);
    const double sy1 = (yhi >= C.ny - 1) ? inf : (C.y0 + (yhi + 1) * C.cs) - ref.y;
    const double safe = std::min(std::min(sx0, sx1), std::min(sy0, sy1));
    if ((int)heap.size() == k && heap.front().first < safe * (1 - 1e-12)) break;
    if (safe == inf) break;
  }
  if ((int)heap.size() < k) throw std::runtime_error("kNearestNeighbors(cells): fewer admissible points than stencil size");
  std::sort_heap(heap.begin(), heap.end());
  std::vector<int> nn(k);
  for (int i = 0; i < k; i++) nn[i] = heap[i].second;
  return nn;
}

void Grid::buildCoeffMatrix(const Pt& point, bool neumann, bool pointBCFlag, int polyDeg, std::vector<double>& M,
                            std::vector<int>& neighbors, std::vector<Pt>& scaledPoints) {  // :263-299
  const int polyTerms = (polyDeg + 1) * (polyDeg + 2) / 2;
  const int stencilSize = (int)(2.5 * (polyDeg + 1) * (polyDeg + 2) / 2);
  neighbors = kNearestNeighbors(point, neumann, pointBCFlag, stencilSize);
  std::vector<Pt> nbPts(neighbors.size());
  for (size_t i = 0; i < neighbors.size(); i++) nbPts[i] = points_[neighbors[i]];
  scaledPoints = shifting_scaling(nbPts, point);
  const int S = stencilSize + polyTerms;
  M.assign((size_t)S * S, 0.0);
  for (int i = 0; i < stencilSize; i++)
    for (int j = i; j < stencilSize; j++) {
      const double r = distance(scaledPoints[i], scaledPoints[j]);
      const double a = std::pow(r, properties_.rbfExp);
      M[i + (size_t)j * S] = a;
      M[j + (size_t)i * S] = a;
    }
  for (int row = 0; row < stencilSize; row++) {
    int col = stencilSize;
    for (int p = 0; p <= polyDeg; p++)
      for (int q = 0; q <= p; q++) {
        const double x = scaledPoints[row].x, y = scaledPoints[row].y;
        const double pc = std::pow(x, p - q) * std::pow(y, q);
        M[row + (size_t)col * S] = pc;
        M[col + (size_t)row * S] = pc;
        col++;
      }
  }
}

std::pair<std::vector<double>, std::vector<int>> Grid::derivx_weights(int pointID) {  // :304-342
  std::vector<double> Mx; std::vector<int> nb; std::vector<Pt> sp;
  buildCoeffMatrix(points_[pointID], neumannFlag_, bcFlags_[pointID] != 0, properties_.polyDeg, Mx, nb, sp);
  const int polyTerms = (properties_.polyDeg + 1) * (properties_.polyDeg + 2) / 2;
  const int n = properties_.stencilSize;
  std::vector<double> rhs(n + polyTerms, 0.0);
  const Pt evalPoint = sp.at(sp.size() - 1);
  const double xEval = evalPoint.x, yEval = evalPoint.y;
  const double M = (double)properties_.rbfExp;
  for (int i = 0; i < n; i++) {
    const double xRef = sp[i].x;
    if (i > 0) rhs[i] = M * std::pow(distance(sp[i], evalPoint), M - 2) * (xEval - xRef);
  }
  int row = n;
  for (int p = 0; p <= properties_.polyDeg; p++)
    for (int q = 0; q <= p; q++) {
      double t = 0;
      if (p - q - 1 >= 0) t += (p - q) * std::pow(xEval, p - q - 1) * std::pow(yEval, q);
      rhs[row++] = t;
    }
  std::vector<double> w;
  fullpivlu_solve(Mx, n + polyTerms, rhs, w);
  const double scale = sp[sp.size() - 2].x;
  for (double& wi : w) wi /= scale;
  return {w, nb};
}

std::pair<std::vector<double>, std::vector<int>> Grid::derivy_weights(int pointID) {  // :343-380
  std::vector<double> Mx; std::vector<int> nb; std::vector<Pt> sp;
  buildCoeffMatrix(points_[pointID], neumannFlag_, bcFlags_[pointID] != 0, properties_.polyDeg, Mx, nb, sp);
  const int polyTerms = (properties_.polyDeg + 1) * (properties_.polyDeg + 2) / 2;
  const int n = properties_.stencilSize;
  std::vector<double> rhs(n + polyTerms, 0.0);
  const Pt evalPoint = sp.at(sp.size() - 1);
  const double xEval = evalPoint.x, yEval = evalPoint.y;
  const double M = (double)properties_.rbfExp;
  for (int i = 0; i < n; i++) {
    const double yRef = sp[i].y;
    if (i > 0) rhs[i] = M * std::pow(distance(sp[i], evalPoint), M - 2) * (yEval - yRef);
  }
  int row = n;
  for (int p = 0; p <= properties_.polyDeg; p++)
    for (int q = 0; q <= p; q++) {
      double t = 0;
      if (q - 1 >= 0) t += q * std::pow(xEval, p - q) * std::pow(yEval, q - 1);
      rhs[row++] = t;
    }
  std::vector<double> w;
  fullpivlu_solve(Mx, n + polyTerms, rhs, w);
  const double scale = sp[sp.size() - 2].x;
  for (double& wi : w) wi /= scale;
  return {w, nb};
}

std::pair<std::vector<double>, std::vector<int>> Grid::laplaceWeights(int pointID) {  // :381-424
  std::vector<double> Mx; std::vector<int> nb; std::vector<Pt> sp;
  buildCoeffMatrix(points_[pointID], neumannFlag_, bcFlags_[pointID] != 0, properties_.polyDeg, Mx, nb, sp);
  const int polyTerms = (properties_.polyDeg + 1) * (properties_.polyDeg + 2) / 2;
  const int n = properties_.stencilSize;
  std::vector<double> rhs(n + polyTerms, 0.0);
  const Pt evalPoint = sp.at(sp.size() - 1);
  const double xEval = evalPoint.x, yEval = evalPoint.y;
  const double M = (double)properties_.rbfExp;
  for (int i = 0; i < n; i++) {
    const double xRef = sp[i].x, yRef = sp[i].y;
    const double D = (xEval * xEval - 2 * xEval * xRef + xRef * xRef + yEval * yEval - 2 * yEval * yRef + yRef * yRef);
    if (D > 0) {
      rhs[i] = (std::pow(2 * xEval - 2 * xRef, 2) + std::pow(2 * yEval - 2 * yRef, 2)) * (M / 2) * (M / 2 - 1) * std::pow(D, M / 2 - 2) +
               2 * M * std::pow(D, M / 2 - 1);
    }
  }
  int row = n;
  for (int p = 0; p <= properties_.polyDeg; p++)
    for (int q = 0; q <= p; q++) {
      double t = 0;
      if (p - q - 2 >= 0) t += (p - q) * (p - q - 1) * std::pow(xEval, p - q - 2) * std::pow(yEval, q);
      if (q - 2 >= 0) t += q * (q - 1) * std::pow(xEval, p - q) * std::pow(yEval, q - 2);
      rhs[row++] = t;
    }
  std::vector<double> w;
  fullpivlu_solve(Mx, n + polyTerms, rhs, w);
  const double scale = sp[sp.size() - 2].x;
  for (double& wi : w) wi /= std::pow(scale, 2);
  return {w, nb};
}

std::pair<std::vector<double>, std::vector<int>> Grid::pointInterpWeights(const Pt& point, int polyDeg) {  // :687-712
  std::vector<double> Mx; std::vector<int> nb; std::vector<Pt> sp;
  buildCoeffMatrix(point, false, false, polyDeg, Mx, nb, sp);
  const int polyTerms = (polyDeg + 1) * (polyDeg + 2) / 2;
  const int stencilSize = (int)(2.5 * polyTerms);
  std::vector<double> rhs(stencilSize + polyTerms, 0.0);
  const Pt evalPoint = sp.at(sp.size() - 1);
  const double xEval = evalPoint.x, yEval = evalPoint.y;
  for (int i = 0; i < stencilSize; i++) rhs[i] = std::pow(distance(evalPoint, sp[i]), properties_.rbfExp);
  int row = stencilSize;
  for (int p = 0; p <= polyDeg; p++)
    for (int q = 0; q <= p; q++) rhs[row++] = std::pow(xEval, p - q) * std::pow(yEval, q);
  std::vector<double> w;
  fullpivlu_solve(Mx, stencilSize + polyTerms, rhs, w);
  return {w, nb};
}

void Grid::build_normal_vecs_square() {  // :442-461 (square branch: inward normals, y tested first)
  const Boundary& b0 = boundaries_[0];
  for (size_t b = 0; b < b0.bcPoints.size(); b++) {
    const Pt& c = points_[b0.bcPoints[b]];
    if (c.y == 0) normalVecs_[b0.bcPoints[b]] = Pt{0, 1, 0};
    else if (c.y == 1) normalVecs_[b0.bcPoints[b]] = Pt{0, -1, 0};
    else if (c.x == 0) normalVecs_[b0.bcPoints[b]] = Pt{1, 0, 0};
    else if (c.x == 1) normalVecs_[b0.bcPoints[b]] = Pt{-1, 0, 0};
  }
}

void Grid::build_normal_vecs(int geom) {  // :442-516 -- the square loop runs for every geomtype, then the circle branches
  build_normal_vecs_square();
  auto radial = [&](const Boundary& bd, double sign) {
    for (size_t b = 0; b < bd.bcPoints.size(); b++) {
      const Pt& c = points_[bd.bcPoints[b]];
      double x = c.x, y = c.y;
      x = x - 0.5; y = y - 0.5;
      const double normPoint = std::sqrt(x * x + y * y);
      x /= normPoint; y /= normPoint;
      normalVecs_[bd.bcPoints[b]] = sign > 0 ? Pt{x, y, 0} : Pt{-x, -y, 0};
    }
  };
  if (geom == GEOM_SQUARE_WITH_CIRCLE) radial(boundaries_.at(1), +1);                       // :480-491
  else if (geom == GEOM_CONCENTRIC_CIRCLES) { radial(boundaries_.at(0), -1); radial(boundaries_.at(1), +1); }   // :492-515
}

void Grid::build_deriv_normal_bound() {  // :520-548
  deriv_normal_coeffs_.clear();
  std::vector<std::pair<int, double>> todo;  // (point, value) in reference visiting order
  for (const Boundary& b : boundaries_)
    if (b.type == 2)
      for (size_t j = 0; j < b.bcPoints.size(); j++) todo.push_back({b.bcPoints[j], b.values[j]});
  deriv_normal_coeffs_.resize(todo.size());
#pragma omp parallel for schedule(dynamic, 16)
  for (long t = 0; t < (long)todo.size(); t++) {
    const int cur = todo[t].first;
    const double xWeight = normalVecs_[cur].x, yWeight = normalVecs_[cur].y;
    auto cx = derivx_weights(cur);
    auto cy = derivy_weights(cur);
    for (size_t i = 0; i < cx.first.size(); i++) {
      cx.first[i] *= xWeight;
      cx.first[i] += yWeight * cy.first[i];
    }
    DerivNormalBC bound;
    bound.pointID = cur;
    bound.value = todo[t].second;
    bound.weights = cx.first;
    bound.neighbors = cx.second;
    deriv_normal_coeffs_[t] = bound;
  }
}

void Grid::build_laplacian() {  // :549-663
  const int N = laplaceMatSize_;
  std::vector<Trip> tripletList, boundaryList;
  std::vector<std::pair<std::vector<double>, std::vector<int>>> W(N);
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = 0; i < N; i++) W[i] = laplaceWeights(i);  // independent per node; emitted in order below
  for (int i = 0; i < N; i++) {
    const auto& weights = W[i];
    if (bcFlags_[i] != 2) {
      for (size_t j = 0; j < weights.second.size(); j++) {
        tripletList.push_back(Trip{i, weights.second[j], weights.first[j]});
        if (bcFlags_[i] == 0 && bcFlags_[weights.second[j]] == 2) boundaryList.push_back(Trip{i, weights.second[j], weights.first[j]});
        if (i == weights.second[j]) diags[i] = weights.first[j];
      }
    }
    if (neumannFlag_ && bcFlags_[i] != 2) tripletList.push_back(Trip{i, N, 1});
  }
  W.clear();
  if (neumannFlag_) {
    for (int i = 0; i < N + 1; i++)
      if (i == N || bcFlags_[i] != 2) tripletList.push_back(Trip{N, i, 1});
    for (const DerivNormalBC& bound : deriv_normal_coeffs_)
      for (size_t j = 0; j < bound.neighbors.size(); j++) {
        tripletList.push_back(Trip{bound.pointID, bound.neighbors[j], bound.weights[j]});
        if (bound.pointID == bound.neighbors[j]) diags[bound.pointID] = bound.weights[j];
      }
  }
  const int A_size = laplaceMat_.rows;
  laplaceMat_ = csr_from_triplets(A_size, A_size, tripletList);
  neumann_boundary_coeffs_ = csr_from_triplets(A_size, A_size, boundaryList);
  if (!implicitFlag_) return;

  // implicit elimination of Neumann boundary unknowns from interior rows (:598-661)
  const Csr& A = laplaceMat_;
  std::vector<std::pair<int, double>> rowBnd;
  for (int i = 0; i < A.rows - 1; i++) {
    if (bcFlags_[i] != 0) continue;
    rowBnd.clear();
    for (int j = A.ptr[i]; j < A.ptr[i + 1]; j++)
      if (A.idx[j] != A.rows - 1 && bcFlags_[A.idx[j]] == 2) rowBnd.push_back({A.idx[j], A.val[j]});
    for (size_t j = 0; j < rowBnd.size(); j++) {
      const int j_col = rowBnd[j].first;
      const double A_ij = rowBnd[j].second;
      const double A_jj = diags[j_col];
      for (int k = A.ptr[j_col]; k < A.ptr[j_col + 1]; k++) {
        const double A_jk = A.val[k];
        if (j_col == A.idx[k]) continue;
        tripletList.push_back(Trip{i, A.idx[k], -A_jk * A_ij / A_jj});
      }
      tripletList.push_back(Trip{i, j_col, -A_ij});
    }
  }
  laplaceMat_ = csr_from_triplets((int)points_.size() + 1, (int)points_.size() + 1, tripletList);
}

void Grid::push_inhomog_to_rhs() {  // :664-685
  if (!implicitFlag_) return;
  const Csr& B = neumann_boundary_coeffs_;
  const std::vector<double> sourceCopy = source_;
  for (int i = 0; i < laplaceMatSize_; i++) {
    if (bcFlags_[i] != 0) continue;
    for (int j = B.ptr[i]; j < B.ptr[i + 1]; j++) {
      const double diag = diags[B.idx[j]];
      const double A_ij = B.val[j];
      source_[i] -= A_ij * sourceCopy[B.idx[j]] / diag;
    }
  }
}

void Grid::rcm_order_points() {  // :713-776
  const int N = (int)points_.size();
  std::vector<std::vector<int>> adjacency(N);
  std::vector<int> order(N);
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < N; i++) adjacency[i] = kNearestNeighbors(points_[i], neumannFlag_, bcFlags_[i] != 0, properties_.stencilSize);
  if (neumannFlag_ && implicitFlag_) {
    for (int i = 0; i < N; i++) {
      if (bcFlags_[i] != 0) continue;
      for (size_t j = 0; j < adjacency[i].size(); j++) {   // size() re-read: appended nodes are visited too
        const int nb = adjacency[i].at(j);
        if (bcFlags_[nb] != 2) continue;
        for (size_t k = 0; k < adjacency[nb].size(); k++) {
          const int cand = adjacency[nb].at(k);
          if (std::find(adjacency[i].begin(), adjacency[i].end(), cand) == adjacency[i].end()) adjacency[i].push_back(cand);
        }
      }
    }
  }
  reverse_cuthill_mckee_ordering(adjacency, order);
  if ((int)order.size() != N) throw std::runtime_error("rcm_order_points: BFS from node 0 did not reach every node (reference reads past the end here)");
  order_ = order;

  std::vector<Pt> newPoints = points_, newNorm = normalVecs_;
  std::vector<double> newSource = source_;
  std::vector<int> newBC = bcFlags_, oldToNew(N);
  for (int i = 0; i < N; i++) {
    newPoints[i] = points_[order[i]];
    newBC[i] = bcFlags_[order[i]];
    newSource[i] = source_[order[i]];
    newNorm[i] = normalVecs_[order[i]];
    oldToNew[order[i]] = i;
  }
  points_ = newPoints; source_ = newSource; bcFlags_ = newBC; normalVecs_ = newNorm;
  for (Boundary& b : boundaries_)
    for (size_t j = 0; j < b.bcPoints.size(); j++) b.bcPoints[j] = oldToNew[b.bcPoints[j]];
  cells_.valid = false;
}

// ---------------------------------------------------------------------------------
// Multicolour mode (oracle restatement of the separately-reported GPU mode; NOT reference code)
// ---------------------------------------------------------------------------------
// Colouring contract (bit-exact integer artefact): rows the smoother visits (bcFlags==0, plus the
// regularisation row of a Neumann grid) are coloured by first-fit in ascending row order on the
// structurally symmetrised graph of laplaceMat_ (i~j iff a_ij or a_ji is stored, explicit zeros
// count); the regularisation row is adjacent to everything and always takes the last colour.
void Grid::build_colouring() {
  const Csr& A = laplaceMat_;
  const int R = A.rows;
  auto swept = [&](int i) { return (neumannFlag_ && i == R - 1) || bcFlags_[i] == 0; };
  const int reg = neumannFlag_ ? R - 1 : -1;
  // transpose structure (without the regularisation row/column)
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
  colour_.assign(R, -1);
  std::vector<int> mark;  // mark[c] == i  <=> colour c used by a neighbour of i
  int ncol = 0;
  for (int i = 0; i < R; i++) {
    if (i == reg || !swept(i)) continue;
    auto touch = [&](int j) {
      if (j < i && j != reg && colour_[j] >= 0) {
        if ((int)mark.size() <= colour_[j]) mark.resize(colour_[j] + 1, -1);
        mark[colour_[j]] = i;
      }
    };
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++) touch(A.idx[k]);
    for (int k = tcnt[i]; k < tcnt[i + 1]; k++) touch(tidx[k]);
    int c = 0;
    while (c < (int)mark.size() && mark[c] == i) c++;
    colour_[i] = c;
    ncol = std::max(ncol, c + 1);
  }
  if (reg >= 0) { colour_[reg] = ncol; ncol++; }
  n_colours_ = ncol;
}

void Grid::sor_multicolour(const Csr& A, std::vector<double>& values, const std::vector<double>& rhs) {
  if (colour_.empty()) build_colouring();
  std::vector<std::vector<int>> rowsOf(n_colours_);
  for (int i = 0; i < A.rows; i++)
    if (colour_[i] >= 0) rowsOf[colour_[i]].push_back(i);
  for (int it = 0; it < properties_.iters; it++) {
    for (int c = 0; c < n_colours_; c++)
      for (int i : rowsOf[c]) {   // same row update as Grid::sor (grid.cpp:122-141)
        double x_i = 0, diagCoeff = 0;
        for (int j = A.ptr[i]; j < A.ptr[i + 1]; j++) {
          if (A.idx[j] == i) { diagCoeff = A.val[j]; continue; }
          x_i -= A.val[j] * values[A.idx[j]];
        }
        x_i += rhs[i];
        x_i *= properties_.omega / diagCoeff;
        x_i += (1 - properties_.omega) * values[i];
        values[i] = x_i;
      }
    bound_eval_neumann();
  }
}

void Grid::build_block_colouring() {
  const Csr& A = laplaceMat_;
  const int R = neumannFlag_ ? A.rows - 1 : A.rows;          // the regularisation row is not part of any block
  const int B = block_size_;
  const int nb = (R + B - 1) / B;
  auto swept = [&](int i) { return i < R && bcFlags_[i] == 0; };
  std::vector<std::vector<int>> adj(nb);
  for (int i = 0; i < R; i++) {
    if (!swept(i)) continue;
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++) {
      const int j = A.idx[k];
      if (j == i || !swept(j)) continue;
      const int a = i / B, b = j / B;
      if (a != b) { adj[a].push_back(b); adj[b].push_back(a); }
    }
  }
  block_colour_.assign(nb, -1);
  n_block_colours_ = 0;
  std::vector<int> mark;
  for (int a = 0; a < nb; a++) {
    for (int b : adj[a])
      if (b < a) {
        if ((int)mark.size() <= block_colour_[b]) mark.resize(block_colour_[b] + 1, -1);
        mark[block_colour_[b]] = a;
      }
    int c = 0;
    while (c < (int)mark.size() && mark[c] == a) c++;
    block_colour_[a] = c;
    n_block_colours_ = std::max(n_block_colours_, c + 1);
  }
}

void Grid::sor_blocklex(const Csr& A, std::vector<double>& values, const std::vector<double>& rhs) {
  if (block_colour_.empty()) build_block_colouring();
  const int R = neumannFlag_ ? A.rows - 1 : A.rows;
  const int B = block_size_, nb = (int)block_colour_.size();
  auto update = [&](int i) {   // same row update as Grid::sor (grid.cpp:122-141)
    double x_i = 0, diagCoeff = 0;
    for (int j = A.ptr[i]; j < A.ptr[i + 1]; j++) {
      if (A.idx[j] == i) { diagCoeff = A.val[j]; continue; }
      x_i -= A.val[j] * values[A.idx[j]];
    }
    x_i += rhs[i];
    x_i *= properties_.omega / diagCoeff;
    x_i += (1 - properties_.omega) * values[i];
    values[i] = x_i;
  };
  for (int it = 0; it < properties_.iters; it++) {
    for (int c = 0; c < n_block_colours_; c++)
      for (int b = 0; b < nb; b++) {
        if (block_colour_[b] != c) continue;
        for (int i = b * B; i < std::min(R, (b + 1) * B); i++)
          if (bcFlags_[i] == 0) update(i);
      }
    if (neumannFlag_) update(A.rows - 1);   // regularisation row last, as in the lexicographic sweep
    bound_eval_neumann();
  }
}

std::vector<int> Grid::lex_levels() const {
  // level(i) = 1 + max level over stored columns j<i that the sweep also visits; skipped rows -1.
  const Csr& A = laplaceMat_;
  const int R = A.rows;
  std::vector<int> lev(R, -1);
  for (int i = 0; i < R; i++) {
    const bool swept = (neumannFlag_ && i == R - 1) || bcFlags_[i] == 0;
    if (!swept) continue;
    int l = 0;
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++) { const int j = A.idx[k]; if (j < i && lev[j] >= 0) l = std::max(l, lev[j] + 1); }
    lev[i] = l;
  }
  return lev;
}

// ---------------------------------------------------------------------------------
// fractionalStepGrid.cpp
// ---------------------------------------------------------------------------------
FractionalStepGrid::FractionalStepGrid(std::vector<Pt> points, std::vector<Boundary> boundaries, GridProperties props, std::vector<double> source)
    : Grid(std::move(points), std::move(boundaries), props, std::move(source)) {  // :2-17
  const int N = laplaceMatSize_;
  u.assign(N, 0); v.assign(N, 0); u_old.assign(N, 0); v_old.assign(N, 0); u_hat.assign(N, 0); v_hat.assign(N, 0);
}

void FractionalStepGrid::set_uv_bound() {  // :41-59
  const double re = rho / mu;
  lambda = 0.5 * re - std::sqrt(0.25 * re * re + 4 * kPi * kPi);
  for (const Boundary& b : boundaries_)
    for (size_t j = 0; j < b.bcPoints.size(); j++) {
      const int c = b.bcPoints.at(j);
      const double x = points_[c].x, y = points_[c].y;
      u[c] = 1 - std::exp(lambda * x) * std::cos(2 * kPi * y);
      v[c] = lambda / (2 * kPi) * std::exp(lambda * x) * std::sin(2 * kPi * y);
      u_old[c] = 1 - std::exp(lambda * x) * std::cos(2 * kPi * y);
      v_old[c] = lambda / (2 * kPi) * std::exp(lambda * x) * std::sin(2 * kPi * y);
    }
}

static Csr build_op(FractionalStepGrid& g, int which) {
  const int N = g.laplaceMatSize_;
  std::vector<std::pair<std::vector<double>, std::vector<int>>> W(N);
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = 0; i < N; i++) W[i] = which == 0 ? g.derivx_weights(i) : which == 1 ? g.derivy_weights(i) : g.laplaceWeights(i);
  std::vector<Trip> t;
  for (int i = 0; i < N; i++)
    for (size_t j = 0; j < W[i].second.size(); j++) t.push_back(Trip{i, W[i].second[j], W[i].first[j]});
  return csr_from_triplets(N, N, t);
}
void FractionalStepGrid::build_derivX_mat() { derivXMat_ = build_op(*this, 0); }      // :60-72
void FractionalStepGrid::build_derivY_mat() { derivYMat_ = build_op(*this, 1); }      // :73-86
void FractionalStepGrid::build_uv_laplace_mat() { uvLaplaceMat_ = build_op(*this, 2); }  // :87-100

void FractionalStepGrid::calc_u_hat() {  // :101-112
  const int N = laplaceMatSize_;
  std::vector<double> u_x(N), u_y(N), del2(N);
  spmv(derivXMat_, u.data(), u_x.data()); spmv(derivYMat_, u.data(), u_y.data()); spmv(uvLaplaceMat_, u.data(), del2.data());
  for (int i = 0; i < N; i++) u_hat[i] = u[i] + dt * (-(u[i] * u_x[i] + v[i] * u_y[i]) + mu / rho * del2[i]);
}
void FractionalStepGrid::calc_v_hat() {  // :113-124
  const int N = laplaceMatSize_;
  std::vector<double> v_x(N), v_y(N), del2(N);
  spmv(derivXMat_, v.data(), v_x.data()); spmv(derivYMat_, v.data(), v_y.data()); spmv(uvLaplaceMat_, v.data(), del2.data());
  for (int i = 0; i < N; i++) v_hat[i] = v[i] + dt * (-(u[i] * v_x[i] + v[i] * v_y[i]) + mu / rho * del2[i]);
}
void FractionalStepGrid::set_ppe_source() {  // :125-145
  const int N = laplaceMatSize_;
  std::vector<double> a(N), b(N);
  spmv(derivXMat_, u_hat.data(), a.data()); spmv(derivYMat_, v_hat.data(), b.data());
  for (int i = 0; i < N; i++) source_[i] = rho / dt * (a[i] + b[i]);
  for (const Boundary& bd : boundaries_)
    for (size_t j = 0; j < bd.bcPoints.size(); j++) {
      const int c = bd.bcPoints[j];
      const double dpdx = -rho / dt * (u[c] - u_hat[c]);
      const double dpdy = -rho / dt * (v[c] - v_hat[c]);
      source_[c] = normalVecs_[c].x * dpdx + normalVecs_[c].y * dpdy;
    }
}
void FractionalStepGrid::correct_u() {  // :146-148
  const int N = laplaceMatSize_;
  std::vector<double> g(N);
  spmv(derivXMat_, values_.data(), g.data());
  for (int i = 0; i < N; i++) u[i] = u_hat[i] - dt / rho * g[i];
}
void FractionalStepGrid::correct_v() {  // :149-151
  const int N = laplaceMatSize_;
  std::vector<double> g(N);
  spmv(derivYMat_, values_.data(), g.data());
  for (int i = 0; i < N; i++) v[i] = v_hat[i] - dt / rho * g[i];
}
double FractionalStepGrid::fs_residual() {  // :152-154
  double s = 0;
  for (int i = 0; i < laplaceMatSize_; i++) s += std::fabs(u[i] - u_hat[i]);
  return s / laplaceMatSize_;
}

// ---------------------------------------------------------------------------------
// multigrid.cpp / FracStepMultigrid.cpp
// ---------------------------------------------------------------------------------
Multigrid::~Multigrid() {
  for (auto& g : grids_) delete g.second;
}

void Multigrid::addGrid(Grid* g) {  // multigrid.cpp:116-122 (sort by (size, pointer))
  grids_.push_back({g->getSize(), g});
  std::sort(grids_.begin(), grids_.end());
}

Csr Multigrid::buildInterpMatrix(Grid* base, Grid* target) {  // multigrid.cpp:17-33; FracStepMultigrid.cpp:17-31
  const int finegridpoly = grids_[grids_.size() - 1].second->properties_.polyDeg;
  const int poly = fracstep ? base->properties_.polyDeg : finegridpoly;
  const int T = target->getSize();
  std::vector<std::pair<std::vector<double>, std::vector<int>>> W(T);
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = 0; i < T; i++) W[i] = base->pointInterpWeights(target->points_[i], poly);
  std::vector<Trip> t;
  for (int i = 0; i < T; i++)
    for (size_t j = 0; j < W[i].second.size(); j++) t.push_back(Trip{i, W[i].second[j], W[i].first[j]});
  // reference stores column-major; the product sums each output row in ascending column order
  // either way, so the oracle keeps rows.
  return csr_from_triplets(T, base->getSize(), t);
}

void Multigrid::buildMatrices() {  // multigrid.cpp:34-60
  const size_t L = grids_.size();
  prolongMatrices_.assign(L, Csr());
  restrictionMatrices_.assign(L, Csr());
  for (size_t i = 0; i + 1 < L; i++) prolongMatrices_[i] = buildInterpMatrix(grids_[i].second, grids_[i + 1].second);
  for (size_t i = 1; i < L; i++) restrictionMatrices_[i] = buildInterpMatrix(grids_[i].second, grids_[i - 1].second);
  for (size_t i = 0; i + 1 < L; i++) grids_[i].second->modify_coeff_neumann("coarse");
}

void Multigrid::smooth(Grid* g) {
  if (blocklex) g->sor_blocklex(g->laplaceMat_, g->values_, g->source_);
  else if (multicolour) g->sor_multicolour(g->laplaceMat_, g->values_, g->source_);
  else g->sor(g->laplaceMat_, g->values_, g->source_);
}

void Multigrid::vCycle() {  // multigrid.cpp:62-110; FracStepMultigrid.cpp:60-112
  const size_t L = grids_.size();
  Grid* currGrid = grids_[L - 1].second;
  if (fracstep && L == 1) { smooth(currGrid); return; }
  const double resid_norm = residual();
  residuals_.push_back(resid_norm);
  currGrid->bound_eval_neumann();
  for (size_t i = L - 1; i > 0; i--) {
    currGrid = grids_[i].second;
    const std::string gridType = (i == L - 1) ? "fine" : "coarse";
    if (i != L - 1) std::fill(currGrid->values_.begin(), currGrid->values_.end(), 0.0);
    currGrid->boundaryOp(gridType);
    smooth(currGrid);
    Grid* coarse = grids_[i - 1].second;
    const std::vector<double> res = currGrid->residual();
    std::vector<double> restricted(coarse->laplaceMatSize_);
    spmv(restrictionMatrices_[i], res.data(), restricted.data());   // only entries 0..N-1 of res are read
    for (int r = 0; r < coarse->laplaceMatSize_; r++) coarse->source_[r] = restricted[r];
    coarse->fix_vector_bound_coarse(coarse->source_);
    if (currGrid->neumannFlag_) {
      coarse->source_[coarse->source_.size() - 1] = 0;
      coarse->modify_coeff_neumann("coarse");
    }
  }
  currGrid->boundaryOp("coarse");   // quirk: still grid 1 (or the only grid) — multigrid.cpp:91
  currGrid = grids_[0].second;
  std::fill(currGrid->values_.begin(), currGrid->values_.end(), 0.0);
  smooth(currGrid);
  smooth(currGrid);
  for (size_t i = 1; i < L; i++) {
    currGrid = grids_[i].second;
    Grid* coarse = grids_[i - 1].second;
    std::vector<double> correction(currGrid->laplaceMatSize_);
    spmv(prolongMatrices_[i - 1], coarse->values_.data(), correction.data());
    if (!currGrid->neumannFlag_) currGrid->fix_vector_bound_coarse(correction);
    for (int r = 0; r < currGrid->laplaceMatSize_; r++) currGrid->values_[r] += correction[r];
    smooth(currGrid);
  }
}

double Multigrid::residual() {  // multigrid.cpp:112-115 (lpNorm<1>; Eigen's packet reduction order is not restated)
  Grid* fine = grids_[grids_.size() - 1].second;
  const std::vector<double> r = fine->residual();
  double num = 0, den = 0;
  for (double t : r) num += std::fabs(t);
  for (double t : fine->source_) den += std::fabs(t);
  return num / den;
}

// ---------------------------------------------------------------------------------
// Problem factories
// ---------------------------------------------------------------------------------
namespace {
// the manufactured right-hand side of the annulus, u = sin(pi k1 r*) with r* = (r - 0.25) / 0.25, written exactly as
// testing_functions.cpp:109-121 / :208-221 write it
double concentric_source(double x, double y, int k1) {
  const double pi = kPi;
  x -= 0.5; y -= 0.5;
  double sum = 0;
  const double r = std::sqrt(x * x + y * y);
  const double rstar = (r - 0.25) / (0.5 - 0.25);
  sum += -pi * k1 * k1 * pi * std::sin(pi * k1 * rstar) * std::pow(4 * x * std::pow(x * x + y * y, -0.5), 2)
         + pi * k1 * std::cos(pi * k1 * rstar) * 4 * (std::pow(x * x + y * y, -0.5) + 2 * x * x * -0.5 * std::pow(x * x + y * y, -1.5));
  sum += -pi * k1 * k1 * pi * std::sin(pi * k1 * rstar) * std::pow(4 * y * std::pow(x * x + y * y, -0.5), 2)
         + pi * k1 * std::cos(pi * k1 * rstar) * 4 * (std::pow(x * x + y * y, -0.5) + 2 * y * y * -0.5 * std::pow(x * x + y * y, -1.5));
  return sum;
}
bool on_circle(double x, double y, double r2) { return std::abs(r2 - (x - 0.5) * (x - 0.5) - (y - 0.5) * (y - 0.5)) <= std::pow(10, -10); }
}  // namespace

Grid* genGridDirichlet(const std::vector<Pt>& points, GridProperties props, int k1, int k2, KnnMode mode, int geom) {  // testing_functions.cpp:68-159
  const double pi = kPi;
  std::vector<int> bPts, bPts_inner;
  std::vector<double> bValues, bValues_inner;
  std::vector<double> source(points.size());
  for (size_t i = 0; i < points.size(); i++) {
    const double x = points[i].x, y = points[i].y;
    if (geom == GEOM_SQUARE) {
      source[i] = -(k1 * k1 + k2 * k2) * pi * pi * std::sin(k1 * pi * x) * std::sin(k2 * pi * y);
      if (x == 0 || x == 1 || y == 0 || y == 1) { bPts.push_back((int)i); bValues.push_back(0.0); }
    } else if (geom == GEOM_SQUARE_WITH_CIRCLE) {   // :92-106 (k1 in both factors, as written)
      source[i] = -(k1 * k1 + k2 * k2) * pi * pi * std::sin(k1 * pi * x) * std::sin(k1 * pi * y);
      if (x == 0 || x == 1 || y == 0 || y == 1) { bPts.push_back((int)i); bValues.push_back(0); }
      else if (on_circle(x, y, 0.0625)) { bPts_inner.push_back((int)i); bValues_inner.push_back(std::sin(k1 * pi * x) * std::sin(k1 * pi * y)); }
    } else {                                        // :107-135
      source[i] = concentric_source(x, y, k1);
      if (on_circle(x, y, 0.25)) { bPts.push_back((int)i); bValues.push_back(0.0); }
      else if (on_circle(x, y, 0.0625)) { bPts_inner.push_back((int)i); bValues_inner.push_back(0.0); }
    }
  }
  Boundary boundary;
  boundary.bcPoints = bPts; boundary.type = 1; boundary.values = bValues;
  std::vector<Boundary> bcs{boundary};
  if (geom != GEOM_SQUARE) {
    Boundary inner;
    inner.bcPoints = bPts_inner; inner.type = 1; inner.values = bValues_inner;
    bcs.push_back(inner);
  }
  Grid* grid = new Grid(points, bcs, props, source);
  grid->knn_mode = mode;
  grid->implicitFlag_ = false;
  grid->setBCFlag(0, "dirichlet", bValues);
  if (geom != GEOM_SQUARE) grid->setBCFlag(1, "dirichlet", bValues_inner);
  grid->rcm_order_points();
  grid->build_laplacian();
  return grid;
}
Grid* genGridDirichletSquare(const std::vector<Pt>& points, GridProperties props, int k1, int k2, KnnMode mode) {
  return genGridDirichlet(points, props, k1, k2, mode, GEOM_SQUARE);
}

Grid* genGridNeumann(const std::vector<Pt>& points, GridProperties props, int k1, int k2, const std::string& coarse, KnnMode mode, int geom) {  // testing_functions.cpp:161-284
  const double pi = kPi;
  std::vector<int> bPts, bPts_inner;
  std::vector<double> bValues, bValues_inner;
  std::vector<double> source(points.size() + 1);
  for (size_t i = 0; i < points.size(); i++) {
    const double x = points[i].x, y = points[i].y;
    if (geom == GEOM_SQUARE) {
      source[i] = -(k1 * k1 + k2 * k2) * pi * pi * std::cos(k1 * pi * x) * std::cos(k2 * pi * y);
      if (x == 0 || x == 1 || y == 0 || y == 1) { bPts.push_back((int)i); bValues.push_back(0.0); }
    } else if (geom == GEOM_SQUARE_WITH_CIRCLE) {   // :186-207
      source[i] = -(k1 * k1 + k2 * k2) * pi * pi * std::cos(k1 * pi * x) * std::cos(k2 * pi * y);
      if (x == 0 || x == 1 || y == 0 || y == 1) { bPts.push_back((int)i); bValues.push_back(0); }
      else if (on_circle(x, y, 0.0625)) {
        double normX = x - 0.5, normY = y - 0.5;
        const double norm = std::sqrt(normX * normX + normY * normY);
        normX /= norm; normY /= norm;
        bPts_inner.push_back((int)i);
        bValues_inner.push_back(-normX * pi * k1 * std::sin(k1 * pi * x) * std::cos(k2 * pi * y) - normY * pi * k2 * std::cos(k1 * pi * x) * std::sin(k2 * pi * y));
      }
    } else {                                        // :208-251
      source[i] = concentric_source(x, y, k1);
      const double xs = x - 0.5, ys = y - 0.5;
      const double r = std::sqrt(xs * xs + ys * ys);
      const double rstar = (r - 0.25) / (0.5 - 0.25);
      if (on_circle(x, y, 0.25)) {
        double normX = x - 0.5, normY = y - 0.5;
        const double norm = std::sqrt(normX * normX + normY * normY);
        normX /= norm; normY /= norm;
        bPts.push_back((int)i);
        bValues.push_back(-normX * k1 * pi * std::cos(k1 * pi * rstar) / r * 4 * (x - 0.5) - normY * k1 * pi * std::cos(k1 * pi * rstar) / r * 4 * (y - 0.5));
      } else if (on_circle(x, y, 0.0625)) {
        double normX = x - 0.5, normY = y - 0.5;
        const double norm = std::sqrt(normX * normX + normY * normY);
        normX /= norm; normY /= norm;
        bPts_inner.push_back((int)i);
        bValues_inner.push_back(normX * k1 * pi * std::cos(k1 * pi * rstar) / r * 4 * (x - 0.5) + normY * k1 * pi * std::cos(k1 * pi * rstar) / r * 4 * (y - 0.5));
      }
    }
  }
  source[source.size() - 1] = 0;
  Boundary boundary;
  boundary.bcPoints = bPts; boundary.type = 2; boundary.values = bValues;
  std::vector<Boundary> bcs{boundary};
  if (geom != GEOM_SQUARE) {
    Boundary inner;
    inner.bcPoints = bPts_inner; inner.type = 2; inner.values = bValues_inner;
    bcs.push_back(inner);
  }
  Grid* grid = new Grid(points, bcs, props, source);
  grid->knn_mode = mode;
  grid->implicitFlag_ = true;
  grid->setBCFlag(0, "neumann", bValues);
  if (geom != GEOM_SQUARE) grid->setBCFlag(1, "neumann", bValues_inner);
  grid->build_normal_vecs(geom);
  grid->rcm_order_points();
  grid->build_deriv_normal_bound();
  grid->build_laplacian();
  grid->modify_coeff_neumann(coarse);
  grid->push_inhomog_to_rhs();
  return grid;
}
Grid* genGridNeumannSquare(const std::vector<Pt>& points, GridProperties props, int k1, int k2, const std::string& coarse, KnnMode mode) {
  return genGridNeumann(points, props, k1, k2, coarse, mode, GEOM_SQUARE);
}

FractionalStepGrid* genFractionalStepGrid(const std::vector<Pt>& points, GridProperties props, double dt, double mu, double rho, double ppe_conv,
                                          const std::string& coarse, KnnMode mode) {  // FractionalStepSim.cpp:3-49
  std::vector<int> bPts;
  std::vector<double> bValues;
  std::vector<double> source(points.size() + 1, 0.0);
  const double re = rho / mu;
  const double lambda = 0.5 * re - std::sqrt(0.25 * re * re + 4 * kPi * kPi);
  for (size_t i = 0; i < points.size(); i++) {
    const double x = points[i].x, y = points[i].y;
    if (x == 0 || x == 1 || y == 0 || y == 1) { bPts.push_back((int)i); bValues.push_back(0.5 * std::exp(2 * lambda * x)); }
  }
  Boundary boundary;
  boundary.bcPoints = bPts; boundary.type = 2; boundary.values = bValues;
  FractionalStepGrid* grid = new FractionalStepGrid(points, {boundary}, props, source);
  grid->knn_mode = mode;
  grid->mu = mu; grid->rho = rho; grid->ppe_conv_res = ppe_conv; grid->dt = dt;
  grid->implicitFlag_ = true;
  grid->setBCFlag(0, "neumann", bValues);
  grid->build_normal_vecs_square();
  grid->rcm_order_points();
  grid->build_deriv_normal_bound();
  grid->build_laplacian();
  grid->modify_coeff_neumann(coarse);
  grid->build_derivX_mat();
  grid->build_derivY_mat();
  grid->build_uv_laplace_mat();
  grid->push_inhomog_to_rhs();
  return grid;
}

Grid* genGridMixedSquare(const std::vector<Pt>& points, GridProperties props, int k1, int k2, const std::string& coarse, KnnMode mode) {
  // u = sin(k1 pi x) cos(k2 pi y): u=0 on x in {0,1} (Dirichlet), du/dn=0 on y in {0,1} (Neumann).
  // Set-up order follows genGmshGridNeumann (testing_functions.cpp:254-283) with two Boundary objects.
  const double pi = kPi;
  std::vector<int> dPts, nPts;
  std::vector<double> dVals, nVals;
  std::vector<double> source(points.size() + 1);
  for (size_t i = 0; i < points.size(); i++) {
    const double x = points[i].x, y = points[i].y;
    source[i] = -(k1 * k1 + k2 * k2) * pi * pi * std::sin(k1 * pi * x) * std::cos(k2 * pi * y);
    if (x == 0 || x == 1) { dPts.push_back((int)i); dVals.push_back(0.0); }
    else if (y == 0 || y == 1) { nPts.push_back((int)i); nVals.push_back(0.0); }
  }
  source[source.size() - 1] = 0;
  // build_normal_vecs visits boundaries_[0] only (grid.cpp:445), so the Neumann boundary goes first.
  Boundary nb, db;
  nb.bcPoints = nPts; nb.type = 2; nb.values = nVals;
  db.bcPoints = dPts; db.type = 1; db.values = dVals;
  Grid* grid = new Grid(points, {nb, db}, props, source);
  grid->knn_mode = mode;
  grid->implicitFlag_ = true;
  grid->setBCFlag(0, "neumann", nVals);
  grid->setBCFlag(1, "dirichlet", dVals);
  grid->build_normal_vecs_square();
  grid->rcm_order_points();
  grid->build_deriv_normal_bound();
  grid->build_laplacian();
  grid->modify_coeff_neumann(coarse);
  grid->push_inhomog_to_rhs();
  return grid;
}

}  // namespace orc
