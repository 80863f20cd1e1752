// ============================================================================
// TEST INFRASTRUCTURE ONLY — CPU oracle for the multigrid solve path.
//
// This is a plain C++17 restatement (no Eigen) of the reference
// MeshlessPoisson/{grid,multigrid,FracStepMultigrid,general_computation_functions}.cpp.
// It is the checker for the CUDA path: only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it.  The product
// (meshlessmultigridpoisson_b200/) never links, imports or executes it.
//
// Parity status: the reference ships NO golden vectors.  The
// oracle is pinned two ways: (1) against the reference's own sources compiled
// in place against an in-repo Eigen-subset shim (oracle/_ref, built by
// oracle/Makefile) — bit-exact agreement is asserted in tests/; (2) against the
// analytic known-answer properties the reference implies (polynomial
// reproduction of every weight set, partition of unity of interpolation rows,
// manufactured solutions).  The arithmetic that lives inside Eigen (FullPivLU,
// setFromTriplets, sparse*dense) is restated from Eigen 3.4.0 semantics; Eigen
// itself is absent from this image, so that part is "parity unpinned" against
// real Eigen (blocked triangular solves would differ at cond*eps level).
//
// Build: g++ -O2 -ffp-contract=off (MSVC never contracts; grid.cpp:397 relies on it).
// ============================================================================
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace orc {

struct Pt { double x, y, z; };

// general_computation_functions.cpp:4-6
double distance(const Pt& a, const Pt& b);
// general_computation_functions.cpp:82-107 — returns scaled neighbours, then the
// sentinel (scale,scale,scale), then the scaled evaluation point.
std::vector<Pt> shifting_scaling(const std::vector<Pt>& pts, const Pt& eval);
// general_computation_functions.cpp:108-134 — BFS from node 0 (no degree sort), reversed.
void reverse_cuthill_mckee_ordering(const std::vector<std::vector<int>>& adjacency, std::vector<int>& order);

// --- Eigen semantics restated -------------------------------------------------
struct Trip { int r, c; double v; };
struct Csr {                      // compressed rows, columns ascending, explicit zeros kept
  int rows = 0, cols = 0;
  std::vector<int> ptr, idx;
  std::vector<double> val;
};
// Eigen::SparseMatrix::setFromTriplets: duplicates summed in triplet order, inner indices sorted.
Csr csr_from_triplets(int rows, int cols, const std::vector<Trip>& t);
// Eigen row-major sparse * dense vector: tmp=0; tmp += a_ij*x_j ascending j; y_i = tmp.
void spmv(const Csr& A, const double* x, double* y);
// Eigen::FullPivLU<MatrixXd>(A).solve(b), A column-major n x n (destroyed).
void fullpivlu_solve(std::vector<double>& A, int n, const std::vector<double>& b, std::vector<double>& x);

// --- gridclasses.hpp:6-28 --------------------------------------------------------
struct GridProperties { int rbfExp = 3, polyDeg = 3, laplaceMatSize = 0, stencilSize = 25; double omega = 1.4; int iters = 5; };
struct Boundary { int type = 0; std::vector<int> bcPoints; std::vector<double> values; };
struct DerivNormalBC { int pointID; std::vector<double> weights; std::vector<int> neighbors; double value; };

enum KnnMode { KNN_BRUTE = 0, KNN_CELLS = 1 };
// the geomtype strings of the reference's factories: "square", "square_with_circle", "concentric_circles"
enum Geom { GEOM_SQUARE = 0, GEOM_SQUARE_WITH_CIRCLE = 1, GEOM_CONCENTRIC_CIRCLES = 2 };

// grid.h:20-79
struct Grid {
  std::vector<double> values_;     // A entries
  std::vector<double> source_;     // A entries (N for Dirichlet grids)
  std::vector<Pt> points_;
  std::vector<Boundary> boundaries_;
  std::vector<Pt> normalVecs_;
  std::vector<DerivNormalBC> deriv_normal_coeffs_;
  GridProperties properties_;
  int laplaceMatSize_ = 0;
  Csr laplaceMat_;
  Csr neumann_boundary_coeffs_;
  std::vector<double> diags;
  std::vector<int> bcFlags_;
  bool neumannFlag_ = false;
  bool implicitFlag_ = false;
  std::vector<int> order_;          // permutation applied by rcm_order_points (new -> old); oracle artefact

  // oracle-only knobs (do not change results; see tests/test_oracle_knn.py)
  KnnMode knn_mode = KNN_BRUTE;
  struct CellIndex { double x0, y0, cs; int nx, ny; std::vector<int> start, ids; bool valid = false; } cells_;

  Grid(std::vector<Pt> points, std::vector<Boundary> boundaries, GridProperties props, std::vector<double> source);
  virtual ~Grid() {}

  void setBCFlag(int bNum, const std::string& type, const std::vector<double>& vals);  // grid.cpp:33-40
  void setNeumannFlag();                                                              // grid.cpp:52-60
  void boundaryOp(const std::string& coarse);                                         // grid.cpp:42-51
  void modify_coeff_neumann(const std::string& coarse);                               // grid.cpp:62-72
  void bound_eval_neumann();                                                          // grid.cpp:73-103
  void sor(const Csr& A, std::vector<double>& values, const std::vector<double>& rhs);// grid.cpp:104-146
  std::vector<double> residual();                                                     // grid.cpp:147-151
  void fix_vector_bound_coarse(std::vector<double>& v);                               // grid.cpp:197-205
  std::vector<int> kNearestNeighbors(int pointID, bool neumann, int k);               // grid.cpp:213-215
  std::vector<int> kNearestNeighbors(const Pt& p, bool neumann, bool pointBCFlag, int k); // grid.cpp:216-260
  // returns column-major (n+m)^2 matrix, neighbour ids, scaled points            // grid.cpp:263-303
  void buildCoeffMatrix(const Pt& p, bool neumann, bool pointBCFlag, int polyDeg,
                        std::vector<double>& M, std::vector<int>& nb, std::vector<Pt>& sp);
  std::pair<std::vector<double>, std::vector<int>> derivx_weights(int pointID);       // grid.cpp:304-342
  std::pair<std::vector<double>, std::vector<int>> derivy_weights(int pointID);       // grid.cpp:343-380
  std::pair<std::vector<double>, std::vector<int>> laplaceWeights(int pointID);       // grid.cpp:381-424
  std::pair<std::vector<double>, std::vector<int>> pointInterpWeights(const Pt& p, int polyDeg); // grid.cpp:687-712
  void build_normal_vecs_square();                                                    // grid.cpp:442-461
  void build_normal_vecs(int geom);                                                   // grid.cpp:442-516 (circle geometries: analytic radial normals)
  void build_deriv_normal_bound();                                                    // grid.cpp:520-548
  void build_laplacian();                                                             // grid.cpp:549-663
  void push_inhomog_to_rhs();                                                         // grid.cpp:664-685
  void rcm_order_points();                                                            // grid.cpp:713-776
  int getSize() const { return laplaceMatSize_; }

  // ---- oracle restatement of the separately-reported multicolour smoother -----
  // (NOT in the reference; restated here so the CUDA multicolour mode has a checker.)
  std::vector<int> colour_;         // per row 0..A-1, -1 for rows the smoother skips
  int n_colours_ = 0;
  void build_colouring();
  void sor_multicolour(const Csr& A, std::vector<double>& values, const std::vector<double>& rhs);
  // dependency-DAG level sets of one lexicographic sweep (integer artefact)
  std::vector<int> lex_levels() const;
  // ---- oracle restatement of the block-lexicographic smoother (NOT in the reference) -----
  // rows are cut into contiguous blocks of block_size_ rows; blocks are coloured first-fit in ascending block
  // order on the symmetrised block graph; a sweep visits colours in order and, inside a block, rows in ascending
  // order with the reference's row update.  Same-colour blocks are independent, so the GPU runs them concurrently.
  int block_size_ = 4096;
  std::vector<int> block_colour_;
  int n_block_colours_ = 0;
  void build_block_colouring();
  void sor_blocklex(const Csr& A, std::vector<double>& values, const std::vector<double>& rhs);

 private:
  void build_cells();
  std::vector<int> knn_cells(const Pt& p, bool neumann, bool pointBCFlag, int k);
};

// fractionalStepGrid.hpp:4-30 — only what the pressure-Poisson path needs plus the
// explicit operators of §8f rank 1-2.
struct FractionalStepGrid : Grid {
  double dt = 0, ppe_conv_res = 0, rho = 1, mu = 1, lambda = 0;
  std::vector<double> u, v, u_old, v_old, u_hat, v_hat;
  Csr derivXMat_, derivYMat_, uvLaplaceMat_;
  FractionalStepGrid(std::vector<Pt> points, std::vector<Boundary> boundaries, GridProperties props, std::vector<double> source);
  void set_uv_bound();          // fractionalStepGrid.cpp:41-59 (kovasznay)
  void build_derivX_mat();      // :60-72
  void build_derivY_mat();      // :73-86
  void build_uv_laplace_mat();  // :87-100
  void calc_u_hat();            // :101-112
  void calc_v_hat();            // :113-124
  void set_ppe_source();        // :125-145
  void correct_u();             // :146-148
  void correct_v();             // :149-151
  double fs_residual();         // :152-154
};

// multigrid.h:4-23 and FracStepMultigrid.hpp:4-25 (flavour flag selects the twin's differences)
struct Multigrid {
  bool fracstep = false;                 // FracStepMultigrid.cpp:23 (interp polyDeg) and :64-67 (1-grid shortcut)
  bool multicolour = false;              // oracle restatement of the multicolour mode
  bool blocklex = false;                 // oracle restatement of the block-lexicographic mode
  std::vector<std::pair<int, Grid*>> grids_;
  std::vector<Csr> restrictionMatrices_; // [i] : N_{i-1} x N_i, i>=1
  std::vector<Csr> prolongMatrices_;     // [i] : N_{i+1} x N_i, i<L-1
  std::vector<double> residuals_;
  ~Multigrid();
  void addGrid(Grid* g);                 // multigrid.cpp:116-122
  Csr buildInterpMatrix(Grid* base, Grid* target);  // multigrid.cpp:17-33 / FracStepMultigrid.cpp:17-31
  void buildMatrices();                  // multigrid.cpp:49-60
  void vCycle();                         // multigrid.cpp:62-110 / FracStepMultigrid.cpp:60-112
  double residual();                     // multigrid.cpp:112-115
 private:
  void smooth(Grid* g);
};

// ---- problem factories (the callers the oracle must mirror) -----------------------
// testing_functions.cpp:68-159 (square branch)
Grid* genGridDirichletSquare(const std::vector<Pt>& points, GridProperties props, int k1, int k2, KnnMode mode);
// testing_functions.cpp:68-159 / :161-284 with the hole and annulus branches (two boundaries, analytic normals)
Grid* genGridDirichlet(const std::vector<Pt>& points, GridProperties props, int k1, int k2, KnnMode mode, int geom);
Grid* genGridNeumann(const std::vector<Pt>& points, GridProperties props, int k1, int k2, const std::string& coarse, KnnMode mode, int geom);
// testing_functions.cpp:161-284 (square branch)
Grid* genGridNeumannSquare(const std::vector<Pt>& points, GridProperties props, int k1, int k2, const std::string& coarse, KnnMode mode);
// FractionalStepSim.cpp:3-49
FractionalStepGrid* genFractionalStepGrid(const std::vector<Pt>& points, GridProperties props, double dt, double mu, double rho,
                                          double ppe_conv, const std::string& coarse, KnnMode mode);
// Mixed BC (BASELINE config 3): x in {0,1} Dirichlet (boundary 0), y in {0,1} (non-corner) Neumann (boundary 1),
// assembled with the reference's own per-boundary semantics (no reference factory exists for it).
Grid* genGridMixedSquare(const std::vector<Pt>& points, GridProperties props, int k1, int k2, const std::string& coarse, KnnMode mode);

}  // namespace orc
