/*
 * mmg.h — C-ABI of the B200-native meshless multigrid Poisson solve path.
 *
 * Drop-in boundary for the reference's C++ surface (michaelxu3/MeshlessMultigridPoisson,
 * paths relative to MeshlessPoisson/):
 *     class Grid                      grid.h:20-79
 *     class Multigrid                 multigrid.h:4-23
 *     class FractionalStepMultigrid   FracStepMultigrid.hpp:4-25
 * The reference has no FFI; its boundary is those classes' public methods and members.  Every
 * entry point below names the method/member it replaces.  The C++ facade
 * (meshlessmultigridpoisson_b200/cpp/mmg_facade.hpp) re-creates the three classes with the
 * reference's method names on top of this header; INTEGRATION.md shows the binding.
 *
 * Conventions: plain pointers and sizes; all host pointers unless a name says `dev`; fp64 values,
 * int32 indices; every function returns MMG_OK (0) or an error code and records a message
 * retrievable with mmg_last_error().  There is NO CPU fallback: without a CUDA device (or when a
 * kernel fails) calls return MMG_ERR_CUDA.
 */
#ifndef MMG_H
#define MMG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mmg_grid mmg_grid;     /* one level: replaces Grid / FractionalStepGrid */
typedef struct mmg_solver mmg_solver; /* the cycle:  replaces Multigrid / FractionalStepMultigrid */

/* GridProperties, gridclasses.hpp:6-14 (laplaceMatSize is dead in the reference and omitted) */
typedef struct mmg_props {
  int rbfExp;
  int polyDeg;
  int stencilSize;
  int iters;
  double omega;
} mmg_props;

enum { MMG_OK = 0, MMG_ERR_ARG = 1, MMG_ERR_CUDA = 2, MMG_ERR_STATE = 3, MMG_ERR_NCCL = 4, MMG_ERR_TIMEOUT = 5 };
enum { MMG_BC_DIRICHLET = 1, MMG_BC_NEUMANN = 2 };                  /* Boundary::type, grid.cpp:35 */
enum { MMG_GEOM_SQUARE = 0, MMG_GEOM_SQUARE_WITH_CIRCLE = 1, MMG_GEOM_CONCENTRIC_CIRCLES = 2 };  /* the geomtype strings, testing_functions.cpp:81,92,107 */
enum { MMG_FINE = 0, MMG_COARSE = 1 };                              /* the "fine"/"coarse" strings, grid.cpp:47,67 */
enum { MMG_SMOOTHER_LEXICOGRAPHIC = 0, MMG_SMOOTHER_MULTICOLOUR = 1, MMG_SMOOTHER_BLOCK_LEXICOGRAPHIC = 2 };
enum { MMG_FLAVOUR_MULTIGRID = 0, MMG_FLAVOUR_FRACSTEP = 1 };
/* row sums folded in the reference's ascending-column order (bit-faithful, default) or with reordered warp reductions (throughput) */
enum { MMG_ARITH_REFERENCE_ORDER = 0, MMG_ARITH_FAST = 1 };
/* matrices addressable through the CSR getters/setters */
enum { MMG_MAT_LAPLACE = 0, MMG_MAT_NEUMANN_COEFFS = 1, MMG_MAT_RESTRICT = 2, MMG_MAT_PROLONG = 3,
       MMG_MAT_DERIVX = 4, MMG_MAT_DERIVY = 5, MMG_MAT_UVLAPLACE = 6 };
/* per-kernel-class timers (mmg_solver_get_timers) */
enum { MMG_T_SOR = 0, MMG_T_RESIDUAL = 1, MMG_T_RESTRICT = 2, MMG_T_PROLONG = 3, MMG_T_OTHER = 4, MMG_T_COUNT = 5 };

const char* mmg_last_error(void);
int mmg_device_count(int* count);
/* library build identity: "sm_100a" etc., so callers can assert the native path is the one loaded */
const char* mmg_build_info(void);

/* ---------------------------------------------------------------- Grid ---------------------- */
/* Grid::Grid(points, boundaries, properties, source)  grid.cpp:5-27.
 * Boundaries arrive the way the drivers build them (testing_functions.cpp:136-141): b_type[i] already
 * set (so the ctor's setNeumannFlag sees it), points of boundary i = b_points[b_ptr[i]..b_ptr[i+1]),
 * same slice of b_values.  source has source_len entries (n for Dirichlet grids, n+1 when any boundary
 * is Neumann, testing_functions.cpp:172-173). */
int mmg_grid_create(mmg_grid** out, int device, int n, const double* x, const double* y, const mmg_props* props,
                    const double* source, int source_len, int n_boundaries, const int* b_type, const int* b_ptr,
                    const int* b_points, const double* b_values);
int mmg_grid_destroy(mmg_grid* g);                                                 /* Grid::~Grid grid.cpp:28-31 */
int mmg_grid_set_implicit(mmg_grid* g, int flag);                                  /* Grid::implicitFlag_ grid.h:38 */
int mmg_grid_set_bc_flag(mmg_grid* g, int boundary, int type, const double* values, int n_values); /* Grid::setBCFlag grid.cpp:33-40 */
int mmg_grid_build_normal_vecs_square(mmg_grid* g);                                /* Grid::build_normal_vecs(.., "square") grid.cpp:442-461 */
int mmg_grid_build_normal_vecs(mmg_grid* g, int geomtype);                          /* Grid::build_normal_vecs(.., geomtype) grid.cpp:442-516, MMG_GEOM_* */
int mmg_grid_set_normal_vecs(mmg_grid* g, const double* nx, const double* ny);     /* Grid::normalVecs_ grid.h:28 (other geometries) */
int mmg_grid_rcm_order_points(mmg_grid* g);                                        /* Grid::rcm_order_points grid.cpp:713-776 */
int mmg_grid_build_deriv_normal_bound(mmg_grid* g);                                /* Grid::build_deriv_normal_bound grid.cpp:520-548 */
int mmg_grid_build_laplacian(mmg_grid* g);                                         /* Grid::build_laplacian grid.cpp:549-663 */
int mmg_grid_modify_coeff_neumann(mmg_grid* g, int coarse);                        /* Grid::modify_coeff_neumann grid.cpp:62-72 */
int mmg_grid_push_inhomog_to_rhs(mmg_grid* g);                                     /* Grid::push_inhomog_to_rhs grid.cpp:664-685 */
int mmg_grid_boundary_op(mmg_grid* g, int coarse);                                 /* Grid::boundaryOp grid.cpp:42-51 */
int mmg_grid_bound_eval_neumann(mmg_grid* g);                                      /* Grid::bound_eval_neumann grid.cpp:73-103 */
int mmg_grid_set_props(mmg_grid* g, const mmg_props* props);                       /* Grid::properties_ (omega, iters) grid.h:30 */
int mmg_grid_set_arithmetic(mmg_grid* g, int arithmetic);                          /* MMG_ARITH_*: no reference counterpart (the reference has one order) */
int mmg_grid_sor(mmg_grid* g, int smoother);                                       /* Grid::sor(laplaceMat_, values_, &source_) grid.cpp:104-146 */
int mmg_grid_residual(mmg_grid* g, double* out);                                   /* Grid::residual grid.cpp:147-151 (A entries) */
int mmg_grid_fix_vector_bound_coarse(mmg_grid* g, double* vec);                    /* Grid::fix_vector_bound_coarse grid.cpp:197-205 */
/* Grid::kNearestNeighbors(point, neumann, pointBCFlag, k) grid.cpp:216-260 for m query points */
int mmg_grid_knn(mmg_grid* g, int m, const double* qx, const double* qy, const int* q_bcflag, int neumann, int k, int* out);
/* Grid::laplaceWeights / derivx_weights / derivy_weights grid.cpp:304-424 for node ids; which = MMG_MAT_LAPLACE|DERIVX|DERIVY;
 * w receives stencilSize entries per node (the entries the reference keeps), nb the neighbour ids */
int mmg_grid_weights(mmg_grid* g, int which, int m, const int* ids, double* w, int* nb);
/* Grid::pointInterpWeights(point, polyDeg) grid.cpp:687-712 for m points */
int mmg_grid_point_interp_weights(mmg_grid* g, int m, const double* px, const double* py, int polyDeg, double* w, int* nb);

/* public data members the drivers touch (grid.h:23-38); vectors are A = n or n+1 long */
int mmg_grid_sizes(mmg_grid* g, int* n, int* a_size, int* neumann_flag);           /* laplaceMatSize_, laplaceMat_->rows(), neumannFlag_ */
int mmg_grid_get_values(mmg_grid* g, double* out);                                 /* *values_ */
int mmg_grid_set_values(mmg_grid* g, const double* in);
int mmg_grid_get_source(mmg_grid* g, double* out);                                 /* source_ */
int mmg_grid_set_source(mmg_grid* g, const double* in);
int mmg_grid_get_points(mmg_grid* g, double* x, double* y);                        /* points_ */
int mmg_grid_get_bcflags(mmg_grid* g, int* flags);                                 /* bcFlags_ */
int mmg_grid_get_normals(mmg_grid* g, double* nx, double* ny);                     /* normalVecs_ */
int mmg_grid_get_diags(mmg_grid* g, double* out);                                  /* diags */
int mmg_grid_get_perm(mmg_grid* g, int* order);                                    /* the order applied by rcm_order_points (new -> old) */
int mmg_grid_get_boundary(mmg_grid* g, int boundary, int* type, int* count, int* points, double* values); /* boundaries_[b] */
/* laplaceMat_ / neumann_boundary_coeffs_ as Eigen stores them: compressed rows, columns ascending */
int mmg_grid_csr_nnz(mmg_grid* g, int which, int64_t* nnz);
int mmg_grid_get_csr(mmg_grid* g, int which, int* ptr, int* idx, double* val);
/* upload path: operators built elsewhere.  Replaces what build_laplacian leaves behind: laplaceMat_, diags,
 * neumann_boundary_coeffs_ (nb_* may be NULL for grids without implicit Neumann elimination). */
int mmg_grid_set_laplacian_csr(mmg_grid* g, int rows, const int* ptr, const int* idx, const double* val, const double* diags,
                               const int* nb_ptr, const int* nb_idx, const double* nb_val);
/* integer artefacts of the GPU schedules (bit-exact against the oracle) */
/* y = M x on the device, host vectors in and out (M one of MMG_MAT_LAPLACE / DERIVX / DERIVY / UVLAPLACE): the products the reference's
 * drivers form with laplaceMat_, derivXMat_, derivYMat_, uvLaplaceMat_ (FractionalStepSim.cpp:80-113) */
int mmg_grid_apply_matrix(mmg_grid* g, int which, const double* x, double* y);
int mmg_grid_get_colouring(mmg_grid* g, int* n_colours, int* colour);              /* per row, -1 for rows the sweep skips */
int mmg_grid_get_colour_counts(mmg_grid* g, int* n_colours, int* counts, int cap);  /* rows per colour class (one multicolour launch each) */
int mmg_grid_set_block_size(mmg_grid* g, int rows_per_block);                      /* block-lexicographic smoother: rows per block (default 4096) */
int mmg_grid_get_block_colouring(mmg_grid* g, int* n_blocks, int* n_colours, int* colour, int cap); /* colour of each block (bit-exact vs oracle) */
int mmg_grid_get_lex_levels(mmg_grid* g, int* n_levels, int* level);               /* dependency-DAG level of each row, -1 if skipped */

/* ---------------------------------------------------------------- FractionalStepGrid (fractionalStepGrid.hpp:4-30) -------- */
enum { MMG_FS_U = 0, MMG_FS_V = 1, MMG_FS_U_OLD = 2, MMG_FS_V_OLD = 3, MMG_FS_U_HAT = 4, MMG_FS_V_HAT = 5, MMG_FS_BOTH = -1 };
int mmg_grid_fs_init(mmg_grid* g, double dt, double mu, double rho);               /* FractionalStepGrid ctor (fractionalStepGrid.cpp:2-17) + dt/mu/rho */
int mmg_grid_fs_build_operators(mmg_grid* g);                                      /* build_derivX_mat, build_derivY_mat, build_uv_laplace_mat :60-100 */
int mmg_grid_fs_set_operator_csr(mmg_grid* g, int which, int rows, const int* ptr, const int* idx, const double* val); /* upload path for derivXMat_/derivYMat_/uvLaplaceMat_ */
int mmg_grid_fs_get_vec(mmg_grid* g, int which, double* out);                      /* u, v, u_old, v_old, u_hat, v_hat (N entries) */
int mmg_grid_fs_set_vec(mmg_grid* g, int which, const double* in);
int mmg_grid_fs_scatter(mmg_grid* g, int which, int count, const int* idx, const double* vals); /* the boundary writes of set_uv_bound :41-59 (values evaluated by the caller's libm) */
int mmg_grid_fs_set_uv_bound(mmg_grid* g);                                         /* set_uv_bound :41-59 (kovasznay), host libm like the reference */
int mmg_grid_fs_calc_hat(mmg_grid* g, int component);                              /* calc_u_hat (MMG_FS_U) :101-112, calc_v_hat (MMG_FS_V) :113-124, or MMG_FS_BOTH */
int mmg_grid_fs_set_ppe_source(mmg_grid* g);                                       /* set_ppe_source :125-145 */
int mmg_grid_fs_correct(mmg_grid* g, int component);                               /* correct_u :146-148, correct_v :149-151, or MMG_FS_BOTH */
int mmg_grid_fs_residual(mmg_grid* g, double* out);                                /* fs_residual :152-154 */

/* ---------------------------------------------------------------- Multigrid ----------------- */
int mmg_solver_create(mmg_solver** out, int flavour);                              /* Multigrid::Multigrid / FractionalStepMultigrid */
int mmg_solver_destroy(mmg_solver* s);                                             /* ~Multigrid (owns its grids) multigrid.cpp:10-16 */
int mmg_solver_add_grid(mmg_solver* s, mmg_grid* g);                               /* Multigrid::addGrid multigrid.cpp:116-122 (takes ownership, keeps sorted by size) */
int mmg_solver_num_grids(mmg_solver* s, int* n);
int mmg_solver_grid(mmg_solver* s, int level, mmg_grid** g);                       /* grids_[level].second (0 = coarsest) */
int mmg_solver_build_matrices(mmg_solver* s);                                      /* Multigrid::buildMatrices multigrid.cpp:49-60 (device assembly of P and R) */
/* upload path for restrictionMatrices_[level] / prolongMatrices_[level]; which = MMG_MAT_RESTRICT|PROLONG; rows given compressed by row */
int mmg_solver_set_interp_csr(mmg_solver* s, int which, int level, int rows, int cols, const int* ptr, const int* idx, const double* val);
int mmg_solver_interp_nnz(mmg_solver* s, int which, int level, int* rows, int* cols, int64_t* nnz);
int mmg_solver_get_interp_csr(mmg_solver* s, int which, int level, int* ptr, int* idx, double* val);
int mmg_solver_finish_build(mmg_solver* s);                                        /* tail of buildMatrices: modify_coeff_neumann("coarse") on non-finest grids, multigrid.cpp:54-59 */
int mmg_solver_set_smoother(mmg_solver* s, int smoother);                          /* lexicographic (reference-faithful) | multicolour */
/* the individual statements of vCycle, for per-operator parity (level i >= 1 unless noted) */
int mmg_solver_restrict(mmg_solver* s, int level);                                 /* source_{i-1} = R_i * residual_i + masks, multigrid.cpp:81-86 */
int mmg_solver_prolong_correct(mmg_solver* s, int level);                          /* values_i += P_{i-1} * values_{i-1}, multigrid.cpp:102-106 */
int mmg_solver_coarse_solve(mmg_solver* s);                                        /* coarsest level: zero guess + 2 sor calls, multigrid.cpp:92-95 */
int mmg_solver_set_block_size(mmg_solver* s, int rows_per_block);                  /* block-lexicographic smoother, every grid */
int mmg_solver_set_omega(mmg_solver* s, double omega);                             /* properties_.omega of every grid (gridclasses.hpp:12) */
int mmg_solver_set_arithmetic(mmg_solver* s, int arithmetic);                      /* MMG_ARITH_* for every grid of the solver */
int mmg_solver_vcycle(mmg_solver* s, int n_cycles);                                /* Multigrid::vCycle multigrid.cpp:62-110, n times, no host sync inside */
int mmg_solver_residual(mmg_solver* s, double* out);                               /* Multigrid::residual multigrid.cpp:112-115 */
int mmg_solver_history_len(mmg_solver* s, int* n);                                 /* residuals_.size() */
int mmg_solver_get_history(mmg_solver* s, double* out, int cap);                   /* residuals_ */
/* while (residual() >= tol) vCycle();   loop shape of FractionalStepSim.cpp:139-142; extra_bound_eval=1 adds the
 * finestGrid->bound_eval_neumann() of :141 after every cycle */
int mmg_solver_solve(mmg_solver* s, double tol, int max_cycles, int extra_bound_eval, int* cycles_done, double* final_residual);
int mmg_solver_sync(mmg_solver* s);
/* measurement hooks: CUDA-event time (ms) and launch counts accumulated per kernel class since the last reset */
int mmg_solver_enable_timers(mmg_solver* s, int on);
int mmg_solver_get_timers(mmg_solver* s, int level, double* ms, int64_t* launches, int64_t* bytes); /* level -1 = all levels; arrays of MMG_T_COUNT */
int mmg_solver_reset_timers(mmg_solver* s);
int mmg_solver_launch_count(mmg_solver* s, int64_t* launches);                     /* kernels launched by this solver so far */
/* CUDA-event timed V-cycles on the solver's stream: runs n cycles, returns elapsed device ms */
int mmg_solver_time_vcycles(mmg_solver* s, int n_cycles, double* ms);

/* ---------------------------------------------------------------- multi-GPU (no reference counterpart: the reference is serial) ---
 * One process per GPU.  Levels with at least `threshold` rows are cut into contiguous row blocks in the reference order
 * (rank r owns rows [bounds[r], bounds[r+1])); each rank works on its block only.  Residual, restriction and prolongation
 * exchange the index ranges their rows read with grouped ncclSend/ncclRecv; the smoother needs no exchange step: its
 * barrier-free sweep stores the rows next to a cut into the neighbour rank's vectors over NVLink peer memory (CUDA IPC,
 * handles shipped through NCCL; falls back to per-colour NCCL exchanges if IPC is unavailable).  Smaller levels are
 * replicated.  Needs the multicolour smoother with fast arithmetic.  After init_comm, vcycle / solve / residual are
 * collective calls: every rank must make them in the same order. */
int mmg_partition_bounds(int n, int world, int* bounds);                           /* the partition map (world+1 offsets), pure host */
int mmg_comm_unique_id(char* out128);                                              /* ncclGetUniqueId on rank 0; ship the 128 bytes to every rank */
int mmg_solver_init_comm(mmg_solver* s, int rank, int world, const char* id128);   /* ncclCommInitRank on the solver's device */
int mmg_solver_set_partition_threshold(mmg_solver* s, int rows);
int mmg_solver_comm_stats(mmg_solver* s, int64_t* messages, int64_t* bytes_sent, int* partitioned_levels);
/* After a partitioned vcycle / solve, mmg_grid_get_values on a rank is current on that rank's row block and halo ranges only.
 * This collective completes values_ of every partitioned level on every rank (call it before reading the whole solution). */
int mmg_solver_gather_values(mmg_solver* s);
/* rows [own_lo, own_hi) of `level` belong to this rank; its kernels read values_ in [need_lo, need_hi) (row block + halo).
 * With the ranged copies below a rank moves only that part of Grid::values_ / source_ between host and device. */
int mmg_solver_owned_range(mmg_solver* s, int level, int* own_lo, int* own_hi, int* need_lo, int* need_hi);
int mmg_grid_get_values_range(mmg_grid* g, int offset, int count, double* out);   /* values_[offset, offset+count) */
int mmg_grid_set_values_range(mmg_grid* g, int offset, int count, const double* in);
int mmg_grid_set_source_range(mmg_grid* g, int offset, int count, const double* in); /* source_[offset, offset+count) */

/* ---------------------------------------------------------------- diagnostics (no reference counterpart) ---
 * Used by the test-suite to prove which kernel instantiation ran and to check host-side schedules. */
int mmg_debug_last_kernel(int slot, char* out, int cap);                           /* slot 0: last smoother kernel, 1: last SpMV-class kernel (this thread) */
int mmg_debug_lex_trace(long long* out, int n);                                    /* clock64 stamps of the chunked lexicographic kernel (MMG_LEX_TRACE=1) */
int mmg_debug_exchange_plan(int rank, int world, const int* need, const int* bounds, int* n_send, int* sends, int* n_recv, int* recvs); /* pure host */

#ifdef __cplusplus
}
#endif
#endif /* MMG_H */
