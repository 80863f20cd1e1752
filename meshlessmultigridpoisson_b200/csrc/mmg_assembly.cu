// RBF-FD assembly on the device (kNN, batched full-pivot LU, CSR build).  Stage-1 placeholder:
// every entry reports MMG_ERR_STATE until the device assembly lands; the solve path works on
// uploaded operators (mmg_grid_set_laplacian_csr / mmg_solver_set_interp_csr).
#include "mmg_internal.hpp"
namespace mmg {
static void nyi(const char* what) { throw Error(MMG_ERR_STATE, std::string(what) + ": device assembly not built into this libmmg yet"); }
void asm_release(Grid&) {}
void asm_knn_points(Grid&, int, const double*, const double*, const int*, int, int, int*) { nyi("kNearestNeighbors"); }
void asm_rcm_order_points(Grid&) { nyi("rcm_order_points"); }
void asm_build_deriv_normal_bound(Grid&) { nyi("build_deriv_normal_bound"); }
void asm_build_laplacian(Grid&) { nyi("build_laplacian"); }
void asm_weights(Grid&, int, int, const int*, double*, int*) { nyi("laplaceWeights"); }
void asm_point_interp_weights(Grid&, int, const double*, const double*, int, double*, int*) { nyi("pointInterpWeights"); }
void asm_build_interp(Grid&, Grid&, int, HybMatrix&) { nyi("buildInterpMatrix"); }
}  // namespace mmg
