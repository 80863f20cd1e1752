// Example driver over the facade: genFractionalStepGrid (FractionalStepSim.cpp:3-49) and the time loop of
// run_fracstep_param (FractionalStepSim.cpp:114-148) for Kovasznay flow, with the statements of the reference kept.
// The pressure-Poisson solve inside the loop is the V-cycle hot path; the explicit operators around it run on the
// device too (include/mmg.h, FractionalStepGrid section).
//
//   run_fracstep <timesteps> <ppe_conv_res> <fine_polyDeg> <coarsest.msh> ... <finest.msh>
//
// Prints one line per time step: "<|resid - oldresid|> <V-cycles used>" with 17 digits, then the mean |u - u_exact|.
// The reference's loop has no cap on the inner while; the facade keeps that, so pick sizes where the all-Neumann V-cycle
// converges or bound the run with max_cycles_per_step (5th optional knob, environment MMG_FS_MAX_CYCLES).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "mmg_facade.hpp"

using namespace mmgf;
using namespace mmgf_io;
static const double PI = 3.141592653589793238462643383279502884;   // EIGEN_PI as a double

static FractionalStepGrid* genFractionalStepGrid(const char* filename, GridProperties props, double dt, double mu, double rho, double ppe_conv, std::string coarse) {
  std::vector<Point> points = pointsFromMshFile(filename);
  std::vector<int> bPts;
  std::vector<double> bValues, source(points.size() + 1, 0.0);
  const double re = rho / mu;
  const double lambda = 0.5 * re - std::sqrt(0.25 * re * re + 4 * PI * PI);
  for (size_t i = 0; i < points.size(); i++) {
    const double x = std::get<0>(points[i]), y = std::get<1>(points[i]);
    if (x == 0 || x == 1 || y == 0 || y == 1) { bPts.push_back((int)i); bValues.push_back(0.5 * std::exp(2 * lambda * x)); }
  }
  Boundary boundary;
  boundary.bcPoints = bPts; boundary.type = 2; boundary.values = bValues;
  FractionalStepGrid* grid = new FractionalStepGrid(points, {boundary}, props, source);
  grid->mu = mu;
  grid->rho = rho;
  grid->ppe_conv_res = ppe_conv;
  grid->dt = dt;
  grid->implicitFlag_ = true;
  grid->flowType = "kovasznay";
  grid->setBCFlag(0, std::string("neumann"), bValues);
  grid->build_normal_vecs(filename, "square");
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

int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: %s <timesteps> <ppe_conv_res> <fine_polyDeg> <coarsest.msh> ... <finest.msh>\n", argv[0]); return 2; }
  try {
    const int steps = atoi(argv[1]), poly_deg = atoi(argv[3]), numGrids = argc - 4;
    const double ppe_conv = atof(argv[2]), dt = 0.0002, mu = 0.025, rho = 1;     // run_frac_step_test, FractionalStepSim.cpp:202
    const char* cap = getenv("MMG_FS_MAX_CYCLES");
    const long max_cycles = cap ? atol(cap) : -1;
    FractionalStepMultigrid mg;
    for (int i = 0; i < numGrids; i++) {                                          // gen_fracstep_param, :60-67
      GridProperties p;
      p.iters = 5; p.polyDeg = (i == numGrids - 1) ? poly_deg : 3; p.omega = 1.4; p.rbfExp = 3;
      p.stencilSize = (int)(2.5 * (p.polyDeg + 1) * (p.polyDeg + 2) / 2);
      mg.addGrid(genFractionalStepGrid(argv[4 + i], p, dt, mu, rho, ppe_conv, i == numGrids - 1 ? "fine" : "coarse"));
    }
    mg.buildMatrices();
    FractionalStepGrid* finestGrid = mg.grids_[mg.grids_.size() - 1].second;
    double resid = 100, oldresid = 1000;
    for (int timesteps = 0; timesteps < steps; timesteps++) {
      *(finestGrid->u_old) = *(finestGrid->u);
      *(finestGrid->v_old) = *(finestGrid->v);
      finestGrid->set_uv_bound();
      finestGrid->calc_u_hat();
      finestGrid->calc_v_hat();
      finestGrid->set_ppe_source();
      finestGrid->push_inhomog_to_rhs();
      long cycles = 0;
      while (mg.residual() >= ppe_conv && (max_cycles < 0 || cycles < max_cycles)) {
        mg.vCycle();
        finestGrid->bound_eval_neumann();
        cycles++;
      }
      finestGrid->correct_u();
      finestGrid->correct_v();
      finestGrid->set_uv_bound();
      resid = finestGrid->fs_residual();
      printf("%.17g %ld\n", std::fabs(resid - oldresid), cycles);
      oldresid = resid;
    }
    const double re = rho / mu, lambda = 0.5 * re - std::sqrt(0.25 * re * re + 4 * PI * PI);
    double err = 0;
    for (int i = 0; i < finestGrid->laplaceMatSize_; i++) {
      const double x = std::get<0>(finestGrid->points_[i]), y = std::get<1>(finestGrid->points_[i]);
      err += std::fabs(1 - std::exp(lambda * x) * std::cos(2 * PI * y) - finestGrid->u->coeff(i));
    }
    printf("l1_error_u %.17g\n", err / finestGrid->laplaceMatSize_);
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
