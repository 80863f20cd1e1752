// TEST INFRASTRUCTURE -- drop-in proof.  This translation unit is linked with the reference's UNMODIFIED testing_functions.cpp,
// FractionalStepSim.cpp, fileReadingFunctions.cpp and general_computation_functions.cpp, compiled where they lie against
// meshlessmultigridpoisson_b200/cpp/dropin/ (grid.h, multigrid.h, gridclasses.hpp, fractionalStepGrid.hpp, FracStepMultigrid.hpp
// forwarding to the facade over libmmg).  Every Grid / Multigrid / FractionalStepMultigrid call those drivers make therefore
// runs on the GPU.  Built by oracle/Makefile (target dropin) into oracle/_ref/dropin_driver; used by tests/test_gpu_dropin.py.
//
//   dropin_driver mg  <geomtype> <neumann 0|1> <directory/> <extension> <cycles> <finePoly> <file0> <file1> ...
//       run_mg_sim(params) exactly as the reference runs it (testing_functions.cpp:328-350; writes resid_/temp_/x_/y_/cond_error_
//       files at the reference's 6 significant digits), then the same factories + loop again, printing the residual history and
//       the solution with 17 digits for the 1e-10 comparison with oracle/_ref.
//   dropin_driver fs  <directory/> <steps> <finePoly> <file0> <file1> ...
//       genFractionalStepGrid + the time-loop body of run_fracstep_param (FractionalStepSim.cpp:130-147) for <steps> steps (the
//       reference hard-codes 2001), printing u, v, p with 17 digits.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "FractionalStepSim.hpp"
#include "testing_functions.hpp"

static GridProperties props_for(int polyDeg) {   // gen_mg_param, testing_functions.cpp:372-380
  GridProperties p;
  p.iters = 5; p.polyDeg = polyDeg; p.omega = 1.4; p.rbfExp = 3;
  p.stencilSize = (int)(2.5 * (p.polyDeg + 1) * (p.polyDeg + 2) / 2);
  return p;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const std::string mode = argv[1];
  try {
    if (mode == "mg") {
      MultigridParameters params;
      params.geomtype = argv[2];
      params.neumann = atoi(argv[3]) != 0;
      params.directory = argv[4];
      params.extension = argv[5];
      params.num_v_cycle = atoi(argv[6]);
      const int finePoly = atoi(argv[7]);
      params.k1 = 1; params.k2 = 1;
      const int nfiles = argc - 8;
      for (int i = 0; i < nfiles; i++) {
        params.filenames.push_back(argv[8 + i]);
        params.filetypes.push_back("msh");
        params.props.push_back(props_for(i == nfiles - 1 ? finePoly : 3));
      }
      run_mg_sim(params);                                     // the reference's own driver, unmodified
      Multigrid mg;                                           // and once more for full-precision output
      for (int i = 0; i < nfiles; i++) {
        const std::string coarse = (i == nfiles - 1) ? "fine" : "coarse";
        const std::string fn = params.directory + params.filenames[i];
        if (params.neumann) mg.addGrid(genGmshGridNeumann(params.geomtype, fn.c_str(), params.props[i], "msh", 1, 1, coarse));
        else mg.addGrid(genGmshGridDirichlet(params.geomtype, fn.c_str(), params.props[i], "msh", 1, 1));
      }
      mg.buildMatrices();
      for (int i = 0; i < params.num_v_cycle; i++) mg.vCycle();
      printf("HISTORY");
      for (double r : mg.residuals_) printf(" %.17g", r);
      printf("\nVALUES");
      Grid* fine = mg.grids_.back().second;
      for (int i = 0; i < (int)fine->values_->rows(); i++) printf(" %.17g", fine->values_->coeff(i));
      printf("\n");
    } else if (mode == "fs") {
      const std::string dir = argv[2];
      const int steps = atoi(argv[3]), finePoly = atoi(argv[4]);
      const int nfiles = argc - 5;
      FractionalStepMultigrid mg;
      for (int i = 0; i < nfiles; i++) {
        const std::string coarse = (i == nfiles - 1) ? "fine" : "coarse";
        mg.addGrid(genFractionalStepGrid((dir + argv[5 + i]).c_str(), props_for(i == nfiles - 1 ? finePoly : 3), 2e-4, 0.025, 1.0, 1e-10, coarse));
      }
      mg.buildMatrices();
      FractionalStepGrid* finestGrid = mg.grids_[mg.grids_.size() - 1].second;
      for (int t = 0; t < steps; t++) {                       // FractionalStepSim.cpp:130-147
        *(finestGrid->u_old) = *(finestGrid->u);
        *(finestGrid->v_old) = *(finestGrid->v);
        finestGrid->set_uv_bound();
        finestGrid->calc_u_hat();
        finestGrid->calc_v_hat();
        finestGrid->set_ppe_source();
        finestGrid->push_inhomog_to_rhs();
        int guard = 0;
        while (mg.residual() >= 1e-10 && guard++ < 400) {
          mg.vCycle();
          finestGrid->bound_eval_neumann();
        }
        finestGrid->correct_u();
        finestGrid->correct_v();
        finestGrid->set_uv_bound();
        printf("STEP %d cycles %d fs_residual %.17g\n", t, guard, finestGrid->fs_residual());
      }
      printf("U");
      for (int i = 0; i < finestGrid->laplaceMatSize_; i++) printf(" %.17g", finestGrid->u->coeff(i));
      printf("\nP");
      for (int i = 0; i < finestGrid->laplaceMatSize_; i++) printf(" %.17g", finestGrid->values_->coeff(i));
      printf("\n");
    } else return 2;
  } catch (const std::exception& e) {
    fprintf(stderr, "dropin_driver: %s\n", e.what());
    return 1;
  }
  return 0;
}
