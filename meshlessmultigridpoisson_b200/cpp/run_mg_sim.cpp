// Example driver over the facade: the body of run_mg_sim (testing_functions.cpp:328-350) with the Dirichlet
// square factory of genGmshGridDirichlet (testing_functions.cpp:68-159), reading Gmsh $Nodes files like
// pointsFromMshFile (fileReadingFunctions.cpp:6-32).  Prints the residual history (residuals_) with 17 digits.
//
//   run_mg_sim <num_v_cycle> <fine_polyDeg> <coarsest.msh> ... <finest.msh>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "mmg_facade.hpp"

using namespace mmgf;
using namespace mmgf_io;
static const double pi = 3.141592653589793238462643383279;   // testing_functions.hpp:9

static Grid* genGridDirichlet(const char* filename, GridProperties props, int k1, int k2) {
  std::vector<Point> points = pointsFromMshFile(filename);
  std::vector<int> bPts;
  std::vector<double> bValues, source(points.size());
  for (size_t i = 0; i < points.size(); i++) {
    const double x = std::get<0>(points[i]), y = std::get<1>(points[i]);
    source[i] = -(k1 * k1 + k2 * k2) * pi * pi * std::sin(k1 * pi * x) * std::sin(k2 * pi * y);
    if (x == 0 || x == 1 || y == 0 || y == 1) { bPts.push_back((int)i); bValues.push_back(0.0); }
  }
  Boundary boundary;
  boundary.bcPoints = bPts; boundary.type = 1; boundary.values = bValues;
  Grid* grid = new Grid(points, {boundary}, props, source);
  grid->implicitFlag_ = false;
  grid->setBCFlag(0, std::string("dirichlet"), bValues);
  grid->rcm_order_points();
  grid->build_laplacian();
  return grid;
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s <num_v_cycle> <fine_polyDeg> <coarsest.msh> ... <finest.msh>\n", argv[0]); return 2; }
  try {
    const int num_v_cycle = atoi(argv[1]), poly_deg = atoi(argv[2]), numGrids = argc - 3;
    Multigrid mg;
    for (int i = 0; i < numGrids; i++) {      // gen_mg_param, testing_functions.cpp:372-380
      GridProperties p;
      p.iters = 5; p.polyDeg = (i == numGrids - 1) ? poly_deg : 3; p.omega = 1.4; p.rbfExp = 3;
      p.stencilSize = (int)(2.5 * (p.polyDeg + 1) * (p.polyDeg + 2) / 2);
      mg.addGrid(genGridDirichlet(argv[3 + i], p, 1, 1));
    }
    mg.buildMatrices();
    for (int i = 0; i < num_v_cycle; i++) mg.vCycle();
    for (double r : mg.residuals_) printf("%.17g\n", r);
    Grid* fine = mg.grids_.back().second;
    double err = 0;
    for (int i = 0; i < fine->laplaceMatSize_; i++)   // calc_l1_error, testing_functions.cpp:3-16
      err += std::fabs((*fine->values_)(i) - std::sin(pi * std::get<0>(fine->points_[i])) * std::sin(pi * std::get<1>(fine->points_[i])));
    printf("l1_error %.17g\n", err / fine->laplaceMatSize_);
    if (getenv("MMG_PRINT_INTERP")) {                 // restrictionMatrices_ / prolongMatrices_: shapes, entries, largest |row sum - 1|
      for (int l = 0; l < numGrids; l++)
        for (int which = 0; which < 2; which++) {
          if ((which == 0 && l == 0) || (which == 1 && l == numGrids - 1)) continue;
          const Multigrid::InterpCsr m = which == 0 ? mg.restrictionMatrix(l) : mg.prolongMatrix(l);
          double dev = 0;
          for (int r = 0; r < m.rows; r++) {
            double sum = 0;
            for (int k = m.ptr[r]; k < m.ptr[r + 1]; k++) sum += m.val[k];
            dev = std::fmax(dev, std::fabs(sum - 1.0));
          }
          printf("interp %s %d %d %d %d %.3g\n", which == 0 ? "R" : "P", l, m.rows, m.cols, m.ptr[m.rows], dev);
        }
    }
    if (const char* dir = getenv("MMG_OUT_DIR")) {    // write_mg_resid + write_temp_contour, testing_functions.cpp:346-349
      const std::string extension = std::to_string(numGrids) + "grid__L=" + std::to_string(poly_deg);
      write_mg_resid(mg, std::string(dir) + "/", extension);
      write_temp_contour(fine, std::string(dir) + "/", extension);
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
