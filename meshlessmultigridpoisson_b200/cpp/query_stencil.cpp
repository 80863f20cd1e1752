// Example driver over the facade's per-point queries (grid.h:60-72): for node ids of one Dirichlet level read from a Gmsh
// $Nodes file it prints Grid::kNearestNeighbors, Grid::laplaceWeights / derivx_weights / derivy_weights and, for the midpoint
// of the node and its nearest neighbour, Grid::pointInterpWeights -- 17 digits, one record per line, for a comparison with
// the ctypes mirror (tests/test_gpu_facade.py).
//
//   query_stencil <polyDeg> <level.msh> <id> [<id> ...]
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "mmg_facade.hpp"

using namespace mmgf;
using namespace mmgf_io;

static double at(const DenseVector& v, int i) {
#ifdef MMG_FACADE_HAVE_EIGEN
  return v(i);
#else
  return v[i];
#endif
}

static void print_weights(const char* tag, int id, const std::pair<DenseVector, std::vector<int>>& w) {
  printf("%s %d", tag, id);
  for (size_t k = 0; k < w.second.size(); k++) printf(" %d:%.17g", w.second[k], at(w.first, (int)k));
  printf("\n");
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s <polyDeg> <level.msh> <id> [<id> ...]\n", argv[0]); return 2; }
  try {
    GridProperties p;
    p.iters = 5; p.polyDeg = atoi(argv[1]); p.omega = 1.4; p.rbfExp = 3;
    p.stencilSize = (int)(2.5 * (p.polyDeg + 1) * (p.polyDeg + 2) / 2);
    std::vector<Point> points = pointsFromMshFile(argv[2]);
    std::vector<int> bPts;
    std::vector<double> bValues, source(points.size(), 0.0);
    for (size_t i = 0; i < points.size(); i++) {
      const double x = std::get<0>(points[i]), y = std::get<1>(points[i]);
      if (x == 0 || x == 1 || y == 0 || y == 1) { bPts.push_back((int)i); bValues.push_back(0.0); }
    }
    Boundary boundary;
    boundary.bcPoints = bPts; boundary.type = 1; boundary.values = bValues;
    Grid grid(points, {boundary}, p, source);
    grid.implicitFlag_ = false;
    grid.setBCFlag(0, std::string("dirichlet"), bValues);
    grid.rcm_order_points();
    grid.build_laplacian();
    const DenseVector d = grid.diags();
    for (int a = 3; a < argc; a++) {
      const int id = atoi(argv[a]);
      const std::vector<int> nb = grid.kNearestNeighbors(id, grid.neumannFlag_, p.stencilSize);
      printf("knn %d", id);
      for (int j : nb) printf(" %d", j);
      printf("\n");
      print_weights("laplace", id, grid.laplaceWeights(id));
      print_weights("derivx", id, grid.derivx_weights(id));
      print_weights("derivy", id, grid.derivy_weights(id));
      const std::vector<Point> two = grid.pointIDs_to_vector({id, nb.at(1)});
      const Point mid(0.5 * (std::get<0>(two[0]) + std::get<0>(two[1])), 0.5 * (std::get<1>(two[0]) + std::get<1>(two[1])), 0.0);
      print_weights("interp", id, grid.pointInterpWeights(mid, p.polyDeg));
      printf("diag %d %.17g\n", id, at(d, id));
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
