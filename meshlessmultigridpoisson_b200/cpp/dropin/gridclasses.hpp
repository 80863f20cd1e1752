// Drop-in replacement of the reference's gridclasses.hpp: the same class names, provided by the facade over libmmg.
// Put this directory BEFORE the reference's own directory on the include path; the reference's drivers
// (testing_functions.cpp, FractionalStepSim.cpp) then compile UNMODIFIED against the CUDA path.
#ifndef GRID_CLASSES_H
#define GRID_CLASSES_H
#include <vector>
#include <Eigen/Dense>
#include <Eigen/Sparse>
#include "../mmg_facade.hpp"
using mmgf::GridProperties;
using mmgf::Boundary;
#endif
