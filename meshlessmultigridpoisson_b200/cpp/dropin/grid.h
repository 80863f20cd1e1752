// Drop-in replacement of the reference's grid.h (see gridclasses.hpp in this directory).
#ifndef GRID_H
#define GRID_H
#include "gridclasses.hpp"
#include <tuple>
#include <string>
#include "fileReadingFunctions.h"
#include "general_computation_functions.h"
#include <stdexcept>
#include "math.h"
#include <iostream>
#include <algorithm>
#include <queue>
#include <unordered_map>
#include <time.h>
using std::vector;
using std::cout;
using std::endl;
using mmgf::Grid;
#endif
