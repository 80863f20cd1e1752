#ifndef FRAC_STEP_GRID_H
#define FRAC_STEP_GRID_H
#include "grid.h"
using mmgf::FractionalStepGrid;
#endif
