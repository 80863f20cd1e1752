#ifndef FRAC_STEP_MG_H
#define FRAC_STEP_MG_H
#include "fractionalStepGrid.hpp"
using mmgf::FractionalStepMultigrid;
#endif
