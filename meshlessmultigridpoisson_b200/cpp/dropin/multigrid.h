#ifndef MULTIGRID_H
#define MULTIGRID_H
#include "grid.h"
using mmgf::Multigrid;
#endif
