/* variableIndex.h -- header name of the reference (src/sbmlsolver/variableIndex.h); variableIndex_t is declared with the ODE model */
#ifndef ODES_B200_FWD_VARIABLEINDEX_H
#define ODES_B200_FWD_VARIABLEINDEX_H
#include "sbmlsolver/odeModel.h"
#endif
