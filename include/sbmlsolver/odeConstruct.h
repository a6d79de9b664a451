/* odeConstruct.h -- header name of the reference (src/sbmlsolver/odeConstruct.h); Model_reduceToOdes, Species_odeFromReactions, Model_getValueById, Model_setValue are declared with the ODE model */
#ifndef ODES_B200_FWD_ODECONSTRUCT_H
#define ODES_B200_FWD_ODECONSTRUCT_H
#include "sbmlsolver/odeModel.h"
#endif
