/* sbml.h -- header name of the reference (src/sbmlsolver/sbml.h); parseModel is declared with the ODE model */
#ifndef ODES_B200_FWD_SBML_H
#define ODES_B200_FWD_SBML_H
#include "sbmlsolver/odeModel.h"
#endif
