/* cvodeSolver.h -- header name of the reference (src/sbmlsolver/cvodeSolver.h); IntegratorInstance_cvodeOneStep and the CVODE statistics printer live with the instance */
#ifndef ODES_B200_FWD_CVODESOLVER_H
#define ODES_B200_FWD_CVODESOLVER_H
#include "sbmlsolver/integratorInstance.h"
#endif
