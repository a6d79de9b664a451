/* sensSolver.h -- header name of the reference (src/sbmlsolver/sensSolver.h:48-62).  Forward sensitivities are part of the
   integrator instance here (integratorInstance.h); the adjoint solver, the quadratures and the objective-function entries
   of this header are out of scope and are not declared. */
#ifndef ODES_B200_FWD_SENSSOLVER_H
#define ODES_B200_FWD_SENSSOLVER_H
#include "sbmlsolver/integratorInstance.h"
#ifdef __cplusplus
extern "C" {
#endif
int IntegratorInstance_printCVODESStatistics(const integratorInstance_t *, FILE *f);
#ifdef __cplusplus
}
#endif
#endif
