/* exportdefs.h -- header name of the reference (src/sbmlsolver/exportdefs.h); no export decoration is needed for the shared library */
#ifndef ODES_B200_FWD_EXPORTDEFS_H
#define ODES_B200_FWD_EXPORTDEFS_H
#ifndef SBML_ODESOLVER_API
#define SBML_ODESOLVER_API
#endif
#endif
