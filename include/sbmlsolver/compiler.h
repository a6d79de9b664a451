/* compiler.h -- header name of the reference (src/sbmlsolver/compiler.h); Compiler_compile / CompiledCode_getFunction / CompiledCode_free are declared with the ODE model */
#ifndef ODES_B200_FWD_COMPILER_H
#define ODES_B200_FWD_COMPILER_H
#include "sbmlsolver/odeModel.h"
#endif
