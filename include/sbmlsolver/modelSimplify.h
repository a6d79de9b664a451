/* modelSimplify.h -- header name of the reference (src/sbmlsolver/modelSimplify.h); the AST_replace* helpers are declared with the formula processing */
#ifndef ODES_B200_FWD_MODELSIMPLIFY_H
#define ODES_B200_FWD_MODELSIMPLIFY_H
#include "sbmlsolver/processAST.h"
#include "sbmlsolver/odeModel.h"
#endif
