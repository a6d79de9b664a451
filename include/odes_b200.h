/* odes_b200.h -- umbrella header of libodes_b200.so (C ABI). */
#ifndef ODES_B200_H
#define ODES_B200_H
#include "sbmlsolver/ast.h"
#include "sbmlsolver/sbmlModel.h"
#include "sbmlsolver/solverError.h"
#include "sbmlsolver/processAST.h"
#include "sbmlsolver/integratorSettings.h"
#include "sbmlsolver/odeModel.h"
#include "sbmlsolver/cvodeData.h"
#include "sbmlsolver/integratorInstance.h"
#include "sbmlsolver/batchIntegrator.h"
#include "sbmlsolver/odeSolver.h"
#endif
