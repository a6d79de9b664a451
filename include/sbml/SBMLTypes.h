/* libSBML's umbrella header as the reference's callers include it; here the SBML front end is part of this library
   (csrc/sbml_io.c, csrc/ast.c -- SURVEY section 8f rank 1), so it forwards to the library's own model and formula types */
#ifndef ODES_B200_COMPAT_SBMLTYPES_H
#define ODES_B200_COMPAT_SBMLTYPES_H
#include "sbmlsolver/sbmlModel.h"
#endif
