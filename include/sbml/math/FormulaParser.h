/* libSBML's infix formula parser header (SBML_parseFormula, SBML_formulaToString): provided by csrc/ast.c */
#ifndef ODES_B200_COMPAT_FORMULAPARSER_H
#define ODES_B200_COMPAT_FORMULAPARSER_H
#include "sbmlsolver/ast.h"
#endif
