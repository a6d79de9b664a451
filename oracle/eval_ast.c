/*
 * eval_ast.c -- ORACLE (test infrastructure, not product code).
 * Independent interpreter for indexed formula trees, restating the operator
 * semantics of the reference's src/evaluateAST.c:137-601 (n-ary relational
 * chains, early-out product, piecewise, odd-degree root, factorial,
 * log(base,x), spelled-out inverse hyperbolics :17-30).  It reads values from
 * a plain value[] vector and is deliberately written separately from the
 * product's host evaluator and from its code emitter so it can check both.
 */
#include <math.h>
#include "oracle.h"

static double ev(const ASTNode_t *n, const double *v, double t);
#define K(i) ev(n->children[i], v, t)

static double chain(const ASTNode_t *n, const double *v, double t)
{
  unsigned int i; double a, b;
  if (n->nchildren <= 1) return 1.0;
  b = K(0);
  for (i = 1; i < n->nchildren; i++) {
    a = b; b = K(i);
    switch (n->type) {
    case AST_RELATIONAL_GEQ: if (a < b) return 0.0; break;
    case AST_RELATIONAL_GT: if (a <= b) return 0.0; break;
    case AST_RELATIONAL_LEQ: if (a > b) return 0.0; break;
    default: if (a >= b) return 0.0; break;            /* LT */
    }
  }
  return 1.0;
}

static double ev(const ASTNode_t *n, const double *v, double t)
{
  unsigned int i, nc = n->nchildren; double a, b, r; int cnt, j, found;
  switch (n->type) {
  case AST_INTEGER: return (double)n->ival;
  case AST_REAL: case AST_REAL_E: return n->rval;
  case AST_RATIONAL: return (double)n->ival / (double)n->denom;
  case AST_NAME: return n->index >= 0 ? v[n->index] : 0.0;
  case AST_NAME_TIME: return (double)(float)t;              /* currenttime is a float in the reference */
  case AST_CONSTANT_E: return exp(1.);
  case AST_CONSTANT_PI: return 4. * atan(1.);
  case AST_CONSTANT_TRUE: return 1.0;
  case AST_CONSTANT_FALSE: return 0.0;
  case AST_PLUS: r = 0.0; for (i = 0; i < nc; i++) r += K(i); return r;
  case AST_MINUS: if (nc < 2) return -K(0); a = K(0); b = K(1); return a - b;
  case AST_TIMES: r = 1.0; for (i = 0; i < nc; i++) { r *= K(i); if (r == 0.0) break; } return r;
  case AST_DIVIDE: a = K(0); b = K(1); return a / b;
  case AST_POWER: case AST_FUNCTION_POWER: a = K(0); b = K(1); return pow(a, b);
  case AST_FUNCTION_ABS: return fabs(K(0));
  case AST_FUNCTION_ARCCOS: return acos(K(0));
  case AST_FUNCTION_ARCCOSH: a = K(0); return log(a + (sqrt(a - 1) * sqrt(a + 1)));
  case AST_FUNCTION_ARCCOT: return atan(1. / K(0));
  case AST_FUNCTION_ARCCOTH: a = K(0); return log((a + 1) / (a - 1)) / 2;
  case AST_FUNCTION_ARCCSC: a = K(0); return atan(1 / sqrt((a - 1) * (a + 1)));
  case AST_FUNCTION_ARCCSCH: a = K(0); return log(1 / a + sqrt(1 / (a * a) + 1));
  case AST_FUNCTION_ARCSEC: return acos(1 / K(0));
  case AST_FUNCTION_ARCSECH: a = 1. / K(0); return log(a + (sqrt(a - 1) * sqrt(a + 1)));
  case AST_FUNCTION_ARCSIN: return asin(K(0));
  case AST_FUNCTION_ARCSINH: a = K(0); return log(a + sqrt((a * a) + 1));
  case AST_FUNCTION_ARCTAN: return atan(K(0));
  case AST_FUNCTION_ARCTANH: a = K(0); return (log(1 + a) - log(1 - a)) / 2;
  case AST_FUNCTION_CEILING: return ceil(K(0));
  case AST_FUNCTION_COS: return cos(K(0));
  case AST_FUNCTION_COSH: return cosh(K(0));
  case AST_FUNCTION_COT: return 1. / tan(K(0));
  case AST_FUNCTION_COTH: return 1 / tanh(K(0));
  case AST_FUNCTION_CSC: return 1. / sin(K(0));
  case AST_FUNCTION_CSCH: return 1. / sinh(K(0));
  case AST_FUNCTION_EXP: return exp(K(0));
  case AST_FUNCTION_FACTORIAL: a = K(0); j = (int)floor(a); for (r = 1; j > 1; --j) r *= j; return r;
  case AST_FUNCTION_FLOOR: return floor(K(0));
  case AST_FUNCTION_LN: return log(K(0));
  case AST_FUNCTION_LOG: a = log10(K(1)); return a / log10(K(0));
  case AST_FUNCTION_PIECEWISE:
    found = 0; r = 0.0;
    for (i = 0; i + 1 < nc; i += 2) if (K(i + 1)) { found++; r = K(i); }
    if (i + 1 == nc && !found) r = K(i);
    return r;
  case AST_FUNCTION_ROOT:
    a = K(1); b = K(0);
    if (b == floor(b) && a < 0 && ((int)b % 2 != 0)) return -pow(fabs(a), 1. / b);
    return pow(a, 1. / b);
  case AST_FUNCTION_SEC: return 1. / cos(K(0));
  case AST_FUNCTION_SECH: return 1. / cosh(K(0));
  case AST_FUNCTION_SIN: return sin(K(0));
  case AST_FUNCTION_SINH: return sinh(K(0));
  case AST_FUNCTION_TAN: return tan(K(0));
  case AST_FUNCTION_TANH: return tanh(K(0));
  case AST_LOGICAL_AND: cnt = 0; for (i = 0; i < nc; i++) cnt += (int)K(i); return cnt == (int)nc ? 1.0 : 0.0;
  case AST_LOGICAL_NOT: return (double)(!(K(0)));
  case AST_LOGICAL_OR: cnt = 0; for (i = 0; i < nc; i++) cnt += (int)K(i); return cnt > 0 ? 1.0 : 0.0;
  case AST_LOGICAL_XOR: cnt = 0; for (i = 0; i < nc; i++) cnt += (int)K(i); return (cnt % 2 != 0) ? 1.0 : 0.0;
  case AST_RELATIONAL_EQ:
    if (nc <= 1) return 1.0;
    a = K(0);
    for (i = 1; i < nc; i++) if (a != K(i)) return 0.0;
    return 1.0;
  case AST_RELATIONAL_GEQ: case AST_RELATIONAL_GT: case AST_RELATIONAL_LEQ: case AST_RELATIONAL_LT: return chain(n, v, t);
  case AST_RELATIONAL_NEQ: a = K(0); b = K(1); return (a != b);
  default: return 0.0;
  }
}
double oracle_eval(const ASTNode_t *n, const double *value, double time) { return n ? ev(n, value, time) : 0.0; }
