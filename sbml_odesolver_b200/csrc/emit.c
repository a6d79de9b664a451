/*
 * emit.c -- code generation: formula trees -> compilable source text.
 *
 * Reference seam (kept): src/processAST.c:2057-2311 generateMacros/generateAST
 * turn an indexed AST into a C expression over value[]; src/odeModel.c:2600-2918
 * wrap those into ode_f / jacobi_f / event_f which src/compiler.c builds with
 * g++ and dlopen's.  Here the same traversal emits
 *   (1) CUDA device functions for sm_100a that are inlined into the batched
 *       BDF kernel (ODEModel_generateCUDASource), and
 *   (2) the host-C flavour of ode_f / jacobi_f (ODEModel_generateCSource) that
 *       corresponds to the reference's compiled mode (used by tests and by the
 *       CPU baseline harness only).
 * Differences to the reference emitter, on purpose: constants are printed
 * with "%.17g" (the reference's "%g" truncates to 6 digits, src/charBuffer.c:103-112),
 * piecewise supports any number of pieces, root() handles odd degrees of
 * negative radicands like evaluateAST does, and the event function follows
 * the interpreter semantics (src/integratorInstance.c:1029-1077), not the
 * re-arming variant of the reference's generated event_f.
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sbmlsolver/odeModel.h"
#include "sbmlsolver/processAST.h"
#include "sbmlsolver/solverError.h"

/* ----------------------------------------------------------- charBuffer */
struct charBuffer { char *b; size_t len, cap; };
charBuffer_t *CharBuffer_create(void) { charBuffer_t *c = (charBuffer_t *)calloc(1, sizeof *c); c->cap = 1024; c->b = (char *)malloc(c->cap); c->b[0] = 0; return c; }
void CharBuffer_free(charBuffer_t *c) { if (c) { free(c->b); free(c); } }
void CharBuffer_append(charBuffer_t *c, const char *s)
{
  size_t n = strlen(s);
  if (c->len + n + 1 > c->cap) { c->cap = 2 * (c->len + n + 1); c->b = (char *)realloc(c->b, c->cap); }
  memcpy(c->b + c->len, s, n + 1); c->len += n;
}
void CharBuffer_appendInt(charBuffer_t *c, long v) { char t[32]; snprintf(t, sizeof t, "%ld", v); CharBuffer_append(c, t); }
/* the public entry prints like the reference (src/charBuffer.c:103-112, "%g"); the emitters below use cb_double17 */
void CharBuffer_appendDouble(charBuffer_t *c, double v) { char t[64]; snprintf(t, sizeof t, "%g", v); CharBuffer_append(c, t); }
/* full precision for generated code: the value is reproduced exactly */
/* the parametric matrix df/dp exists and every entry could be differentiated */
#define SENS_OK(os) ((os) && (os)->sensitivity && (os)->sens)

static void cb_double17(charBuffer_t *c, double v)
{
  char t[64];
  if (isnan(v)) strcpy(t, "(0.0/0.0)");
  else if (isinf(v)) strcpy(t, v > 0 ? "(1.0/0.0)" : "(-1.0/0.0)");
  else {
    snprintf(t, sizeof t, "%.17g", v);
    if (!strpbrk(t, ".eEn")) strcat(t, ".0");
  }
  CharBuffer_append(c, t);
}
const char *CharBuffer_getBuffer(charBuffer_t *c) { return c->b; }
static void cbprintf(charBuffer_t *c, const char *fmt, ...)
{
  char t[2048]; va_list ap;
  va_start(ap, fmt); vsnprintf(t, sizeof t, fmt, ap); va_end(ap);
  CharBuffer_append(c, t);
}

/* --------------------------------------------------------- expressions */
/* the time symbol: currenttime is a FLOAT in the reference (src/sbmlsolver/cvodeData.h:116; set from t in f, JacODE and
   updateData, src/cvodeSolver.c:964, :1068), so a formula sees t rounded to 24 bits */
#define REF_TIME "((double)(float)t)"
typedef struct emit_ctx {
  int cuda;                       /* 1: device flavour with exact-case strength reduction */
  int interp;                     /* 1: output-time code (rules, events, stored rows): the reference evaluates these with
                                     evaluateAST even in compiled mode (src/integratorInstance.c:1039-1072, :1139-1161), so
                                     pow(x, 2) stays libm's pow; in ode_f / jacobi_f the reference's g++ -O build turns the
                                     literal-2 call into x * x */
  /* resolves value[idx] to source text */
  void (*name)(charBuffer_t *, int idx, const struct emit_ctx *);
  const char *time;               /* text of the current time */
  const void *info;
  struct leafrec *leaf;           /* non-NULL: leaf mode -- numbers and names are printed as operand placeholders x<k> and
                                     recorded, so that expressions of identical shape print identical text */
} emit_ctx_t;

/* leaf mode (warp flavour): operands of one expression.  An operand is a word of the warp's value space V (state,
   rule values, per-trajectory constants) or a number that goes to the literal pool behind it. */
#define LEAF_MAX 16
struct leafrec { int n, failed; int isnum[LEAF_MAX]; int off[LEAF_MAX]; double num[LEAF_MAX]; int dedup; /* 1: an operand that occurs again gets its first placeholder */ };
static void leaf_number(charBuffer_t *s, const emit_ctx_t *c, double v)
{
  struct leafrec *lr = c->leaf; char t[24]; int k;
  /* 1.0 (compartment sizes, stoichiometries) stays a literal in the text: x / 1.0 and 1.0 * x fold away at compile time */
  if (lr->dedup && v == 1.0) { CharBuffer_append(s, "1.0"); return; }
  for (k = 0; lr->dedup && k < lr->n; k++) if (lr->isnum[k] && memcmp(&lr->num[k], &v, sizeof v) == 0) { snprintf(t, sizeof t, "x%d", k); CharBuffer_append(s, t); return; }
  if (lr->n >= LEAF_MAX) { lr->failed = 1; CharBuffer_append(s, "0.0"); return; }
  lr->isnum[lr->n] = 1; lr->num[lr->n] = v; lr->off[lr->n] = -1;
  snprintf(t, sizeof t, "x%d", lr->n++); CharBuffer_append(s, t);
}
static void leaf_word(charBuffer_t *s, const emit_ctx_t *c, int off)
{
  struct leafrec *lr = c->leaf; char t[24]; int k;
  for (k = 0; lr->dedup && k < lr->n; k++) if (!lr->isnum[k] && lr->off[k] == off) { snprintf(t, sizeof t, "x%d", k); CharBuffer_append(s, t); return; }
  if (lr->n >= LEAF_MAX) { lr->failed = 1; CharBuffer_append(s, "0.0"); return; }
  lr->isnum[lr->n] = 0; lr->num[lr->n] = 0.0; lr->off[lr->n] = off;
  snprintf(t, sizeof t, "x%d", lr->n++); CharBuffer_append(s, t);
}

static void gen(charBuffer_t *s, const ASTNode_t *n, const emit_ctx_t *c);

static int is_const_tree(const ASTNode_t *n)
{
  unsigned int i;
  if (ASTNode_isName(n) || n->type == AST_FUNCTION || n->type == AST_FUNCTION_DELAY) return 0;
  for (i = 0; i < n->nchildren; i++) if (!is_const_tree(n->children[i])) return 0;
  return 1;
}
static void gen_nested(charBuffer_t *s, const ASTNode_t *n, const emit_ctx_t *c)
{
  switch (n->type) {
  case AST_PLUS: case AST_MINUS: case AST_TIMES: case AST_DIVIDE:
  case AST_LOGICAL_AND: case AST_LOGICAL_OR: case AST_LOGICAL_NOT:
  case AST_RELATIONAL_EQ: case AST_RELATIONAL_GEQ: case AST_RELATIONAL_GT:
  case AST_RELATIONAL_LEQ: case AST_RELATIONAL_LT: case AST_RELATIONAL_NEQ:
    CharBuffer_append(s, "("); gen(s, n, c); CharBuffer_append(s, ")"); break;
  default: gen(s, n, c);
  }
}
static void gen_nary(charBuffer_t *s, const ASTNode_t *n, const char *op, const emit_ctx_t *c)
{
  unsigned int i;
  for (i = 0; i < n->nchildren; i++) {
    gen_nested(s, n->children[i], c);
    if (i + 1 != n->nchildren) { CharBuffer_append(s, " "); CharBuffer_append(s, op); CharBuffer_append(s, " "); }
  }
}
static void gen_call(charBuffer_t *s, const ASTNode_t *n, const char *fn, const emit_ctx_t *c)
{
  unsigned int i;
  CharBuffer_append(s, fn); CharBuffer_append(s, "(");
  for (i = 0; i < n->nchildren; i++) { gen(s, n->children[i], c); if (i + 1 != n->nchildren) CharBuffer_append(s, ", "); }
  CharBuffer_append(s, ")");
}
/* n-ary relational chain a < b < c  ->  ((a < b) && (b < c)) as doubles; children are pure */
static void gen_chain(charBuffer_t *s, const ASTNode_t *n, const char *op, const emit_ctx_t *c)
{
  unsigned int i;
  if (n->nchildren <= 1) { CharBuffer_append(s, "1.0"); return; }
  CharBuffer_append(s, "(double)(");
  for (i = 0; i + 1 < n->nchildren; i++) {
    if (i) CharBuffer_append(s, " && ");
    CharBuffer_append(s, "("); gen_nested(s, n->children[i], c); cbprintf(s, " %s ", op); gen_nested(s, n->children[i + 1], c); CharBuffer_append(s, ")");
  }
  CharBuffer_append(s, ")");
}
static void gen_piecewise(charBuffer_t *s, const ASTNode_t *n, const emit_ctx_t *c)
{
  /* evaluateAST lets the LAST true piece win, falls back to the odd trailing child, else 0:
     nest so that later pieces are tested first */
  unsigned int np = n->nchildren / 2, i;
  int opened = 0;
  for (i = np; i-- > 0;) {
    CharBuffer_append(s, "(("); gen(s, n->children[2 * i + 1], c); CharBuffer_append(s, ") != 0.0 ? ("); gen(s, n->children[2 * i], c); CharBuffer_append(s, ") : ");
    opened++;
  }
  if (n->nchildren % 2) { CharBuffer_append(s, "("); gen(s, n->children[n->nchildren - 1], c); CharBuffer_append(s, ")"); }
  else CharBuffer_append(s, "0.0");
  while (opened--) CharBuffer_append(s, ")");
}

static void gen(charBuffer_t *s, const ASTNode_t *n, const emit_ctx_t *c)
{
  unsigned int i;
  /* constant sub-expressions of numbers are folded at emit time with the host evaluator
     (same IEEE operations the interpreter would perform at run time) */
  if (c->cuda && n->nchildren > 0 && is_const_tree(n) && !ASTNode_isLogical(n) && !ASTNode_isRelational(n) && n->type != AST_FUNCTION_PIECEWISE) {
    if (c->leaf) leaf_number(s, c, evaluateAST((ASTNode_t *)n, NULL));
    else cb_double17(s, evaluateAST((ASTNode_t *)n, NULL));
    return;
  }
  switch (n->type) {
  case AST_PLUS: gen_nary(s, n, "+", c); break;
  case AST_TIMES: gen_nary(s, n, "*", c); break;
  case AST_MINUS:
    if (n->nchildren == 1) { CharBuffer_append(s, "-"); gen_nested(s, n->children[0], c); }
    else gen_nary(s, n, "-", c);
    break;
  case AST_DIVIDE: gen_nary(s, n, "/", c); break;
  case AST_POWER: case AST_FUNCTION_POWER:
    if (c->cuda && n->nchildren == 2 && is_const_tree(n->children[1])) {
      double e = evaluateAST(n->children[1], NULL);
      if (e == 0.0) { CharBuffer_append(s, "1.0"); break; }               /* pow(x,0) == 1 exactly  */
      if (e == 1.0) { gen_nested(s, n->children[0], c); break; }          /* pow(x,1) == x exactly  */
      if (e == 2.0 && !c->interp) { CharBuffer_append(s, "odes_sqr("); gen(s, n->children[0], c); CharBuffer_append(s, ")"); break; }
    }
    gen_call(s, n, c->cuda ? "odes_pow" : "pow", c); break;
  case AST_INTEGER: if (c->leaf) leaf_number(s, c, (double)n->ival); else cb_double17(s, (double)n->ival); break;
  case AST_REAL: case AST_REAL_E: case AST_RATIONAL: if (c->leaf) leaf_number(s, c, ASTNode_getReal(n)); else cb_double17(s, ASTNode_getReal(n)); break;
  case AST_NAME:
    if (ASTNode_isSetIndex(n)) c->name(s, (int)ASTNode_getIndex(n), c);
    else {
      SolverError_error(FATAL_ERROR_TYPE, SOLVER_ERROR_AST_COMPILATION_FAILED_MISSING_VALUE,
                        "generateAST: No value found for AST_NAME %s. Defaults to Zero to avoid program crash", ASTNode_getName(n));
      CharBuffer_append(s, "0.0");
    }
    break;
  case AST_NAME_TIME: CharBuffer_append(s, c->time); break;
  case AST_CONSTANT_E: cb_double17(s, exp(1.)); break;
  case AST_CONSTANT_FALSE: CharBuffer_append(s, "0.0"); break;
  case AST_CONSTANT_PI: cb_double17(s, 4. * atan(1.)); break;
  case AST_CONSTANT_TRUE: CharBuffer_append(s, "1.0"); break;
  case AST_FUNCTION_ABS: gen_call(s, n, "fabs", c); break;
  case AST_FUNCTION_ARCCOS: gen_call(s, n, "acos", c); break;
  case AST_FUNCTION_ARCCOSH: gen_call(s, n, "odes_acosh", c); break;
  case AST_FUNCTION_ARCCOT: gen_call(s, n, "odes_acot", c); break;
  case AST_FUNCTION_ARCCOTH: gen_call(s, n, "odes_acoth", c); break;
  case AST_FUNCTION_ARCCSC: gen_call(s, n, "odes_acsc", c); break;
  case AST_FUNCTION_ARCCSCH: gen_call(s, n, "odes_acsch", c); break;
  case AST_FUNCTION_ARCSEC: gen_call(s, n, "odes_asec", c); break;
  case AST_FUNCTION_ARCSECH: gen_call(s, n, "odes_asech", c); break;
  case AST_FUNCTION_ARCSIN: gen_call(s, n, "asin", c); break;
  case AST_FUNCTION_ARCSINH: gen_call(s, n, "odes_asinh", c); break;
  case AST_FUNCTION_ARCTAN: gen_call(s, n, "atan", c); break;
  case AST_FUNCTION_ARCTANH: gen_call(s, n, "odes_atanh", c); break;
  case AST_FUNCTION_CEILING: gen_call(s, n, "ceil", c); break;
  case AST_FUNCTION_COS: gen_call(s, n, "cos", c); break;
  case AST_FUNCTION_COSH: gen_call(s, n, "cosh", c); break;
  case AST_FUNCTION_COT: gen_call(s, n, "odes_cot", c); break;
  case AST_FUNCTION_COTH: gen_call(s, n, "odes_coth", c); break;
  case AST_FUNCTION_CSC: gen_call(s, n, "odes_csc", c); break;
  case AST_FUNCTION_CSCH: gen_call(s, n, "odes_csch", c); break;
  case AST_FUNCTION_EXP: gen_call(s, n, c->cuda ? "odes_exp" : "exp", c); break;      /* the host libm's exp, restated (glibc_pow.inc) */
  case AST_FUNCTION_FACTORIAL: gen_call(s, n, "odes_factorial", c); break;
  case AST_FUNCTION_FLOOR: gen_call(s, n, "floor", c); break;
  case AST_FUNCTION_LN: gen_call(s, n, c->cuda ? "odes_log" : "log", c); break;
  case AST_FUNCTION_LOG: gen_call(s, n, "odes_logbase", c); break;
  case AST_FUNCTION_PIECEWISE: gen_piecewise(s, n, c); break;
  case AST_FUNCTION_ROOT: gen_call(s, n, "odes_root", c); break;
  case AST_FUNCTION_SEC: gen_call(s, n, "odes_sec", c); break;
  case AST_FUNCTION_SECH: gen_call(s, n, "odes_sech", c); break;
  case AST_FUNCTION_SIN: gen_call(s, n, "sin", c); break;
  case AST_FUNCTION_SINH: gen_call(s, n, "sinh", c); break;
  case AST_FUNCTION_TAN: gen_call(s, n, "tan", c); break;
  case AST_FUNCTION_TANH: gen_call(s, n, "tanh", c); break;
  case AST_LOGICAL_AND: case AST_LOGICAL_OR: case AST_LOGICAL_XOR: {
      /* evaluateAST sums the (int-truncated) truth values: and = all, or = any, xor = odd count */
      CharBuffer_append(s, "(double)((");
      for (i = 0; i < n->nchildren; i++) { if (i) CharBuffer_append(s, " + "); CharBuffer_append(s, "(int)("); gen(s, n->children[i], c); CharBuffer_append(s, ")"); }
      if (n->nchildren == 0) CharBuffer_append(s, "0");
      if (n->type == AST_LOGICAL_AND) cbprintf(s, ") == %u)", n->nchildren);
      else if (n->type == AST_LOGICAL_OR) CharBuffer_append(s, ") > 0)");
      else CharBuffer_append(s, ") % 2 != 0)");
      break;
    }
  case AST_LOGICAL_NOT: CharBuffer_append(s, "(double)(!("); gen(s, n->children[0], c); CharBuffer_append(s, "))"); break;
  case AST_RELATIONAL_EQ:
    if (n->nchildren <= 1) { CharBuffer_append(s, "1.0"); break; }
    CharBuffer_append(s, "(double)(");
    for (i = 1; i < n->nchildren; i++) { if (i > 1) CharBuffer_append(s, " && "); CharBuffer_append(s, "("); gen_nested(s, n->children[0], c); CharBuffer_append(s, " == "); gen_nested(s, n->children[i], c); CharBuffer_append(s, ")"); }
    CharBuffer_append(s, ")");
    break;
  case AST_RELATIONAL_GEQ: gen_chain(s, n, ">=", c); break;
  case AST_RELATIONAL_GT: gen_chain(s, n, ">", c); break;
  case AST_RELATIONAL_LEQ: gen_chain(s, n, "<=", c); break;
  case AST_RELATIONAL_LT: gen_chain(s, n, "<", c); break;
  case AST_RELATIONAL_NEQ: CharBuffer_append(s, "(double)("); gen_nested(s, n->children[0], c); CharBuffer_append(s, " != "); gen_nested(s, n->children[1], c); CharBuffer_append(s, ")"); break;
  default:
    SolverError_error(FATAL_ERROR_TYPE, SOLVER_ERROR_AST_COMPILATION_FAILED_STRANGE_NODE_TYPE,
                      "Found strange node type %d whilst generating code.", (int)n->type);
    CharBuffer_append(s, "0.0");
  }
}

/* value[] naming of the C flavour */
static void name_value(charBuffer_t *s, int idx, const emit_ctx_t *c) { (void)c; cbprintf(s, "value[%d]", idx); }

/* The reference's public entry (src/processAST.c:2117-2311): an indexed AST as the C expression text the reference
   writes into its generated sources -- numbers as ((realtype)%g), indexed names as value[i], time as
   data->currenttime, the macro names of generateMacros (acot, sec ...), operator children in parentheses.  This is the
   text API callers and the reference's unit tests see (unittest/test_processAST.c:190-247); the code THIS library
   compiles comes from gen() above (full-precision constants, exact-case strength reduction). */
static void genref(charBuffer_t *s, const ASTNode_t *n);
static int ref_is_operator(const ASTNode_t *n)
{
  switch (n->type) {
  case AST_PLUS: case AST_MINUS: case AST_TIMES: case AST_DIVIDE: case AST_POWER:
  case AST_LOGICAL_AND: case AST_LOGICAL_OR: case AST_LOGICAL_XOR: case AST_LOGICAL_NOT:
  case AST_RELATIONAL_EQ: case AST_RELATIONAL_GEQ: case AST_RELATIONAL_GT: case AST_RELATIONAL_LEQ: case AST_RELATIONAL_LT: case AST_RELATIONAL_NEQ:
    return 1;
  default: return 0;
  }
}
static void genref_nested(charBuffer_t *s, const ASTNode_t *n)
{
  if (ref_is_operator(n)) { CharBuffer_append(s, "("); genref(s, n); CharBuffer_append(s, ")"); }
  else genref(s, n);
}
static void genref_join(charBuffer_t *s, const ASTNode_t *n, const char *op)
{
  unsigned int i;
  for (i = 0; i < n->nchildren; i++) { if (i) CharBuffer_append(s, op); genref_nested(s, n->children[i]); }
}
static void genref_call(charBuffer_t *s, const ASTNode_t *n, const char *fn)
{
  unsigned int i;
  CharBuffer_append(s, fn); CharBuffer_append(s, "(");
  for (i = 0; i < n->nchildren; i++) { if (i) CharBuffer_append(s, ", "); genref(s, n->children[i]); }
  CharBuffer_append(s, ")");
}
static void genref_number(charBuffer_t *s, double v) { CharBuffer_append(s, "((realtype)"); CharBuffer_appendDouble(s, v); CharBuffer_append(s, ")"); }
static void genref(charBuffer_t *s, const ASTNode_t *n)
{
  unsigned int i;
  switch (n->type) {
  case AST_PLUS: genref_join(s, n, " + "); break;
  case AST_TIMES: genref_join(s, n, " * "); break;
  case AST_DIVIDE: genref_join(s, n, " / "); break;
  case AST_MINUS:
    if (n->nchildren == 1) { CharBuffer_append(s, "-"); genref_nested(s, n->children[0]); }
    else genref_join(s, n, " - ");
    break;
  case AST_POWER: case AST_FUNCTION_POWER: genref_call(s, n, "pow"); break;
  case AST_INTEGER: genref_number(s, (double)n->ival); break;
  case AST_REAL: case AST_REAL_E: case AST_RATIONAL: genref_number(s, ASTNode_getReal(n)); break;
  case AST_NAME:
    if (ASTNode_isSetIndex(n)) cbprintf(s, "value[%d]", (int)ASTNode_getIndex(n));
    else {
      SolverError_error(FATAL_ERROR_TYPE, SOLVER_ERROR_AST_COMPILATION_FAILED_MISSING_VALUE,
                        "generateAST: No value found for AST_NAME %s. Defaults to Zero to avoid program crash", ASTNode_getName(n));
      CharBuffer_append(s, "0.0");
    }
    break;
  case AST_NAME_TIME: CharBuffer_append(s, "data->currenttime"); break;
  case AST_CONSTANT_E: CharBuffer_appendDouble(s, exp(1.)); break;
  case AST_CONSTANT_FALSE: CharBuffer_append(s, "0"); break;
  case AST_CONSTANT_PI: CharBuffer_appendDouble(s, 4. * atan(1.)); break;
  case AST_CONSTANT_TRUE: CharBuffer_append(s, "1"); break;
  case AST_FUNCTION_ABS: genref_call(s, n, "fabs"); break;
  case AST_FUNCTION_ARCCOS: genref_call(s, n, "acos"); break;
  case AST_FUNCTION_ARCCOSH: genref_call(s, n, "acosh"); break;
  case AST_FUNCTION_ARCCOT: genref_call(s, n, "acot"); break;
  case AST_FUNCTION_ARCCOTH: genref_call(s, n, "acoth"); break;
  case AST_FUNCTION_ARCCSC: genref_call(s, n, "acsc"); break;
  case AST_FUNCTION_ARCCSCH: genref_call(s, n, "acsch"); break;
  case AST_FUNCTION_ARCSEC: genref_call(s, n, "asec"); break;
  case AST_FUNCTION_ARCSECH: genref_call(s, n, "asech"); break;
  case AST_FUNCTION_ARCSIN: genref_call(s, n, "asin"); break;
  case AST_FUNCTION_ARCSINH: genref_call(s, n, "asinh"); break;
  case AST_FUNCTION_ARCTAN: genref_call(s, n, "atan"); break;
  case AST_FUNCTION_ARCTANH: genref_call(s, n, "atanh"); break;
  case AST_FUNCTION_CEILING: genref_call(s, n, "ceil"); break;
  case AST_FUNCTION_COS: genref_call(s, n, "cos"); break;
  case AST_FUNCTION_COSH: genref_call(s, n, "cosh"); break;
  case AST_FUNCTION_COT: genref_call(s, n, "cot"); break;
  case AST_FUNCTION_COTH: genref_call(s, n, "coth"); break;
  case AST_FUNCTION_CSC: genref_call(s, n, "csc"); break;
  case AST_FUNCTION_CSCH: genref_call(s, n, "csch"); break;
  case AST_FUNCTION_EXP: genref_call(s, n, "exp"); break;
  case AST_FUNCTION_FACTORIAL: genref_call(s, n, "factorial"); break;
  case AST_FUNCTION_FLOOR: genref_call(s, n, "floor"); break;
  case AST_FUNCTION_LN: case AST_FUNCTION_LOG: genref_call(s, n, "log"); break;
  case AST_FUNCTION_ROOT:
    if (n->nchildren == 2 && n->denom == -2) { CharBuffer_append(s, "root("); genref(s, n->children[1]); CharBuffer_append(s, ")"); }   /* written as root(x) */
    else genref_call(s, n, "root");
    break;
  case AST_FUNCTION_SEC: genref_call(s, n, "sec"); break;
  case AST_FUNCTION_SECH: genref_call(s, n, "sech"); break;
  case AST_FUNCTION_SIN: genref_call(s, n, "sin"); break;
  case AST_FUNCTION_SINH: genref_call(s, n, "sinh"); break;
  case AST_FUNCTION_TAN: genref_call(s, n, "tan"); break;
  case AST_FUNCTION_TANH: genref_call(s, n, "tanh"); break;
  case AST_LOGICAL_AND: genref_join(s, n, " && "); break;
  case AST_LOGICAL_OR: genref_join(s, n, " || "); break;
  case AST_LOGICAL_NOT: CharBuffer_append(s, "!"); genref_nested(s, n->children[0]); break;
  case AST_LOGICAL_XOR:
    CharBuffer_append(s, "((");
    for (i = 0; i < n->nchildren; i++) { if (i) CharBuffer_append(s, " + "); CharBuffer_append(s, "("); genref_nested(s, n->children[i]); CharBuffer_append(s, " ? 1 : 0)"); }
    CharBuffer_append(s, ") % 2) != 0");
    break;
  case AST_RELATIONAL_EQ: genref_join(s, n, " == "); break;
  case AST_RELATIONAL_GEQ: genref_join(s, n, " >= "); break;
  case AST_RELATIONAL_GT: genref_join(s, n, " > "); break;
  case AST_RELATIONAL_LEQ: genref_join(s, n, " <= "); break;
  case AST_RELATIONAL_LT: genref_join(s, n, " < "); break;
  case AST_RELATIONAL_NEQ: genref_join(s, n, " != "); break;
  default: {                                   /* piecewise, delay ...: the text this library compiles */
      emit_ctx_t c; c.cuda = 0; c.interp = 0; c.name = name_value; c.time = "data->currenttime"; c.info = NULL; c.leaf = NULL;
      gen(s, n, &c);
    }
  }
}
void generateAST(charBuffer_t *s, const ASTNode_t *node) { genref(s, node); }

/* helper library shared by both flavours; DEV expands to __device__ __forceinline__ or static inline */
void generateMacros(charBuffer_t *b)
{
  CharBuffer_append(b,
    "DEV double odes_sqr(double x) { return x * x; }\n"
    "DEV double odes_acosh(double x) { return log(x + (sqrt(x - 1) * sqrt(x + 1))); }\n"
    "DEV double odes_asinh(double x) { return log(x + sqrt((x * x) + 1)); }\n"
    "DEV double odes_atanh(double x) { return (log(1 + x) - log(1 - x)) / 2; }\n"
    "DEV double odes_acot(double x) { return atan(1. / x); }\n"
    "DEV double odes_acoth(double x) { return log((x + 1) / (x - 1)) / 2; }\n"
    "DEV double odes_acsc(double x) { return atan(1 / sqrt((x - 1) * (x + 1))); }\n"
    "DEV double odes_acsch(double x) { return log(1 / x + sqrt(1 / (x * x) + 1)); }\n"
    "DEV double odes_asec(double x) { return acos(1 / x); }\n"
    "DEV double odes_asech(double x) { return odes_acosh(1. / x); }\n"
    "DEV double odes_cot(double x) { return 1. / tan(x); }\n"
    "DEV double odes_coth(double x) { return 1 / tanh(x); }\n"
    "DEV double odes_csc(double x) { return 1. / sin(x); }\n"
    "DEV double odes_csch(double x) { return 1. / sinh(x); }\n"
    "DEV double odes_sec(double x) { return 1. / cos(x); }\n"
    "DEV double odes_sech(double x) { return 1. / cosh(x); }\n"
    "DEV double odes_logbase(double base, double x) { return log10(x) / log10(base); }\n"
    "DEV double odes_root(double degree, double x)\n"
    "{ if (degree == floor(degree) && x < 0 && ((int)degree % 2 != 0)) return -ODES_POW(fabs(x), 1. / degree); return ODES_POW(x, 1. / degree); }\n"
    "DEV double odes_factorial(double x) { double r; int j = (int)floor(x); for (r = 1; j > 1; --j) r *= j; return r; }\n");
}

/* ------------------------------------------------------- host C flavour */
char *ODEModel_generateCSource(odeModel_t *om) { return ODEModel_generateCSourceSens(om, NULL); }
char *ODEModel_generateCSourceSens(odeModel_t *om, const odeSense_t *os)
{
  charBuffer_t *b = CharBuffer_create();
  emit_ctx_t c; int i; char *out;
  c.cuda = 0; c.interp = 0; c.name = name_value; c.time = REF_TIME; c.info = NULL; c.leaf = NULL;
  if (!om->jacob && !om->jacobianFailed) ODEModel_constructJacobian(om);
  CharBuffer_append(b, "/* generated by ODEModel_generateCSource: the reference's compiled mode (ode_f, jacobi_f) */\n#include <math.h>\n#define DEV static inline\n#define ODES_POW pow\n");
  generateMacros(b);
  CharBuffer_append(b, "void ode_f(double t, const double *y, double *ydot, double *value)\n{\n  (void)t;\n");
  for (i = 0; i < om->neq; i++) cbprintf(b, "  value[%d] = y[%d];\n", i, i);
  for (i = 0; i < om->nassbeforeodes; i++) {
    nonzeroElem_t *o = om->assignmentsBeforeODEs[i];
    cbprintf(b, "  value[%d] = ", o->i); gen(b, o->ij, &c); CharBuffer_append(b, ";\n");
  }
  for (i = 0; i < om->neq; i++) { cbprintf(b, "  ydot[%d] = ", i); gen(b, om->ode[i], &c); CharBuffer_append(b, ";\n"); }
  CharBuffer_append(b, "}\nvoid jacobi_f(double t, const double *y, double *J, double *value)\n{\n  (void)t;\n");
  for (i = 0; i < om->neq; i++) cbprintf(b, "  value[%d] = y[%d];\n", i, i);
  for (i = 0; i < om->sparsesize; i++) {
    nonzeroElem_t *nz = om->jacobSparse[i];
    cbprintf(b, "  J[%d] = ", nz->j * om->neq + nz->i); gen(b, nz->ij, &c); CharBuffer_append(b, ";\n");
  }
  CharBuffer_append(b, "}\n");
  /* sense_f: src/odeModel.c:2994-3100 -- per row: 0, += J_ij * yS[j] for the non-zero Jacobian entries, += df_i/dp for
     the parameter of sensitivity iS */
  if (os && om->jacobian) {
    int j, k;
    CharBuffer_append(b, "void sense_f(int iS, double t, const double *y, const double *yS, double *ySdot, double *value)\n{\n  (void)t; (void)iS;\n");
    for (i = 0; i < om->neq; i++) cbprintf(b, "  value[%d] = y[%d];\n", i, i);
    for (i = 0; i < om->neq; i++) {
      cbprintf(b, "  ySdot[%d] = 0.0;\n", i);
      for (j = 0; j < om->sparsesize; j++) {
        nonzeroElem_t *nz = om->jacobSparse[j];
        if (nz->i != i) continue;
        cbprintf(b, "  ySdot[%d] += ( ", i); gen(b, nz->ij, &c); cbprintf(b, ") * yS[%d];\n", nz->j);
      }
      for (k = 0; k < os->nsens; k++) {
        if (os->index_sensP[k] == -1 || !SENS_OK(os) || !os->sensLogic[i][os->index_sensP[k]]) continue;
        cbprintf(b, "  if ( %d == iS ) ySdot[%d] += ", k, i); gen(b, os->sens[i][os->index_sensP[k]], &c); CharBuffer_append(b, ";\n");
      }
    }
    CharBuffer_append(b, "}\n");
  }
  out = b->b; b->b = NULL; CharBuffer_free(b);
  return out;
}

/* -------------------------------------------------------- CUDA flavour */
/* storage classes of value[] entries inside the kernel */
enum { K_ODE, K_ASSIGNED, K_THREAD, K_UNIFORM };
typedef struct {
  const odeModel_t *om;
  int *kind;          /* [nvalues] storage class                                   */
  int *slot;          /* [nvalues] K_THREAD: index into pv[]; K_ASSIGNED static: pv slot or -1 */
  int *local;         /* [nvalues] 1 if an assigned value is a fresh local a[] in this function */
  const double *base; /* [nvalues] values after reset, or NULL: uniform constants are then read from c_base[] */
} kinfo_t;

static void name_cuda(charBuffer_t *s, int idx, const emit_ctx_t *c)
{
  const kinfo_t *k = (const kinfo_t *)c->info;
  const odeModel_t *om = k->om;
  switch (k->kind[idx]) {
  case K_ODE: cbprintf(s, "y[%d]", idx); break;
  case K_ASSIGNED:
    if (k->local[idx]) cbprintf(s, "a[%d]", idx - om->neq);
    else cbprintf(s, "pv[%d]", k->slot[idx]);
    break;
  case K_THREAD: cbprintf(s, "pv[%d]", k->slot[idx]); break;
  default:
    /* a constant that no trajectory changes: its value is known now, so x / 1.0, x * 1.0 ... fold at compile time
       (exactly: these identities hold in IEEE arithmetic) */
    if (k->base) { CharBuffer_append(s, "("); cb_double17(s, k->base[idx]); CharBuffer_append(s, ")"); }
    else cbprintf(s, "c_base[%d]", idx);
  }
}

static void mark_refs(const ASTNode_t *n, int *used)
{
  unsigned int i;
  if (ASTNode_isSetIndex(n)) used[ASTNode_getIndex(n)] = 1;
  for (i = 0; i < n->nchildren; i++) mark_refs(n->children[i], used);
}

static void emit_refresh_all(charBuffer_t *b, const odeModel_t *om, const emit_ctx_t *c, const char *indent)
{
  int i;
  for (i = 0; i < om->nass; i++) {
    nonzeroElem_t *o = om->assignmentOrder[i];
    cbprintf(b, "%sa[%d] = ", indent, o->i - om->neq); gen(b, o->ij, c); CharBuffer_append(b, ";\n");
  }
}


/* ------------------------------------------------ warp-per-trajectory flavour (bdf_warp_kernel.cuh)
   The same expressions, one element per lane.  The warp keeps a value space V in shared memory:
     V[0..32) y, V[W_AS_OFF + r] rule value r, V[W_PV_OFF + k] per-trajectory constant k, V[W_LIT_OFF + n] literal n.
   Expressions of identical shape (identical text once numbers and names are replaced by operand placeholders) are
   evaluated by all their lanes together, each lane reading ITS operands through a packed offset word it keeps in a
   register (w_group_task); right-hand sides that are signed sums of rule values -- the form odeConstruct builds from
   the stoichiometry -- are evaluated from a packed term list (w_ode_task) in the emitted left-to-right order; whatever
   fits neither is evaluated per lane in a switch (w_rules_level, w_ode_case).  Every lane performs exactly the IEEE
   operations of its element's emitted expression, in the emitted order. */
#define W_MAXGROUPS 8
#define W_GROUP_LEAVES 5
#define W_ODE_TERMS_MAX 10
#define JE_MAX_ROUNDS 8
typedef struct { int npv, nass, neq, nlit; double *lit; int naspad, npvpad; const kinfo_t *k; } wspace_t;
static int w_as_off(const wspace_t *w) { (void)w; return 32; }
static int w_pv_off(const wspace_t *w) { return 32 + w->naspad; }
static int w_lit_off(const wspace_t *w) { return 32 + w->naspad + w->npvpad; }
static int w_literal(wspace_t *w, double v)
{
  int i;
  for (i = 0; i < w->nlit; i++) if (memcmp(&w->lit[i], &v, sizeof v) == 0) return i;
  w->lit = (double *)realloc(w->lit, sizeof(double) * (size_t)(w->nlit + 1));
  w->lit[w->nlit] = v;
  return w->nlit++;
}
static const wspace_t *g_wspace;      /* the resolver below has no other way to reach it */
static void name_leaf(charBuffer_t *s, int idx, const emit_ctx_t *c)
{
  const kinfo_t *k = (const kinfo_t *)c->info; const odeModel_t *om = k->om;
  switch (k->kind[idx]) {
  case K_ODE: leaf_word(s, c, idx); break;
  case K_ASSIGNED:
    if (k->local[idx]) leaf_word(s, c, w_as_off(g_wspace) + idx - om->neq);
    else leaf_word(s, c, w_pv_off(g_wspace) + k->slot[idx]);
    break;
  case K_THREAD: leaf_word(s, c, w_pv_off(g_wspace) + k->slot[idx]); break;
  default:
    if (k->base) leaf_number(s, c, k->base[idx]);
    else c->leaf->failed = 1, CharBuffer_append(s, "0.0");
  }
}
/* a name that is one word of V, or -1 */
static int w_plain_word(const ASTNode_t *n, const wspace_t *w)
{
  const kinfo_t *k = w->k; int idx;
  if (n->type != AST_NAME || !ASTNode_isSetIndex(n)) return -1;
  idx = (int)ASTNode_getIndex(n);
  switch (k->kind[idx]) {
  case K_ODE: return idx;
  case K_ASSIGNED: return k->local[idx] ? w_as_off(w) + idx - w->neq : w_pv_off(w) + k->slot[idx];
  case K_THREAD: return w_pv_off(w) + k->slot[idx];
  default: return -1;
  }
}
/* left-to-right signed sum of plain words: off[], neg[]; returns the number of terms or -1 */
static int w_flatten_sum(const ASTNode_t *n, int neg, const wspace_t *w, int *off, int *negs, int cap)
{
  unsigned int i; int cnt, o;
  if ((o = w_plain_word(n, w)) >= 0) { if (cap < 1) return -1; off[0] = o; negs[0] = neg; return 1; }
  if (n->type == AST_MINUS && n->nchildren == 1) {
    if ((o = w_plain_word(n->children[0], w)) < 0 || cap < 1) return -1;
    off[0] = o; negs[0] = !neg; return 1;
  }
  if ((n->type != AST_PLUS && n->type != AST_MINUS) || n->nchildren < 2 || neg) return -1;
  cnt = w_flatten_sum(n->children[0], 0, w, off, negs, cap);
  if (cnt < 0) return -1;
  for (i = 1; i < n->nchildren; i++) {
    if ((o = w_plain_word(n->children[i], w)) < 0 || cnt >= cap) return -1;
    off[cnt] = o; negs[cnt] = (n->type == AST_MINUS); cnt++;
  }
  return cnt;
}
/* ---- the entries as ONE expression DAG, evaluated level by level (w_dag_pass in bdf_warp_kernel.cuh) ----
   Every entry's tree is taken apart into binary IEEE operations exactly as gen() writes them (n-ary + - * / fold from the
   left, a literal-2 power is x * x, constant sub-trees are folded by the host evaluator, multiplications / divisions by the
   literal 1.0 drop out -- x * 1.0 == x / 1.0 == x for every x); identical sub-expressions become one node (same operands,
   same operation: same bits).  A node's level is one above its operands'; the nodes of a level with the same operation
   form a step of at most 32 tasks, one per lane.  The kernel then walks a table of steps with one small loop: the whole
   Jacobian and parametric matrix cost (levels x operation kinds) short passes instead of one code body per shape. */
enum { D_LEAF = 0, D_ADD, D_SUB, D_MUL, D_DIV, D_NEG, D_POW, D_EXP, D_LOG, D_NOPS };
typedef struct { int op, a, b, level, off, isone; } dagnode_t;
typedef struct { dagnode_t *n; int cnt, cap, failed; wspace_t *w; const kinfo_t *k; const emit_ctx_t *c; } dag_t;
static int dag_new(dag_t *d, int op, int a, int b, int level, int off, int isone)
{
  if (d->cnt == d->cap) { d->cap = d->cap ? 2 * d->cap : 256; d->n = (dagnode_t *)realloc(d->n, sizeof(dagnode_t) * (size_t)d->cap); }
  d->n[d->cnt].op = op; d->n[d->cnt].a = a; d->n[d->cnt].b = b; d->n[d->cnt].level = level; d->n[d->cnt].off = off; d->n[d->cnt].isone = isone;
  return d->cnt++;
}
static int dag_leaf(dag_t *d, int off, int isone)
{
  int i;
  for (i = 0; i < d->cnt; i++) if (d->n[i].op == D_LEAF && d->n[i].off == off) return i;
  return dag_new(d, D_LEAF, -1, -1, 0, off, isone);
}
static int dag_num(dag_t *d, double v) { return dag_leaf(d, w_lit_off(d->w) + w_literal(d->w, v), v == 1.0 && !signbit(v)); }
static int dag_op(dag_t *d, int op, int a, int b)
{
  int i, lv;
  if (a < 0 || (b < 0 && op != D_NEG && op != D_EXP && op != D_LOG)) { d->failed = 1; return -1; }
  if (op == D_MUL && d->n[a].isone) return b;
  if ((op == D_MUL || op == D_DIV) && d->n[b].isone) return a;
  if (b < 0) b = a;
  for (i = 0; i < d->cnt; i++) if (d->n[i].op == op && d->n[i].a == a && d->n[i].b == b) return i;
  lv = d->n[a].level > d->n[b].level ? d->n[a].level : d->n[b].level;
  return dag_new(d, op, a, b, lv + 1, -1, 0);
}
static int dag_new_or_find_sqr(dag_t *d, int a)
{
  int i;
  for (i = 0; i < d->cnt; i++) if (d->n[i].op == D_MUL && d->n[i].a == a && d->n[i].b == a) return i;
  return dag_new(d, D_MUL, a, a, d->n[a].level + 1, -1, 0);
}
static int dag_build(dag_t *d, const ASTNode_t *n)
{
  unsigned int i; int r, op;
  if (d->failed) return -1;
  if (n->nchildren > 0 && is_const_tree(n) && !ASTNode_isLogical(n) && !ASTNode_isRelational(n) && n->type != AST_FUNCTION_PIECEWISE)
    return dag_num(d, evaluateAST((ASTNode_t *)n, NULL));
  switch (n->type) {
  case AST_PLUS: case AST_TIMES: case AST_DIVIDE: case AST_MINUS:
    if (n->nchildren < 1) { d->failed = 1; return -1; }
    if (n->type == AST_MINUS && n->nchildren == 1) return dag_op(d, D_NEG, dag_build(d, n->children[0]), -1);
    op = n->type == AST_PLUS ? D_ADD : n->type == AST_TIMES ? D_MUL : n->type == AST_DIVIDE ? D_DIV : D_SUB;
    r = dag_build(d, n->children[0]);
    for (i = 1; i < n->nchildren; i++) r = dag_op(d, op, r, dag_build(d, n->children[i]));
    return r;
  case AST_POWER: case AST_FUNCTION_POWER:
    if (n->nchildren != 2) { d->failed = 1; return -1; }
    if (is_const_tree(n->children[1])) {
      const double e = evaluateAST(n->children[1], NULL);
      if (e == 0.0) return dag_num(d, 1.0);
      if (e == 1.0) return dag_build(d, n->children[0]);
      if (e == 2.0 && !d->c->interp) { r = dag_build(d, n->children[0]); return d->failed ? -1 : dag_new_or_find_sqr(d, r); }
    }
    r = dag_build(d, n->children[0]);
    return dag_op(d, D_POW, r, dag_build(d, n->children[1]));
  case AST_INTEGER: return dag_num(d, (double)n->ival);
  case AST_REAL: case AST_REAL_E: case AST_RATIONAL: return dag_num(d, ASTNode_getReal(n));
  case AST_NAME: {
      const kinfo_t *k = d->k; int idx;
      if (!ASTNode_isSetIndex(n)) { d->failed = 1; return -1; }
      idx = (int)ASTNode_getIndex(n);
      switch (k->kind[idx]) {
      case K_ODE: return dag_leaf(d, idx, 0);
      case K_ASSIGNED: if (k->local[idx]) { d->failed = 1; return -1; } return dag_leaf(d, w_pv_off(d->w) + k->slot[idx], 0);   /* rule values are not refreshed here */
      case K_THREAD: return dag_leaf(d, w_pv_off(d->w) + k->slot[idx], 0);
      default: if (k->base) return dag_num(d, k->base[idx]); d->failed = 1; return -1;
      }
    }
  case AST_FUNCTION_EXP: return dag_op(d, D_EXP, dag_build(d, n->children[0]), -1);
  case AST_FUNCTION_LN: return dag_op(d, D_LOG, dag_build(d, n->children[0]), -1);
  default: d->failed = 1; return -1;        /* time, piecewise, other functions: the shape form evaluates those */
  }
}
static void emit_warp_flavour(charBuffer_t *b, const odeModel_t *om, kinfo_t *k, emit_ctx_t *c, int *ode_local, int usejac, int npv, const odeSense_t *os)
{
  int nb = om->nassbeforeodes, nlev = 0, lev, i, j, g, ngroups = 0, nvalues = om->neq + om->nass + om->nconst;
  int *depth = (int *)calloc((size_t)nb + 1, sizeof(int)), *refs = (int *)calloc((size_t)nvalues + 1, sizeof(int));
  int *group = (int *)calloc((size_t)nb + 1, sizeof(int));            /* table group of rule i, or -1 */
  char **text = (char **)calloc((size_t)nb + 1, sizeof(char *));
  struct leafrec *leaves = (struct leafrec *)calloc((size_t)nb + 1, sizeof(struct leafrec));
  int glevel[W_MAXGROUPS], grep_[W_MAXGROUPS], gcount[W_MAXGROUPS];
  int *odeoff = (int *)calloc((size_t)om->neq * W_ODE_TERMS_MAX + 1, sizeof(int)), *odeneg = (int *)calloc((size_t)om->neq * W_ODE_TERMS_MAX + 1, sizeof(int));
  int *odecnt = (int *)calloc((size_t)om->neq + 1, sizeof(int)); double *odediv = (double *)calloc((size_t)om->neq + 1, sizeof(double));
  int odeterms = 0, odefallback = 0, odedivide = 0, negzero;
  char **je_text = NULL; struct leafrec *je_leaf = NULL; int *je_dest = NULL, *je_neg = NULL, *je_shape = NULL, je_n = 0, je_shapes = 0, je_total = 0;
  dag_t dag; int *je_node = NULL, *je_dneg = NULL, dag_ok = 0, ntemp = 0, dag_one = 0;
  memset(&dag, 0, sizeof dag);
  wspace_t w;
  w.npv = npv; w.nass = om->nass; w.neq = om->neq; w.nlit = 0; w.lit = NULL; w.k = k;
  w.naspad = 8 * ((om->nass + 8) / 8); w.npvpad = 8 * ((npv + 7) / 8) + 8;
  g_wspace = &w;
  k->local = ode_local; c->interp = 0;
  negzero = w_literal(&w, -0.0);

  /* dependency depth of the rules the right-hand sides need */
  for (i = 0; i < nb; i++) {
    int d = 0, r;
    memset(refs, 0, sizeof(int) * (size_t)(nvalues + 1));
    mark_refs(om->assignmentsBeforeODEs[i]->ij, refs);
    for (r = 0; r < i; r++) if (refs[om->assignmentsBeforeODEs[r]->i] && depth[r] + 1 > d) d = depth[r] + 1;
    depth[i] = d;
    if (d + 1 > nlev) nlev = d + 1;
  }
  /* shapes: the text of each rule with operand placeholders */
  for (i = 0; i < nb; i++) {
    charBuffer_t *tb = CharBuffer_create(); emit_ctx_t lc = *c;
    lc.leaf = &leaves[i]; lc.name = name_leaf;
    gen(tb, om->assignmentsBeforeODEs[i]->ij, &lc);
    text[i] = tb->b; tb->b = NULL; CharBuffer_free(tb);
    group[i] = -1;
    if (leaves[i].failed || leaves[i].n > W_GROUP_LEAVES) { free(text[i]); text[i] = NULL; }
  }
  for (i = 0; i < nb; i++) {
    int members = 1;
    if (!text[i] || group[i] >= 0) continue;
    for (j = i + 1; j < nb; j++) if (text[j] && group[j] < 0 && depth[j] == depth[i] && strcmp(text[i], text[j]) == 0) members++;
    if (members < 2 || ngroups >= W_MAXGROUPS) continue;
    glevel[ngroups] = depth[i]; grep_[ngroups] = i; gcount[ngroups] = 0;
    for (j = i; j < nb && gcount[ngroups] < 32; j++)
      if (text[j] && group[j] < 0 && depth[j] == depth[i] && strcmp(text[i], text[j]) == 0) { group[j] = ngroups; gcount[ngroups]++; }
    ngroups++;
  }
  /* right-hand sides that are signed sums of words, optionally divided by a constant */
  for (i = 0; i < om->neq; i++) {
    const ASTNode_t *n = om->ode[i]; double div = 1.0; int cnt;
    if (n->type == AST_DIVIDE && n->nchildren == 2) {
      const ASTNode_t *d = n->children[1]; int ok = 0;
      if (is_const_tree(d)) { div = evaluateAST((ASTNode_t *)d, NULL); ok = 1; }
      else if (d->type == AST_NAME && ASTNode_isSetIndex(d) && k->kind[ASTNode_getIndex(d)] == K_UNIFORM && k->base) { div = k->base[ASTNode_getIndex(d)]; ok = 1; }
      if (ok) n = n->children[0];
    }
    cnt = w_flatten_sum(n, 0, &w, odeoff + i * W_ODE_TERMS_MAX, odeneg + i * W_ODE_TERMS_MAX, W_ODE_TERMS_MAX);
    if (cnt < 1 || i >= 32) { odecnt[i] = 0; odefallback = 1; continue; }
    odecnt[i] = cnt; odediv[i] = div;
    if (cnt > odeterms) odeterms = cnt;
    if (div != 1.0) odedivide = 1;
  }
  /* ---- Jacobian and parametric-matrix ENTRIES of identical shape, one entry per lane ----
     The row form below (w_jac_row, w_sens_row) makes lane i walk the expressions of row i: NEQ different code bodies one
     after the other, each run by one lane -- with the simultaneous corrector that is most of a trajectory's instructions
     (the Jacobian is evaluated at every corrector iteration).  Here every structurally non-zero entry d f_i / d y_j and
     d f_i / d p_k is a task of its own: a top-level negation is split off as a flag (-(x) is exact), repeated operands
     share a placeholder, entries whose text is then identical form a shape, and a lane evaluates the entry its task words
     name -- all entries of a shape at once, every one with exactly the operations of its emitted expression. */
  {
    int total = usejac ? om->sparsesize : 0, e;
    if (usejac && SENS_OK(os)) for (i = 0; i < om->neq; i++) for (j = 0; j < os->nsensP; j++) if (os->sensLogic[i][j]) total++;
    je_n = 0; je_shapes = 0;
    if (usejac && total > 0 && total <= 32 * JE_MAX_ROUNDS) {
      je_total = total;
      je_text = (char **)calloc((size_t)total, sizeof(char *)); je_leaf = (struct leafrec *)calloc((size_t)total, sizeof(struct leafrec));
      je_dest = (int *)calloc((size_t)total, sizeof(int)); je_neg = (int *)calloc((size_t)total, sizeof(int)); je_shape = (int *)calloc((size_t)total, sizeof(int));
      for (e = 0; e < total; e++) {
        const ASTNode_t *ast; charBuffer_t *tb = CharBuffer_create(); emit_ctx_t lc = *c; int q;
        if (e < om->sparsesize) { ast = om->jacobSparse[e]->ij; je_dest[e] = 32 * om->jacobSparse[e]->j + om->jacobSparse[e]->i; }
        else {
          int cnt = om->sparsesize, found = 0;
          ast = NULL;
          for (i = 0; i < om->neq && !found; i++) for (j = 0; j < os->nsensP && !found; j++) if (os->sensLogic[i][j]) { if (cnt == e) { ast = os->sens[i][j]; je_dest[e] = 32 * om->neq + 32 * j + i; found = 1; } cnt++; }
        }
        if (ast && ast->type == AST_MINUS && ast->nchildren == 1) { je_neg[e] = 1; ast = ast->children[0]; }
        je_leaf[e].dedup = 1; lc.leaf = &je_leaf[e]; lc.name = name_leaf;
        if (ast) gen(tb, ast, &lc); else je_leaf[e].failed = 1;
        je_text[e] = tb->b; tb->b = NULL; CharBuffer_free(tb);
        for (q = 0; q < je_leaf[e].n; q++) if (!je_leaf[e].isnum[q] && je_leaf[e].off[q] >= w_as_off(&w) && je_leaf[e].off[q] < w_pv_off(&w)) je_leaf[e].failed = 1;   /* rule values are not refreshed here */
        if (je_leaf[e].failed) break;
        for (q = 0; q < e; q++) if (strcmp(je_text[q], je_text[e]) == 0 && je_leaf[q].n == je_leaf[e].n) break;
        je_shape[e] = q < e ? je_shape[q] : je_shapes++;
      }
      if (e == total && je_shapes <= 64) {
        int q2;
        je_n = total;
        /* the same entries as a DAG (falls back to the shapes when a tree holds something the DAG has no operation for) */
        dag.w = &w; dag.k = k; dag.c = c;
        dag_one = w_lit_off(&w) + w_literal(&w, 1.0);            /* idle lanes of a step compute 1.0 (op) 1.0 into a scrap word */
        je_node = (int *)calloc((size_t)total, sizeof(int)); je_dneg = (int *)calloc((size_t)total, sizeof(int));
        for (q2 = 0; q2 < total && !dag.failed; q2++) {
          const ASTNode_t *ast = NULL; int cnt2 = om->sparsesize, ii, jj;
          if (q2 < om->sparsesize) ast = om->jacobSparse[q2]->ij;
          else for (ii = 0; ii < om->neq && !ast; ii++) for (jj = 0; jj < os->nsensP && !ast; jj++) if (os->sensLogic[ii][jj]) { if (cnt2 == q2) ast = os->sens[ii][jj]; cnt2++; }
          je_node[q2] = ast ? dag_build(&dag, ast) : -1;
          if (je_node[q2] < 0) dag.failed = 1;
          else if (dag.n[je_node[q2]].op == D_NEG) { je_dneg[q2] = 1; je_node[q2] = dag.n[je_node[q2]].a; }
        }
        dag_ok = !dag.failed;
        for (e = 0; e < total; e++) for (j = 0; j < je_leaf[e].n; j++) if (je_leaf[e].isnum[j]) je_leaf[e].off[j] = w_lit_off(&w) + w_literal(&w, je_leaf[e].num[j]);
      }
    }
  }
  /* pool the literals of the grouped rules */
  for (i = 0; i < nb; i++) if (group[i] >= 0) for (j = 0; j < leaves[i].n; j++) if (leaves[i].isnum[j]) leaves[i].off[j] = w_lit_off(&w) + w_literal(&w, leaves[i].num[j]);

  /* (the literal pool is complete from here on: the DAG's temporaries live behind it) */
  cbprintf(b, "#ifndef ODES_WARP\n#define ODES_WARP 0\n#endif\n#if ODES_WARP\n#define W_NLEVELS %d\n#define W_AS_OFF %d\n#define W_PV_OFF %d\n#define W_LIT_OFF %d\n#define W_NLIT %d\n#define W_NGROUPS %d\n"
              "#define W_ODE_NTERMS %d\n#define W_ODE_DIVIDE %d\n#define W_ODE_FALLBACK %d\n#define W_NEGZERO_OFF %d\n",
           nlev, w_as_off(&w), w_pv_off(&w), w_lit_off(&w), w.nlit, ngroups, odeterms, odedivide, odefallback, w_lit_off(&w) + negzero);
  CharBuffer_append(b, "DEV double w_literal(int n)\n{\n  switch (n) {\n");
  for (i = 0; i < w.nlit; i++) { cbprintf(b, "  case %d: return ", i); if (i == negzero) CharBuffer_append(b, "-0.0"); else cb_double17(b, w.lit[i]); CharBuffer_append(b, ";\n"); }
  CharBuffer_append(b, "  default: return 0.0;\n  }\n}\n");
  /* packed operand words: bit 63 valid, bits 50..59 destination, 10 bits per operand from bit 0 */
  CharBuffer_append(b, "DEV unsigned long long w_group_task(int g, int lane)\n{\n  (void)g; (void)lane;\n  switch (g * 32 + lane) {\n");
  for (g = 0; g < ngroups; g++) {
    int lane = 0;
    for (i = 0; i < nb; i++) {
      unsigned long long word = 1ull << 63;
      if (group[i] != g) continue;
      word |= (unsigned long long)(w_as_off(&w) + om->assignmentsBeforeODEs[i]->i - om->neq) << 50;
      for (j = 0; j < leaves[i].n; j++) word |= (unsigned long long)leaves[i].off[j] << (10 * j);
      cbprintf(b, "  case %d: return 0x%llxull;\n", g * 32 + lane, word);
      lane++;
    }
  }
  CharBuffer_append(b, "  default: return 0ull;\n  }\n}\n");
  CharBuffer_append(b, "DEV void w_groups_level(int level, const unsigned long long (&task)[A1(W_NGROUPS)], double t, double *V)\n{\n  (void)level; (void)task; (void)t; (void)V;\n  switch (level) {\n");
  for (lev = 0; lev < nlev; lev++) {
    cbprintf(b, "  case %d:\n", lev);
    for (g = 0; g < ngroups; g++) {
      if (glevel[g] != lev) continue;
      cbprintf(b, "    if (task[%d] >> 63) {\n      const unsigned long long w_ = task[%d];\n", g, g);
      for (j = 0; j < leaves[grep_[g]].n; j++) cbprintf(b, "      const double x%d = V[(w_ >> %d) & 1023u];\n", j, 10 * j);
      cbprintf(b, "      V[(w_ >> 50) & 1023u] = %s;\n    }\n", text[grep_[g]]);
    }
    CharBuffer_append(b, "    break;\n");
  }
  CharBuffer_append(b, "  default: break;\n  }\n}\n");
  /* rules outside the groups: per lane */
  CharBuffer_append(b, "DEV void w_rules_level(int level, int lane, double t, const double *y, double *a, const PvRef pv)\n{\n  (void)level; (void)lane; (void)t; (void)y; (void)a; (void)pv;\n  switch (level) {\n");
  for (lev = 0; lev < nlev; lev++) {
    int slot, any = 0;
    for (i = 0; i < nb; i++) if (depth[i] == lev && group[i] < 0) any = 1;
    cbprintf(b, "  case %d:\n", lev);
    if (any) {
      CharBuffer_append(b, "    switch (lane) {\n");
      for (slot = 0; slot < 32; slot++) {
        int open = 0, cnt = 0;
        for (i = 0; i < nb; i++) {
          if (depth[i] != lev || group[i] >= 0) continue;
          if (cnt % 32 == slot) {
            nonzeroElem_t *o = om->assignmentsBeforeODEs[i];
            if (!open) { cbprintf(b, "    case %d:", slot); open = 1; }
            cbprintf(b, " a[%d] = ", o->i - om->neq); gen(b, o->ij, c); CharBuffer_append(b, ";");
          }
          cnt++;
        }
        if (open) CharBuffer_append(b, " break;\n");
      }
      CharBuffer_append(b, "    default: break;\n    }\n");
    }
    CharBuffer_append(b, "    break;\n");
  }
  CharBuffer_append(b, "  default: break;\n  }\n}\n");
  { int any = 0; for (i = 0; i < nb; i++) if (group[i] < 0) any = 1; cbprintf(b, "#define W_RULES_FALLBACK %d\n", any); }
  /* right-hand sides: term lists (bit 63 valid; 11 bits per term: offset, sign), divisors, and the per-lane rest */
  cbprintf(b, "#define W_ODE_WORDS %d\n", odeterms > 5 ? 2 : 1);
  CharBuffer_append(b, "DEV unsigned long long w_ode_task(int lane, int word)\n{\n  switch (2 * lane + word) {\n");
  for (i = 0; i < om->neq; i++) {
    int wd;
    if (!odecnt[i]) continue;
    for (wd = 0; wd < (odeterms > 5 ? 2 : 1); wd++) {
      unsigned long long word = 1ull << 63;
      for (j = 5 * wd; j < odeterms && j < 5 * wd + 5; j++) {
        const int off = j < odecnt[i] ? odeoff[i * W_ODE_TERMS_MAX + j] : w_lit_off(&w) + negzero, neg = j < odecnt[i] ? odeneg[i * W_ODE_TERMS_MAX + j] : 0;
        word |= (unsigned long long)(off | (neg << 10)) << (11 * (j - 5 * wd));
      }
      cbprintf(b, "  case %d: return 0x%llxull;\n", 2 * i + wd, word);
    }
  }
  CharBuffer_append(b, "  default: return 0ull;\n  }\n}\nDEV double w_ode_divisor(int lane)\n{\n  switch (lane) {\n");
  for (i = 0; i < om->neq; i++) if (odecnt[i] && odediv[i] != 1.0) { cbprintf(b, "  case %d: return ", i); cb_double17(b, odediv[i]); CharBuffer_append(b, ";\n"); }
  CharBuffer_append(b, "  default: return 1.0;\n  }\n}\n");
  CharBuffer_append(b, "DEV double w_ode_case(int i, double t, const double *y, const double *a, const PvRef pv)\n{\n  (void)t; (void)y; (void)a; (void)pv;\n  switch (i) {\n");
  for (i = 0; i < om->neq; i++) { if (odecnt[i]) continue; cbprintf(b, "  case %d: return ", i); gen(b, om->ode[i], c); CharBuffer_append(b, ";\n"); }
  CharBuffer_append(b, "  default: return 0.0;\n  }\n}\n");
  CharBuffer_append(b, "#define JR(j) jr[32 * (j) + lane]\nDEV void w_jac_row(int lane, double t, const double *y, const PvRef pv, double *jr)\n{\n  (void)t; (void)y; (void)pv; (void)jr;\n  switch (lane) {\n");
  if (usejac) {
    for (i = 0; i < om->neq; i++) {
      int open = 0;
      for (j = 0; j < om->sparsesize; j++) {
        if (om->jacobSparse[j]->i != i) continue;
        if (!open) { cbprintf(b, "  case %d:\n", i); open = 1; }
        cbprintf(b, "    JR(%d) = ", om->jacobSparse[j]->j); gen(b, om->jacobSparse[j]->ij, c); CharBuffer_append(b, ";\n");
      }
      if (open) CharBuffer_append(b, "    break;\n");
    }
  }
  CharBuffer_append(b, "  default: break;\n  }\n}\n");
  /* forward sensitivities: which variable a sensitivity belongs to (initial value 1 on that element), and the emitted
     sense_f row by row -- 0, += J_ij * yS[j] over the row's non-zero Jacobian entries, += df_i/dp_is */
  if (os) {
    CharBuffer_append(b, "DEV int w_sens_variable(int is)\n{\n  switch (is) {\n");
    for (i = 0; i < os->nsens; i++) if (os->index_sensP[i] == -1) cbprintf(b, "  case %d: return %d;\n", i, os->index_sens[i]);
    CharBuffer_append(b, "  default: return -1;\n  }\n}\n");
    /* the per-trajectory slot of a sensitivity parameter (the difference quotients perturb it there), -1 for a variable */
    CharBuffer_append(b, "DEV int w_sens_pvslot(int is)\n{\n  switch (is) {\n");
    for (i = 0; i < os->nsens; i++) if (os->index_sensP[i] != -1 && k->slot[os->index_sens[i]] >= 0) cbprintf(b, "  case %d: return %d;\n", i, k->slot[os->index_sens[i]]);
    CharBuffer_append(b, "  default: return -1;\n  }\n}\n");
    CharBuffer_append(b, "DEV double w_sens_row(int lane, int is, double t, const double *y, const PvRef pv, const double *jr, const double *ys)\n{\n"
                         "  double r = 0.0; (void)is; (void)t; (void)y; (void)pv; (void)jr; (void)ys;\n  switch (lane) {\n");
    for (i = 0; i < om->neq; i++) {
      int kk, any = 0;
      cbprintf(b, "  case %d:\n", i);
      if (usejac) for (j = 0; j < om->sparsesize; j++) if (om->jacobSparse[j]->i == i) cbprintf(b, "    r += JR(%d) * ys[%d];\n", om->jacobSparse[j]->j, om->jacobSparse[j]->j);
      for (kk = 0; kk < os->nsens; kk++) if (os->index_sensP[kk] != -1 && usejac && SENS_OK(os) && os->sensLogic[i][os->index_sensP[kk]]) any = 1;
      if (any) {
        CharBuffer_append(b, "    switch (is) {\n");
        for (kk = 0; kk < os->nsens; kk++) {
          if (os->index_sensP[kk] == -1 || !os->sensLogic[i][os->index_sensP[kk]]) continue;
          cbprintf(b, "    case %d: r += ", kk); gen(b, os->sens[i][os->index_sensP[kk]], c); CharBuffer_append(b, "; break;\n");
        }
        CharBuffer_append(b, "    default: break;\n    }\n");
      }
      CharBuffer_append(b, "    break;\n");
    }
    CharBuffer_append(b, "  default: break;\n  }\n  return r;\n}\n");
  }
  /* the entry tasks: word 0 = operands 0..3 (10 bits each), destination (bits 40..51: 32 * column + row in the row-per-lane
     matrix, the parametric columns behind the NEQ Jacobian columns), shape (52..57), negate (62), valid (63); words 1, 2 =
     operands 4..9 and 10..15.  Slots are filled shape by shape so that a round of 32 entries holds few shapes. */
  cbprintf(b, "#define W_JE_N %d\n#define W_JE_ROUNDS %d\n#define W_NSENSP %d\n", je_n, (je_n + 31) / 32, (usejac && SENS_OK(os)) ? os->nsensP : 0);
  if (je_n > 0) {
    int slot = 0, sh, e, q;
    CharBuffer_append(b, "DEV unsigned long long w_je_task(int slot, int word)\n{\n  switch (3 * slot + word) {\n");
    for (sh = 0; sh < je_shapes; sh++)
      for (e = 0; e < je_n; e++) {
        unsigned long long wd[3] = {0ull, 0ull, 0ull};
        if (je_shape[e] != sh) continue;
        wd[0] = (1ull << 63) | ((unsigned long long)je_neg[e] << 62) | ((unsigned long long)sh << 52) | ((unsigned long long)je_dest[e] << 40);
        for (q = 0; q < je_leaf[e].n; q++) {
          if (q < 4) wd[0] |= (unsigned long long)je_leaf[e].off[q] << (10 * q);
          else wd[1 + (q - 4) / 6] |= (unsigned long long)je_leaf[e].off[q] << (10 * ((q - 4) % 6));
        }
        for (q = 0; q < 3; q++) if (wd[q]) cbprintf(b, "  case %d: return 0x%llxull;\n", 3 * slot + q, wd[q]);
        slot++;
      }
    CharBuffer_append(b, "  default: return 0ull;\n  }\n}\n");
    CharBuffer_append(b, "DEV void w_je_eval(const unsigned long long w0, const unsigned long long w1, const unsigned long long w2, double t, const double *V, double *dst, unsigned int limit)\n{\n"
                         "  (void)t; (void)w1; (void)w2;\n  const unsigned int d_ = (unsigned int)(w0 >> 40) & 4095u;\n  if (!(w0 >> 63) || d_ >= limit) return;\n  double r_ = 0.0;\n  switch ((int)((w0 >> 52) & 63u)) {\n");
    for (sh = 0; sh < je_shapes; sh++) {
      for (e = 0; e < je_n && je_shape[e] != sh; e++) ;
      cbprintf(b, "  case %d: {\n", sh);
      for (q = 0; q < je_leaf[e].n; q++) {
        if (q < 4) cbprintf(b, "    const double x%d = V[(w0 >> %d) & 1023u];\n", q, 10 * q);
        else cbprintf(b, "    const double x%d = V[(w%d >> %d) & 1023u];\n", q, 1 + (q - 4) / 6, 10 * ((q - 4) % 6));
      }
      cbprintf(b, "    r_ = %s;\n  } break;\n", je_text[e]);
    }
    CharBuffer_append(b, "  default: break;\n  }\n  dst[d_] = ((w0 >> 62) & 1ull) ? -r_ : r_;\n}\n");
    /* rows of the sensitivity right-hand side: the non-zero Jacobian columns of a row (bit j), the parametric column of a
       sensitivity (or -1: it belongs to a variable) and the non-zero parametric columns of a row */
    CharBuffer_append(b, "DEV unsigned int w_row_mask(int i)\n{\n  switch (i) {\n");
    for (i = 0; i < om->neq && i < 32; i++) {
      unsigned int m = 0u; int last = -1, sorted = 1;
      for (j = 0; j < om->sparsesize; j++) if (om->jacobSparse[j]->i == i) { if (om->jacobSparse[j]->j <= last) sorted = 0; last = om->jacobSparse[j]->j; m |= 1u << om->jacobSparse[j]->j; }
      cbprintf(b, "  case %d: return 0x%xu;%s\n", i, m, sorted ? "" : " /* UNSORTED */");
      if (!sorted) je_n = -1;
    }
    CharBuffer_append(b, "  default: return 0u;\n  }\n}\n");
    CharBuffer_append(b, "DEV int w_sens_pcol(int is)\n{\n  switch (is) {\n");
    if (os) for (i = 0; i < os->nsens; i++) if (os->index_sensP[i] != -1) cbprintf(b, "  case %d: return %d;\n", i, os->index_sensP[i]);
    CharBuffer_append(b, "  default: return -1;\n  }\n}\nDEV unsigned int w_sens_pmask(int i)\n{\n  switch (i) {\n");
    if (usejac && SENS_OK(os)) for (i = 0; i < om->neq; i++) {
      unsigned int m = 0u;
      for (j = 0; j < os->nsensP && j < 32; j++) if (os->sensLogic[i][j]) m |= 1u << j;
      if (m) cbprintf(b, "  case %d: return 0x%xu;\n", i, m);
    }
    CharBuffer_append(b, "  default: return 0u;\n  }\n}\n");
    if (je_n < 0) CharBuffer_append(b, "#undef W_JE_N\n#define W_JE_N 0\n");          /* column order of a row not ascending: keep the row form */
  }
  /* the DAG form: steps (level, operation) of at most 32 tasks; a task is one 32-bit word, operand a | operand b << 10 |
     result << 20, offsets into V (temporaries behind the literal pool; idle lanes take 1.0 (op) 1.0 into a scrap word, so
     a step needs no predicate); the operation of a step carries bit 8 when the next step belongs to a higher level (the
     warp synchronises there); then the entries: node | destination << 12 | negate << 62 | valid << 63 */
  {
    int nsteps = 0, opmask = 0, lv, op, maxlv = 0, q;
    if (dag_ok && je_n > 0) {
      const int tmp_off = 8 * ((w_lit_off(&w) + w.nlit + 7) / 8);
      ntemp = 1;                                                 /* word 0: scrap */
      for (q = 0; q < dag.cnt; q++) { if (dag.n[q].op != D_LEAF) dag.n[q].off = tmp_off + ntemp++; if (dag.n[q].level > maxlv) maxlv = dag.n[q].level; }
      if (tmp_off + ntemp >= 1024) dag_ok = 0;
    }
    if (dag_ok && je_n > 0) {
      const int tmp_off = 8 * ((w_lit_off(&w) + w.nlit + 7) / 8);
      const unsigned int idle = (unsigned int)dag_one | ((unsigned int)dag_one << 10) | ((unsigned int)tmp_off << 20);
      charBuffer_t *tasks = CharBuffer_create(), *ops = CharBuffer_create();
      int *stepop = (int *)calloc(1024, sizeof(int)), *steplv = (int *)calloc(1024, sizeof(int));
      for (lv = 1; lv <= maxlv && nsteps < 1000; lv++) for (op = D_ADD; op < D_NOPS && nsteps < 1000; op++) {
        int lane = 0, any = 0;
        for (q = 0; q < dag.cnt && nsteps < 1000; q++) {
          unsigned int word;
          if (dag.n[q].op != op || dag.n[q].level != lv) continue;
          if (lane == 32) { stepop[nsteps] = op; steplv[nsteps] = lv; nsteps++; lane = 0; }
          word = (unsigned int)dag.n[dag.n[q].a].off | ((unsigned int)dag.n[dag.n[q].b].off << 10) | ((unsigned int)dag.n[q].off << 20);
          cbprintf(tasks, "  case %d: return 0x%xu;\n", 32 * nsteps + lane, word);
          lane++; any = 1; opmask |= 1 << op;
        }
        if (any) { stepop[nsteps] = op; steplv[nsteps] = lv; nsteps++; }
      }
      for (q = 0; q < nsteps; q++) cbprintf(ops, "  case %d: return %d;\n", q, stepop[q] | ((q + 1 == nsteps || steplv[q + 1] != steplv[q]) ? 256 : 0));
      if (nsteps > 0 && nsteps <= 192) {
        /* does the DAG pay?  instruction estimates: a shape body costs its operations (a division ~22 instructions, pow / exp /
           log a few hundred) plus operand fetches, a DAG step its one operation plus ~16 of task decoding and traffic -- but a
           step is also a dependent chain of two loads, the operation and a store, so the DAG must win by a factor (measured,
           see bdf_warp_kernel.cuh) */
        int shape_cost = 0, dag_cost = 0, sh;
        for (sh = 0; sh < je_shapes; sh++) {
          const char *t; int e2, nd = 0, no = 0, nf = 0;
          for (e2 = 0; e2 < je_n && je_shape[e2] != sh; e2++) ;
          for (t = je_text[e2]; *t; t++) { if (*t == '/') nd++; else if (*t == '+' || *t == '*' || (*t == '-' && t[1] == ' ')) no++; else if (*t == 'o' && (!strncmp(t, "odes_pow", 8) || !strncmp(t, "odes_exp", 8) || !strncmp(t, "odes_log", 8))) nf++; }
          shape_cost += 22 * nd + 2 * no + 250 * nf + 6 * je_leaf[e2].n + 12;
        }
        for (q = 0; q < nsteps; q++) dag_cost += (stepop[q] == D_DIV ? 38 : (stepop[q] == D_POW || stepop[q] == D_EXP || stepop[q] == D_LOG) ? 266 : 18);
        cbprintf(b, "#define W_DAG_STEPS %d\n#define W_DAG_OPMASK %d\n#define W_DAG_NTEMP %d\n#define W_DAG_NODES %d\n#define W_DAG_IDLE 0x%xu\n#define W_DAG_PAYS %d /* estimate: shapes %d, DAG %d instructions */\n",
                 nsteps, opmask, ntemp, dag.cnt, idle, shape_cost > 2 * dag_cost, shape_cost, dag_cost);
        CharBuffer_append(b, "DEV unsigned int w_dag_task(int slot)\n{\n  switch (slot) {\n"); if (tasks->b) CharBuffer_append(b, tasks->b);
        CharBuffer_append(b, "  default: return W_DAG_IDLE;\n  }\n}\nDEV int w_dag_op(int step)\n{\n  switch (step) {\n"); if (ops->b) CharBuffer_append(b, ops->b);
        CharBuffer_append(b, "  default: return 0;\n  }\n}\n");
        CharBuffer_append(b, "DEV unsigned long long w_dag_out(int slot)\n{\n  switch (slot) {\n");
        for (q = 0; q < je_n; q++)
          cbprintf(b, "  case %d: return 0x%llxull;\n", q, (1ull << 63) | ((unsigned long long)je_dneg[q] << 62) | ((unsigned long long)je_dest[q] << 12) | (unsigned long long)dag.n[je_node[q]].off);
        CharBuffer_append(b, "  default: return 0ull;\n  }\n}\n");
      } else dag_ok = 0;
      CharBuffer_free(tasks); CharBuffer_free(ops); free(stepop); free(steplv);
    }
    if (!(dag_ok && je_n > 0)) CharBuffer_append(b, "#define W_DAG_STEPS 0\n#define W_DAG_NTEMP 0\n");
  }
  CharBuffer_append(b, "#endif\n");
  for (i = 0; je_text && i < je_total; i++) free(je_text[i]);
  free(je_text); free(je_leaf); free(je_dest); free(je_neg); free(je_shape); free(je_node); free(je_dneg); free(dag.n);
  for (i = 0; i < nb; i++) free(text[i]);
  free(text); free(leaves); free(group); free(depth); free(refs); free(odeoff); free(odeneg); free(odecnt); free(odediv); free(w.lit);
}

/* Reaction fluxes as a device function.  The reference maps results after every run by evaluating each
   kinetic law -- local parameters and constants substituted -- with evaluateAST on the stored values of every time point
   (src/odeSolver.c:463-471, :519-531).  Here the same indexed trees (SBMLResults_fluxASTs, sbml_results.c) are emitted
   with the interpreter's arithmetic over one stored row v[0..NVAL) and run by odes_flux_kernel (bdf_kernel.cuh) over the
   trajectories in HBM.  Reactions without a kinetic law keep the zero of their freshly created time course. */
static void name_row(charBuffer_t *s, int idx, const emit_ctx_t *c) { (void)c; cbprintf(s, "v[%d]", idx); }
char *ODEModel_generateCUDAFluxSection(int nflux, ASTNode_t *const *kls)
{
  charBuffer_t *b = CharBuffer_create(); emit_ctx_t c; char *out; int r;
  c.cuda = 1; c.interp = 1; c.name = name_row; c.time = REF_TIME; c.info = NULL; c.leaf = NULL;
  cbprintf(b, "#define NFLUX %d\nDEV void flux_row(double t, const double *v, double *flux)\n{\n  (void)t; (void)v;\n", nflux);
  for (r = 0; r < nflux; r++) {
    cbprintf(b, "  flux[%d] = ", r);
    if (kls[r]) gen(b, kls[r], &c); else CharBuffer_append(b, "0.0");
    CharBuffer_append(b, ";\n");
  }
  CharBuffer_append(b, "}\n");
  out = b->b; b->b = NULL; CharBuffer_free(b);
  return out;
}

char *ODEModel_generateCUDASource(odeModel_t *om, const cvodeSettings_t *opt, int nvary, const int *vary,
                                  int nstore, const int *store, const double *base)
{ return ODEModel_generateCUDASourceSens(om, opt, nvary, vary, nstore, store, base, NULL); }
char *ODEModel_generateCUDASourceSens(odeModel_t *om, const cvodeSettings_t *opt, int nvary, const int *vary,
                                      int nstore, const int *store, const double *base, const odeSense_t *os)
{
  int nvalues = om->neq + om->nass + om->nconst, i, j, e, npv = 0, usejac, store_needs_pv = 0, static_slots = 0;
  charBuffer_t *b = CharBuffer_create();
  kinfo_t k; emit_ctx_t c; char *out;
  int *used_in_ode = (int *)calloc((size_t)nvalues + 1, sizeof(int));
  int *all_local = (int *)calloc((size_t)nvalues + 1, sizeof(int)), *ode_local = (int *)calloc((size_t)nvalues + 1, sizeof(int));

  if (opt->UseJacobian && !om->jacob && !om->jacobianFailed) ODEModel_constructJacobian(om);
  usejac = opt->UseJacobian && om->jacobian;

  k.om = om;
  k.base = base;
  k.kind = (int *)calloc((size_t)nvalues + 1, sizeof(int));
  k.slot = (int *)calloc((size_t)nvalues + 1, sizeof(int));
  k.local = all_local;
  for (i = 0; i < nvalues; i++) { k.kind[i] = i < om->neq ? K_ODE : i < om->neq + om->nass ? K_ASSIGNED : K_UNIFORM; k.slot[i] = -1; }
  /* per-trajectory constants: varied inputs and event-assignment targets */
  for (j = 0; j < nvary; j++) if (vary[j] >= om->neq + om->nass && k.kind[vary[j]] == K_UNIFORM) { k.kind[vary[j]] = K_THREAD; k.slot[vary[j]] = npv++; }
  for (e = 0; e < om->nevents; e++) for (j = 0; j < om->neventAss[e]; j++) {
    int idx = om->eventIndex[e][j];
    if (idx >= om->neq + om->nass && k.kind[idx] == K_UNIFORM) { k.kind[idx] = K_THREAD; k.slot[idx] = npv++; }
  }
  /* difference-quotient sensitivities (no parametric matrix or no Jacobian, src/sensSolver.c:312-321): the kernel perturbs
     the sensitivity parameters (CVSensRhs1DQ), so they are per-trajectory constants even where the batch does not vary them */
  if (os && !(opt->UseJacobian && om->jacobian && os->sensitivity))
    for (i = 0; i < os->nsens; i++) {
      int idx = os->index_sens[i];
      if (os->index_sensP[i] != -1 && idx >= om->neq + om->nass && k.kind[idx] == K_UNIFORM) { k.kind[idx] = K_THREAD; k.slot[idx] = npv++; }
    }
  /* assigned values read by the ODEs but not recomputed before them keep a per-trajectory slot */
  for (i = 0; i < om->neq; i++) mark_refs(om->ode[i], used_in_ode);
  for (i = 0; i < om->nassbeforeodes; i++) { ode_local[om->assignmentsBeforeODEs[i]->i] = 1; mark_refs(om->assignmentsBeforeODEs[i]->ij, used_in_ode); }
  for (i = om->neq; i < om->neq + om->nass; i++) { all_local[i] = 1; if (used_in_ode[i] && !ode_local[i]) { k.slot[i] = npv++; static_slots++; } }

  c.cuda = 1; c.interp = 0; c.name = name_cuda; c.time = REF_TIME; c.info = &k; c.leaf = NULL;

  cbprintf(b, "/* generated by ODEModel_generateCUDASource for model '%s' (sm_100a device functions) */\n",
           om->simple && om->simple->id ? om->simple->id : "odes");
  cbprintf(b, "#define NEQ %d\n#define NASS %d\n#define NVAL %d\n#define NEVENTS %d\n#define NPV %d\n#define NVARY %d\n#define NSTORE %d\n#define NNZ %d\n#define HAS_JAC %d\n#define USE_TSTOP %d\n",
           om->neq, om->nass, nvalues, om->nevents, npv, nvary, nstore, usejac ? om->sparsesize : 0, usejac, (opt->SetTStop || om->npiecewise) ? 1 : 0);
  cbprintf(b, "#define RESET_ON_EVENT %d\n#define HALT_ON_EVENT %d\n", opt->ResetCvodeOnEvent ? 1 : 0, opt->HaltOnEvent ? 1 : 0);
  cbprintf(b, "#define NSENS %d\n#define SENS_STAGGERED %d\n#define W_SENS_DQ %d\n#define W_STATIC_RULE_SLOTS %d\n", os ? os->nsens : 0, os ? opt->SensMethod : 0, (os && !(usejac && os->sensitivity)) ? 1 : 0, static_slots);
  CharBuffer_append(b, "#define A1(n) ((n) > 0 ? (n) : 1)\n#define DEV __device__ __forceinline__\n");
  /* per-trajectory constants live in strided shared memory (word k of this thread at p[k*s]) */
  CharBuffer_append(b, "#ifndef ODES_BLOCK\n#define ODES_BLOCK 64\n#endif\n#ifndef ODES_MAT_LOCAL\n#define ODES_MAT_LOCAL 0\n#endif\n"
                       "#if defined(ODES_HOST_EMULATION) || ODES_MAT_LOCAL\n#define ODES_SSTRIDE 1\n#else\n#define ODES_SSTRIDE ODES_BLOCK\n#endif\n"
                       "struct PvRef { double *p; unsigned int s; DEV double &operator[](int k) const { return p[(size_t)k * s]; } };\n");
  CharBuffer_append(b, "__constant__ double c_base[A1(NVAL)];\n__constant__ int c_trigger0[A1(NEVENTS)];\n");
  /* pow as the host libm computes it (glibc_pow.inc, defined after this section) */
  CharBuffer_append(b, "DEV double odes_pow(double x, double y);\nDEV double odes_exp(double x);\nDEV double odes_log(double x);\n#define ODES_POW odes_pow\n");
  generateMacros(b);
  /* scatter of the saved sparse Jacobian into column j of M = -gamma*J (the kernel adds the +1 on the diagonal);
     j is warp-uniform at the call site, rows and non-zero indices are compile-time constants */
  /* ODES_LA_PACKED (set by the host where it halves the tensor-memory allocation of a lane): the per-trajectory constants
     follow the saved Jacobian's non-zeros in the SAME chunks instead of starting a padded region of their own */
  CharBuffer_append(b, "#ifndef ODES_LA_PACKED\n#define ODES_LA_PACKED 0\n#endif\n#define NP (8 * ((NEQ + 7) / 8))\n#define NSJW (HAS_JAC ? NNZ : NEQ * NP)\n"
                       "#if ODES_LA_PACKED && HAS_JAC\n#define NSJP (8 * ((NSJW + NPV + 7) / 8))\n#define NPVP 0\n#else\n#define NSJP (8 * ((NSJW + 7) / 8))\n#define NPVP (8 * ((NPV + 7) / 8))\n#endif\n");
  CharBuffer_append(b, "DEV void scatter_col(int j, double (&c)[A1(NP)], const double (&jnz)[A1(NSJP)], double mg)\n{\n  (void)c; (void)jnz; (void)mg;\n  switch (j) {\n");
  if (usejac) {
    for (j = 0; j < om->neq; j++) {
      cbprintf(b, "  case %d:", j);
      for (i = 0; i < om->sparsesize; i++) if (om->jacobSparse[i]->j == j) cbprintf(b, " c[%d] = jnz[%d] * mg;", om->jacobSparse[i]->i, i);
      CharBuffer_append(b, " break;\n");
    }
  }
  CharBuffer_append(b, "  default: break;\n  }\n}\n");

  /* model_init: uniform base values, then the varied inputs of this trajectory (SoA, batch fastest) */
  CharBuffer_append(b, "DEV void model_init(double (&y)[A1(NEQ)], const PvRef pv, const double *__restrict__ vary, size_t stride, size_t set)\n{\n");
  for (i = 0; i < om->neq; i++) cbprintf(b, "  y[%d] = c_base[%d];\n", i, i);
  for (i = 0; i < nvalues; i++) if (k.slot[i] >= 0) cbprintf(b, "  pv[%d] = c_base[%d];\n", k.slot[i], i);
  for (j = 0; j < nvary; j++) {
    int idx = vary[j];
    if (idx < om->neq) cbprintf(b, "  y[%d] = vary[%d * stride + set];\n", idx, j);
    else if (k.kind[idx] == K_THREAD) cbprintf(b, "  pv[%d] = vary[%d * stride + set];\n", k.slot[idx], j);
  }
  CharBuffer_append(b, "  (void)vary; (void)stride; (void)set; (void)pv; (void)y;\n}\n");

  /* refresh of all assignment rules; static ones are written back to their slots */
  CharBuffer_append(b, "DEV void rules_all(double t, const double (&y)[A1(NEQ)], const PvRef pv, double (&a)[A1(NASS)])\n{\n  (void)t; (void)y; (void)pv; (void)a;\n");
  k.local = all_local; c.interp = 1;
  emit_refresh_all(b, om, &c, "  ");
  for (i = om->neq; i < om->neq + om->nass; i++) if (k.slot[i] >= 0) cbprintf(b, "  pv[%d] = a[%d];\n", k.slot[i], i - om->neq);
  CharBuffer_append(b, "}\n");

  /* ode_f: rules needed by the ODEs in topological order, then the right-hand sides */
  CharBuffer_append(b, "DEV void ode_f(double t, const double (&y)[A1(NEQ)], const PvRef pv, double (&ydot)[A1(NEQ)])\n{\n  double a[A1(NASS)]; (void)t; (void)pv; (void)a;\n");
  k.local = ode_local; c.interp = 0;
  for (i = 0; i < om->nassbeforeodes; i++) {
    nonzeroElem_t *o = om->assignmentsBeforeODEs[i];
    cbprintf(b, "  a[%d] = ", o->i - om->neq); gen(b, o->ij, &c); CharBuffer_append(b, ";\n");
  }
  for (i = 0; i < om->neq; i++) { cbprintf(b, "  ydot[%d] = ", i); gen(b, om->ode[i], &c); CharBuffer_append(b, ";\n"); }
  CharBuffer_append(b, "}\n");

  /* jacobi_f: structurally non-zero entries in jacobSparse order (rules are inlined) */
  CharBuffer_append(b, "DEV void jacobi_f(double t, const double (&y)[A1(NEQ)], const PvRef pv, double (&jnz)[A1(NNZ)])\n{\n  (void)t; (void)pv; (void)y; (void)jnz;\n");
  k.local = ode_local;
  if (usejac) for (i = 0; i < om->sparsesize; i++) { cbprintf(b, "  jnz[%d] = ", i); gen(b, om->jacobSparse[i]->ij, &c); CharBuffer_append(b, ";\n"); }
  CharBuffer_append(b, "}\n");

  emit_warp_flavour(b, om, &k, &c, ode_local, usejac, npv, os);

  /* event_f: interpreter semantics of IntegratorInstance_processEventsAndAssignments */
  CharBuffer_append(b, "DEV unsigned int event_f(double t, double (&y)[A1(NEQ)], const PvRef pv, int (&trigger)[A1(NEVENTS)], int &reinit)\n{\n"
                       "  unsigned int fired = 0; double a[A1(NASS)]; (void)t; (void)a; (void)y; (void)pv; (void)trigger; (void)reinit;\n");
  k.local = all_local; c.interp = 1;
  if (om->nevents > 0) {
    emit_refresh_all(b, om, &c, "  ");
    for (e = 0; e < om->nevents; e++) {
      cbprintf(b, "  if (trigger[%d] == 0) {\n    if ((", e); gen(b, om->event[e], &c); cbprintf(b, ") != 0.0) {\n      fired |= %uu; trigger[%d] = 1;\n", 1u << e, e);
      for (j = 0; j < om->neventAss[e]; j++) {
        int idx = om->eventIndex[e][j];
        if (idx >= om->neq && idx < om->neq + om->nass) continue;       /* assigned values cannot be set */
        CharBuffer_append(b, "      { const double r = "); gen(b, om->eventAssignment[e][j], &c); CharBuffer_append(b, "; ");
        if (idx < om->neq) cbprintf(b, "y[%d] = r; reinit = 1; }\n", idx);
        else cbprintf(b, "pv[%d] = r; if (RESET_ON_EVENT) reinit = 1; }\n", k.slot[idx]);
        emit_refresh_all(b, om, &c, "      ");
      }
      cbprintf(b, "    }\n  } else if (!((", e); gen(b, om->event[e], &c); cbprintf(b, ") != 0.0)) trigger[%d] = 0;\n", e);
    }
    for (i = om->neq; i < om->neq + om->nass; i++) if (k.slot[i] >= 0) cbprintf(b, "  pv[%d] = a[%d];\n", k.slot[i], i - om->neq);
  }
  CharBuffer_append(b, "  return fired;\n}\n");
  /* trigger_f: the truth values of all event triggers at (t, y) as a bit mask, evaluated like event_f evaluates them but
     without assignments and without touching the trigger state (event location inside a step, ODES_ROOTFIND) */
  CharBuffer_append(b, "DEV unsigned int trigger_f(double t, const double (&y)[A1(NEQ)], const PvRef pv)\n{\n"
                       "  unsigned int m = 0; double a[A1(NASS)]; (void)t; (void)a; (void)y; (void)pv;\n");
  if (om->nevents > 0) {
    emit_refresh_all(b, om, &c, "  ");
    for (e = 0; e < om->nevents; e++) { CharBuffer_append(b, "  if (("); gen(b, om->event[e], &c); cbprintf(b, ") != 0.0) m |= %uu;\n", 1u << e); }
  }
  CharBuffer_append(b, "  return m;\n}\n");

  /* store_row: refresh all rules, write the selected values (coalesced: consecutive sets are consecutive addresses) */
  CharBuffer_append(b, "DEV void store_row(double *__restrict__ row, size_t stride, double t, const double (&y)[A1(NEQ)], const PvRef pv)\n{\n  double a[A1(NASS)]; (void)t; (void)a; (void)pv; (void)y;\n");
  {
    int need_rules = 0, need_pv = 0;
    for (j = 0; j < nstore; j++) { if (store[j] >= om->neq && store[j] < om->neq + om->nass) need_rules = 1; if (store[j] >= om->neq) need_pv = 1; }
    store_needs_pv = need_rules || need_pv;
    if (need_rules) emit_refresh_all(b, om, &c, "  ");
    for (j = 0; j < nstore; j++) { cbprintf(b, "  row[%d * stride] = ", j); name_cuda(b, store[j], &c); CharBuffer_append(b, ";\n"); }
  }
  CharBuffer_append(b, "}\n");
  /* whether store_row reads anything but y (the output block then skips fetching the per-trajectory constants) */
  cbprintf(b, "#define STORE_NEEDS_PV %d\n", store_needs_pv);
  /* row 0: values after initial assignments with the model's own parameters, varied entries overwritten
     (setVariableValue at t == 0 patches results->value[idx][0] only, integratorInstance.c:1571-1572) */
  CharBuffer_append(b, "DEV void store_row0(double *__restrict__ row, size_t stride, const double (&y)[A1(NEQ)], const PvRef pv)\n{\n  (void)y; (void)pv;\n");
  for (j = 0; j < nstore; j++) {
    int idx = store[j], varied = 0, jj;
    for (jj = 0; jj < nvary; jj++) if (vary[jj] == idx) varied = 1;
    cbprintf(b, "  row[%d * stride] = ", j);
    if (varied && idx < om->neq) cbprintf(b, "y[%d]", idx);
    else if (varied && k.kind[idx] == K_THREAD) cbprintf(b, "pv[%d]", k.slot[idx]);
    else cbprintf(b, "c_base[%d]", idx);
    CharBuffer_append(b, ";\n");
  }
  CharBuffer_append(b, "}\n");

  /* steady_f: IntegratorInstance_checkSteadyState (src/integratorInstance.c:1442-1507) at an output time, after the row
     has been stored: mean and standard deviation of the right-hand sides, evaluated the way the reference's
     interpreter does on the refreshed values; non-zero when mean + std is below the threshold */
  cbprintf(b, "#define HALT_ON_STEADY %d\n", (opt->SteadyState == 1 && om->neq > 0) ? 1 : 0);
  if (opt->SteadyState == 1 && om->neq > 0) {
    CharBuffer_append(b, "#define SS_THRESHOLD "); cb_double17(b, opt->ssThreshold); CharBuffer_append(b, "\n");
    CharBuffer_append(b, "DEV int steady_f(double t, const double (&y)[A1(NEQ)], const PvRef pv)\n{\n  double a[A1(NASS)], dy[A1(NEQ)], mean = 0.0, var = 0.0; (void)t; (void)a; (void)pv;\n");
    k.local = all_local; c.interp = 1;
    emit_refresh_all(b, om, &c, "  ");
    for (i = 0; i < om->neq; i++) { cbprintf(b, "  dy[%d] = ", i); gen(b, om->ode[i], &c); CharBuffer_append(b, ";\n"); }
    CharBuffer_append(b, "  _Pragma(\"unroll\") for (int i = 0; i < NEQ; i++) mean += fabs(dy[i]);\n  mean = mean / NEQ;\n"
                         "  _Pragma(\"unroll\") for (int i = 0; i < NEQ; i++) { const double d = dy[i] - mean; var += d * d; }\n"
                         "  var = var / (NEQ - 1);\n  return (mean + sqrt(var)) < SS_THRESHOLD;\n}\n");
  }

  free(k.kind); free(k.slot); free(used_in_ode); free(all_local); free(ode_local);
  out = b->b; b->b = NULL; CharBuffer_free(b);
  return out;
}
