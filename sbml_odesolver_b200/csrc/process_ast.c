/*
 * process_ast.c -- symbolic processing of formula trees (host side, ISO C).
 *
 * Restates, for the AST type of ast.h, the behaviour of the reference's
 *   src/processAST.c:321-1139   differentiateAST  (rules per node kind)
 *   src/processAST.c:1276-1360  indexAST
 *   src/processAST.c:1389-1749  simplifyAST
 *   src/processAST.c:1800-1873  index helpers, containsTime/Piecewise
 *   src/modelSimplify.c:66-378  AST_replace*
 *   src/evaluateAST.c:106-604   evaluateAST
 * The result trees must be structurally identical to the reference's because
 * the equation strings and the Jacobian sparsity are pinned by its unit tests
 * (unittest/test_processAST.c:248-367, unittest/test_odeModel.c:564-719).
 *
 * Derivative rules are written as infix templates over the operands
 * (A, B: copies of the first/second child, dA, dB: their derivatives,
 *  F: copy of the whole node) and instantiated by substitution; this keeps the
 * tree shapes explicit and auditable against the rule comments of the
 * reference.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sbmlsolver/processAST.h"
#include "sbmlsolver/cvodeData.h"
#include "sbmlsolver/solverError.h"

static char *sdup_(const char *s) { char *r = (char *)malloc(strlen(s) + 1); strcpy(r, s); return r; }
ASTNode_t *copyAST(const ASTNode_t *f) { return ASTNode_deepCopy(f); }

void AST_dump(const char *context, ASTNode_t *node)
{
  char *s = SBML_formulaToString(node);
  printf("%s %s\n", context, s);
  free(s);
}

/* ------------------------------------------------------------ predicates */
static int user_defined(const ASTNode_t *n) { return ASTNode_isName(n) || ASTNode_getType(n) == AST_FUNCTION; }
static int is_name(const ASTNode_t *n) { return ASTNode_isName(n); }
static int is_function(const ASTNode_t *n) { return ASTNode_isFunction(n); }

static int zero(const ASTNode_t *f)
{
  if (ASTNode_isReal(f)) return ASTNode_getReal(f) == 0.0;
  if (ASTNode_isInteger(f)) return ASTNode_getInteger(f) == 0;
  return 0;
}
static int one(const ASTNode_t *f)
{
  if (ASTNode_isReal(f)) return ASTNode_getReal(f) == 1.0;
  if (ASTNode_isInteger(f)) return ASTNode_getInteger(f) == 1;
  return 0;
}
/* detach and return the only child of a unary node, freeing the node */
static ASTNode_t *cut_root(ASTNode_t *old)
{
  ASTNode_t *c = copyAST(ASTNode_getChild(old, 0));
  ASTNode_free(old);
  return c;
}
static int depends_on(const ASTNode_t *f, const char *x)
{
  unsigned int i; int found = 0;
  List_t *l = ASTNode_getListOfNodes(f, user_defined);
  for (i = 0; i < List_size(l); i++) {
    const char *nm = ASTNode_getName((ASTNode_t *)List_get(l, i));
    if (nm && strcmp(nm, x) == 0) found = 1;
  }
  List_free(l);
  return found;
}

/* ------------------------------------------------------------- indexAST
   Copy of an expression in which every symbol that names a model value carries its position in value[] (reference
   src/processAST.c:1276-1360).  A symbol "<id>_data" stands for the observation data of <id>: it is indexed like <id> and
   flagged.  Whichever of the two readings occurs FIRST in names[] wins, as in the reference's single pass over names[]. */
static int name_position(char **names, int nvalues, const char *s, size_t len)
{
  int i;
  for (i = 0; i < nvalues; i++) if (strlen(names[i]) == len && strncmp(names[i], s, len) == 0) return i;
  return nvalues;
}
static ASTNode_t *indexed_symbol(const ASTNode_t *f, int nvalues, char **names)
{
  const char *id = ASTNode_getName(f);
  const size_t len = id ? strlen(id) : 0;
  const int whole = id ? name_position(names, nvalues, id, len) : nvalues;
  const int stem = (id && len > 5 && strstr(id, "_data")) ? name_position(names, nvalues, id, len - 5) : nvalues;
  ASTNode_t *n;
  if (whole == nvalues && stem == nvalues) { n = ASTNode_create(); ASTNode_setName(n, id); }
  else {
    n = ASTNode_createIndexName();
    if (whole < stem) { ASTNode_setName(n, id); ASTNode_setIndex(n, (unsigned int)whole); }
    else { ASTNode_setName(n, names[stem]); ASTNode_setIndex(n, (unsigned int)stem); ASTNode_setData(n); }
  }
  ASTNode_setType(n, ASTNode_getType(f));            /* AST_NAME_TIME stays what it is */
  return n;
}
ASTNode_t *indexAST(const ASTNode_t *f, int nvalues, char **names)
{
  ASTNode_t *n; unsigned int k;
  if (ASTNode_isName(f)) return indexed_symbol(f, nvalues, names);
  n = ASTNode_create();
  if (ASTNode_isInteger(f)) ASTNode_setInteger(n, ASTNode_getInteger(f));
  else if (ASTNode_isReal(f)) ASTNode_setReal(n, ASTNode_getReal(f));
  else {
    if (ASTNode_getType(f) == AST_FUNCTION) ASTNode_setName(n, ASTNode_getName(f));
    ASTNode_setType(n, ASTNode_getType(f));
    for (k = 0; k < ASTNode_getNumChildren(f); k++) ASTNode_addChild(n, indexAST(ASTNode_getChild(f, k), nvalues, names));
  }
  return n;
}

int ASTNode_getIndices(const ASTNode_t *node, List_t *indices)
{
  unsigned int i;
  if (ASTNode_isSetIndex(node)) {
    unsigned int *idx = (unsigned int *)malloc(sizeof *idx);
    *idx = ASTNode_getIndex(node);
    List_add(indices, idx);
  }
  for (i = 0; i < ASTNode_getNumChildren(node); i++) ASTNode_getIndices(ASTNode_getChild(node, i), indices);
  return 1;
}
int *ASTNode_getIndexArray(const ASTNode_t *node, int nvalues)
{
  int *result = (int *)calloc((size_t)(nvalues > 0 ? nvalues : 1), sizeof(int));
  List_t *indices;
  if (!node) return result;
  indices = List_create();
  ASTNode_getIndices(node, indices);
  while (List_size(indices)) {
    unsigned int *k = (unsigned int *)List_remove(indices, 0);
    if ((int)*k < nvalues) result[*k] = 1;
    free(k);
  }
  List_free(indices);
  return result;
}
int ASTNode_containsTime(const ASTNode_t *node)
{
  unsigned int i;
  if (ASTNode_getType(node) == AST_NAME_TIME) return 1;
  for (i = 0; i != ASTNode_getNumChildren(node); i++) if (ASTNode_containsTime(ASTNode_getChild(node, i))) return 1;
  return 0;
}
int ASTNode_containsPiecewise(const ASTNode_t *node)
{
  unsigned int i;
  if (ASTNode_getType(node) == AST_FUNCTION_PIECEWISE) return 1;
  for (i = 0; i != ASTNode_getNumChildren(node); i++) if (ASTNode_containsPiecewise(ASTNode_getChild(node, i))) return 1;
  return 0;
}

/* ----------------------------------------------------------- simplifyAST */
static ASTNode_t *unary_of(ASTNodeType_t t, ASTNode_t *c) { ASTNode_t *n = ASTNode_createWithType(t); ASTNode_addChild(n, c); return n; }
static ASTNode_t *binary_of(ASTNodeType_t t, ASTNode_t *a, ASTNode_t *b)
{ ASTNode_t *n = ASTNode_createWithType(t); ASTNode_addChild(n, a); ASTNode_addChild(n, b); return n; }

ASTNode_t *simplifyAST(const ASTNode_t *f)
{
  unsigned int i, childnum;
  ASTNode_t *simple, *left, *right, *helper;
  ASTNodeType_t type = ASTNode_getType(f);

  if (ASTNode_isInteger(f)) { simple = ASTNode_create(); ASTNode_setInteger(simple, ASTNode_getInteger(f)); return simple; }
  if (ASTNode_isReal(f)) { simple = ASTNode_create(); ASTNode_setReal(simple, ASTNode_getReal(f)); return simple; }
  if (ASTNode_isName(f)) {
    simple = ASTNode_create();
    ASTNode_setName(simple, ASTNode_getName(f));
    ASTNode_setType(simple, type);
    if (ASTNode_isSetIndex(f)) { ASTNode_setIndex(simple, ASTNode_getIndex(f)); if (ASTNode_isSetData(f)) ASTNode_setData(simple); }
    return simple;
  }
  if (ASTNode_isUMinus(f)) {
    left = simplifyAST(ASTNode_getLeftChild(f));
    if (zero(left)) return left;                       /* -0 -> 0   */
    if (ASTNode_isUMinus(left)) return cut_root(left); /* --x -> x  */
    return unary_of(AST_MINUS, left);
  }
  if (ASTNode_isOperator(f) || type == AST_FUNCTION_POWER) {
    childnum = ASTNode_getNumChildren(f);
    if (childnum == 0) {                               /* neutral element */
      simple = ASTNode_create();
      if (type == AST_PLUS) ASTNode_setInteger(simple, 0);
      else if (type == AST_TIMES) ASTNode_setInteger(simple, 1);
      return simple;
    }
    if (childnum == 1) return simplifyAST(ASTNode_getChild(f, 0));
    if (childnum > 2) {
      /* n-ary: first operand op (rest), re-associating to the right (reference :1467-1483) */
      simple = ASTNode_createWithType(type);
      ASTNode_addChild(simple, simplifyAST(ASTNode_getChild(f, 0)));
      helper = ASTNode_createWithType(type);
      for (i = 1; i < childnum; i++) ASTNode_addChild(helper, simplifyAST(ASTNode_getChild(f, i)));
      ASTNode_addChild(simple, simplifyAST(helper));
      ASTNode_free(helper);
      return simple;
    }
    left = simplifyAST(ASTNode_getLeftChild(f));
    right = simplifyAST(ASTNode_getRightChild(f));
    switch (type) {
    case AST_PLUS:
      if (zero(right)) { ASTNode_free(right); return left; }
      if (zero(left)) { ASTNode_free(left); return right; }
      if (ASTNode_isUMinus(left) && ASTNode_isUMinus(right))        /* -x + -y -> -(x+y) */
        return unary_of(AST_MINUS, binary_of(AST_PLUS, cut_root(left), cut_root(right)));
      if (ASTNode_isUMinus(left)) return binary_of(AST_MINUS, right, cut_root(left));   /* -x + y -> y - x */
      if (ASTNode_isUMinus(right)) return binary_of(AST_MINUS, left, cut_root(right));  /* x + -y -> x - y */
      break;
    case AST_MINUS:
      if (zero(right)) { ASTNode_free(right); return left; }
      if (zero(left)) { ASTNode_free(left); return unary_of(AST_MINUS, right); }         /* 0 - x -> -x */
      if (ASTNode_isUMinus(left) && ASTNode_isUMinus(right)) {                             /* -x - -y -> y - x */
        ASTNode_t *r = cut_root(right), *l = cut_root(left);
        return binary_of(AST_MINUS, r, l);
      }
      if (ASTNode_isUMinus(left)) return unary_of(AST_MINUS, binary_of(AST_PLUS, cut_root(left), right)); /* -x - y */
      if (ASTNode_isUMinus(right)) return binary_of(AST_PLUS, left, cut_root(right));     /* x - -y -> x + y */
      break;
    case AST_TIMES:
      if (zero(right)) { ASTNode_free(left); return right; }
      if (zero(left)) { ASTNode_free(right); return left; }
      if (one(right)) { ASTNode_free(right); return left; }
      if (one(left)) { ASTNode_free(left); return right; }
      if (ASTNode_isUMinus(left) && ASTNode_isUMinus(right)) {
        ASTNode_t *l = cut_root(left), *r = cut_root(right);
        return binary_of(AST_TIMES, l, r);
      }
      if (ASTNode_isUMinus(left)) return unary_of(AST_MINUS, binary_of(AST_TIMES, cut_root(left), right));
      if (ASTNode_isUMinus(right)) return unary_of(AST_MINUS, binary_of(AST_TIMES, left, cut_root(right)));
      break;
    case AST_DIVIDE:
      if (zero(left)) { ASTNode_free(right); return left; }
      if (one(right)) { ASTNode_free(right); return left; }
      if (ASTNode_isUMinus(left) && ASTNode_isUMinus(right)) {
        ASTNode_t *l = cut_root(left), *r = cut_root(right);
        return binary_of(AST_DIVIDE, l, r);
      }
      if (ASTNode_isUMinus(left)) return unary_of(AST_MINUS, binary_of(AST_DIVIDE, cut_root(left), right));
      if (ASTNode_isUMinus(right)) return unary_of(AST_MINUS, binary_of(AST_DIVIDE, left, cut_root(right)));
      break;
    case AST_POWER:
    case AST_FUNCTION_POWER:
      if (zero(right)) { ASTNode_free(left); ASTNode_free(right); simple = ASTNode_create(); ASTNode_setReal(simple, 1.0); return simple; }
      if (one(right)) { ASTNode_free(right); return left; }
      if (zero(left)) { ASTNode_free(left); ASTNode_free(right); simple = ASTNode_create(); ASTNode_setReal(simple, 0.0); return simple; }
      if (one(left)) { ASTNode_free(left); ASTNode_free(right); simple = ASTNode_create(); ASTNode_setReal(simple, 1.0); return simple; }
      break;
    default:
      SolverError_error(WARNING_ERROR_TYPE, SOLVER_ERROR_AST_UNKNOWN_FAILURE, "simplifyAST: unknown failure for operator type");
      break;
    }
    return binary_of(type, left, right);
  }
  /* constants, functions, logical, relational: copy with simplified children */
  simple = ASTNode_createWithType(type);
  if (type == AST_FUNCTION || type == AST_FUNCTION_DELAY) { simple->type = type; ASTNode_setName(simple, ASTNode_getName(f)); simple->type = type; }
  for (i = 0; i < ASTNode_getNumChildren(f); i++) ASTNode_addChild(simple, simplifyAST(ASTNode_getChild(f, i)));
  return simple;
}

/* ------------------------------------------------------ differentiateAST */
typedef struct { const char *key; const ASTNode_t *tree; } subst_t;

/* replaces placeholder names of a parsed template by copies of the given trees */
static ASTNode_t *instantiate(ASTNode_t *t, const subst_t *s, int ns)
{
  unsigned int i; int k;
  if (t->type == AST_NAME && t->name) {
    for (k = 0; k < ns; k++)
      if (strcmp(t->name, s[k].key) == 0) { ASTNode_t *c = copyAST(s[k].tree); ASTNode_free(t); return c; }
    if (strcmp(t->name, "UM1") == 0) {            /* explicit unary minus of integer 1 */
      ASTNode_t *o = ASTNode_create(); ASTNode_setInteger(o, 1);
      ASTNode_free(t);
      return unary_of(AST_MINUS, o);
    }
    if (strcmp(t->name, "R1") == 0) { ASTNode_setReal(t, 1.0); return t; }   /* real 1.0 */
  }
  for (i = 0; i < t->nchildren; i++) t->children[i] = instantiate(t->children[i], s, ns);
  return t;
}
static ASTNode_t *build(const char *tmpl, const subst_t *s, int ns)
{
  ASTNode_t *t = SBML_parseFormula(tmpl);
  return instantiate(t, s, ns);
}
static ASTNode_t *failed(errorCode_t code, const char *msg)
{
  ASTNode_t *n = ASTNode_create();
  SolverError_error(WARNING_ERROR_TYPE, code, "%s", msg);
  ASTNode_setName(n, "differentiation_failed");
  return n;
}

ASTNode_t *differentiateAST(ASTNode_t *f, char *x)
{
  unsigned int i, j, childnum;
  ASTNodeType_t type;
  ASTNode_t *fprime = NULL, *simple, *A, *B, *dA = NULL, *dB = NULL;
  subst_t s[6]; int ns = 0;
  const char *tmpl = NULL;

  if (!depends_on(f, x)) { fprime = ASTNode_create(); ASTNode_setReal(fprime, 0.0); return fprime; }

  type = ASTNode_getType(f);
  childnum = ASTNode_getNumChildren(f);
  A = ASTNode_getChild(f, 0);
  B = ASTNode_getChild(f, 1);

  if (ASTNode_isName(f)) {
    fprime = ASTNode_create();
    ASTNode_setReal(fprime, ASTNode_isSetData(f) ? 0.0 : 1.0);
  }
  else if (ASTNode_isOperator(f) || type == AST_FUNCTION_POWER) {
    switch (type) {
    case AST_PLUS: case AST_MINUS:
      fprime = ASTNode_createWithType(type);
      for (i = 0; i < childnum; i++) ASTNode_addChild(fprime, differentiateAST(ASTNode_getChild(f, i), x));
      break;
    case AST_TIMES:
      if (childnum != 2) {                       /* decompose n-ary product first */
        ASTNode_t *helper = simplifyAST(f);
        fprime = differentiateAST(helper, x);
        ASTNode_free(helper);
      } else tmpl = "dA * B + A * dB";
      break;
    case AST_DIVIDE: tmpl = "dA / B - A / B^2 * dB"; break;
    case AST_POWER: case AST_FUNCTION_POWER:
      if (!depends_on(B, x)) tmpl = "B * (A^(B - R1) * dA)";
      else if (!depends_on(A, x)) tmpl = "F * (ln(A) * dB)";
      else tmpl = "F * (B / A * dA + ln(A) * dB)";
      break;
    default:
      fprime = failed(SOLVER_ERROR_AST_DIFFERENTIATION_FAILED_OPERATOR, "differentiateAST: operator: impossible case");
    }
  }
  else if (ASTNode_isFunction(f) || type == AST_LAMBDA) {
    switch (type) {
    case AST_LAMBDA: fprime = failed(SOLVER_ERROR_AST_DIFFERENTIATION_FAILED_LAMBDA, "differentiateAST: lambda: not implemented"); break;
    case AST_FUNCTION: {
      const char *fname = ASTNode_getName(f);
      char *dfname = (char *)malloc(strlen(fname) + 40);
      if (strcmp(fname, x) == 0) {               /* derivative w.r.t. the function itself */
        fprime = ASTNode_createWithType(AST_FUNCTION);
        sprintf(dfname, "d_%s", fname);
        ASTNode_setName(fprime, dfname); fprime->type = AST_FUNCTION;
        for (j = 0; j < childnum; j++) ASTNode_addChild(fprime, copyAST(ASTNode_getChild(f, j)));
      } else {                                   /* chain rule over the arguments */
        ASTNode_t *sum = NULL, *tmp = NULL;
        for (i = 0; i < childnum; i++) {
          ASTNode_t *prod;
          if (!depends_on(ASTNode_getChild(f, i), x)) { prod = ASTNode_create(); ASTNode_setReal(prod, 0.0); }
          else {
            ASTNode_t *helper = ASTNode_createWithType(AST_FUNCTION);
            sprintf(dfname, "d%i_%s", (int)i, fname);
            ASTNode_setName(helper, dfname); helper->type = AST_FUNCTION;
            for (j = 0; j < childnum; j++) ASTNode_addChild(helper, copyAST(ASTNode_getChild(f, j)));
            prod = binary_of(AST_TIMES, helper, differentiateAST(ASTNode_getChild(f, i), x));
          }
          sum = (i == 0) ? prod : binary_of(AST_PLUS, tmp, prod);
          tmp = sum;
        }
        fprime = sum;
      }
      free(dfname);
      break;
    }
    case AST_FUNCTION_ABS: tmpl = "piecewise(-1, lt(A, 0), 0, eq(A, 0), 1, gt(A, 0)) * dA"; break;
    case AST_FUNCTION_ARCCOS: tmpl = "-dA / sqrt(1 - A^2)"; break;
    case AST_FUNCTION_ARCCOSH: tmpl = "dA / sqrt(A^2 - 1)"; break;
    case AST_FUNCTION_ARCCOT: tmpl = "-dA / (1 + A^2)"; break;
    case AST_FUNCTION_ARCCOTH: tmpl = "-dA / (UM1 + A^2)"; break;
    case AST_FUNCTION_ARCCSC: tmpl = "-dA / (abs(A) * sqrt(A^2 - 1))"; break;
    case AST_FUNCTION_ARCCSCH: tmpl = "-dA / (A^2 * sqrt(1 + 1 / A^2))"; break;
    case AST_FUNCTION_ARCSEC: tmpl = "-dA / (A^2 * sqrt(1 - 1 / A^2))"; break;
    case AST_FUNCTION_ARCSECH: tmpl = "dA * (sqrt((1 - A) / (1 + A)) / (A * (A - 1)))"; break;
    case AST_FUNCTION_ARCSIN: tmpl = "dA / sqrt(1 - A^2)"; break;
    case AST_FUNCTION_ARCSINH: tmpl = "dA / sqrt(1 + A^2)"; break;
    case AST_FUNCTION_ARCTAN: tmpl = "dA / (1 + A^2)"; break;
    case AST_FUNCTION_ARCTANH: tmpl = "dA / (1 - A^2)"; break;
    case AST_FUNCTION_CEILING: case AST_FUNCTION_FLOOR:
      fprime = unary_of(type, differentiateAST(A, x));      /* sic: same function of a' */
      break;
    case AST_FUNCTION_COS: tmpl = "-sin(A) * dA"; break;
    case AST_FUNCTION_COSH: tmpl = "sinh(A) * dA"; break;
    case AST_FUNCTION_COT: tmpl = "-dA / sin(A)^2"; break;
    case AST_FUNCTION_COTH: tmpl = "-dA / sinh(A)^2"; break;
    case AST_FUNCTION_CSC: tmpl = "-dA * (csc(A) * cot(A))"; break;
    case AST_FUNCTION_CSCH: tmpl = "-dA * (csch(A) * coth(A))"; break;
    case AST_FUNCTION_DELAY: fprime = failed(SOLVER_ERROR_AST_DIFFERENTIATION_FAILED_DELAY, "differentiateAST: delay: not implemented"); break;
    case AST_FUNCTION_EXP: tmpl = "F * dA"; break;
    case AST_FUNCTION_FACTORIAL: fprime = failed(SOLVER_ERROR_AST_DIFFERENTIATION_FAILED_FACTORIAL, "differentiateAST: factorial: impossible case"); break;
    case AST_FUNCTION_LN: tmpl = "R1 / A * dA"; break;
    case AST_FUNCTION_LOG: {                 /* log(b, x) = 1/ln(b) * ln(x): (1/ln(A)) * d ln(B) */
      ASTNode_t *helper = unary_of(AST_FUNCTION_LN, copyAST(B));
      ASTNode_t *dh = differentiateAST(helper, x);
      s[0].key = "A"; s[0].tree = A; s[1].key = "X"; s[1].tree = dh;
      fprime = build("R1 / ln(A) * X", s, 2);
      ASTNode_free(helper); ASTNode_free(dh);
      break;
    }
    case AST_FUNCTION_PIECEWISE: fprime = failed(SOLVER_ERROR_AST_DIFFERENTIATION_FAILED_PIECEWISE, "differentiateAST: piecewise: not implemented"); break;
    case AST_FUNCTION_ROOT: {                /* treated as child0 ^ (1 / child1), as the reference does */
      ASTNode_t *helper;
      s[0].key = "A"; s[0].tree = A; s[1].key = "B"; s[1].tree = B;
      helper = build("pow(A, R1 / B)", s, 2);
      fprime = differentiateAST(helper, x);
      ASTNode_free(helper);
      break;
    }
    case AST_FUNCTION_SEC: tmpl = "dA * (sec(A) * tan(A))"; break;
    case AST_FUNCTION_SECH: tmpl = "-dA * (sech(A) * tanh(A))"; break;
    case AST_FUNCTION_SIN: tmpl = "cos(A) * dA"; break;
    case AST_FUNCTION_SINH: tmpl = "cosh(A) * dA"; break;
    case AST_FUNCTION_TAN: tmpl = "dA / cos(A)^2"; break;
    case AST_FUNCTION_TANH: tmpl = "dA / cosh(A)^2"; break;
    default: fprime = failed(SOLVER_ERROR_AST_UNKNOWN_FAILURE, "differentiateAST: unknown failure for function type");
    }
  }
  else if (ASTNode_isLogical(f) || ASTNode_isRelational(f))
    fprime = failed(SOLVER_ERROR_AST_DIFFERENTIATION_FAILED_LOGICAL_OR_RELATIONAL, "differentiateAST: logical and relational not possible");
  else
    fprime = failed(SOLVER_ERROR_AST_UNKNOWN_NODE_TYPE, "differentiateAST: unknown ASTNode type");

  if (tmpl) {
    if (A) { dA = differentiateAST(A, x); s[ns].key = "A"; s[ns++].tree = A; s[ns].key = "dA"; s[ns++].tree = dA; }
    if (B && strstr(tmpl, "B")) { dB = differentiateAST(B, x); s[ns].key = "B"; s[ns++].tree = B; s[ns].key = "dB"; s[ns++].tree = dB; }
    s[ns].key = "F"; s[ns++].tree = f;
    fprime = build(tmpl, s, ns);
    ASTNode_free(dA); ASTNode_free(dB);
  }
  simple = simplifyAST(fprime);
  ASTNode_free(fprime);
  return simple;
}

/* ---------------------------------------------------------- replacements */
void AST_replaceNameByName(ASTNode_t *math, const char *name, const char *newname)
{
  unsigned int i; List_t *names = ASTNode_getListOfNodes(math, is_name);
  for (i = 0; i < List_size(names); i++) {
    ASTNode_t *n = (ASTNode_t *)List_get(names, i);
    if (ASTNode_getName(n) && strcmp(ASTNode_getName(n), name) == 0) ASTNode_setName(n, newname);
  }
  List_free(names);
}
void AST_replaceNameByValue(ASTNode_t *math, const char *name, double x)
{
  unsigned int i; List_t *names = ASTNode_getListOfNodes(math, is_name);
  for (i = 0; i < List_size(names); i++) {
    ASTNode_t *n = (ASTNode_t *)List_get(names, i);
    if (ASTNode_getName(n) && strcmp(ASTNode_getName(n), name) == 0) ASTNode_setReal(n, x);
  }
  List_free(names);
}
void AST_replaceNameByParameters(ASTNode_t *math, ListOf_t *lp)
{
  unsigned int i;
  for (i = 0; i < ListOf_size(lp); i++) {
    Parameter_t *p = (Parameter_t *)ListOf_get(lp, i);
    if (!p->constant) continue;
    AST_replaceNameByValue(math, p->id, p->value);
  }
}
void AST_replaceNameByFormula(ASTNode_t *math, const char *name, const ASTNode_t *formula)
{
  unsigned int i, j;
  List_t *names = ASTNode_getListOfNodes(math, is_name);
  for (i = 0; i < List_size(names); i++) {
    ASTNode_t *old = (ASTNode_t *)List_get(names, i);
    if (!ASTNode_getName(old) || strcmp(ASTNode_getName(old), name) != 0) continue;
    if (ASTNode_isName(formula)) { ASTNode_setName(old, ASTNode_getName(formula)); old->type = formula->type; old->index = formula->index; old->isData = formula->isData; }
    else if (ASTNode_isInteger(formula)) ASTNode_setInteger(old, ASTNode_getInteger(formula));
    else if (ASTNode_isReal(formula)) ASTNode_setReal(old, ASTNode_getReal(formula));
    else {
      ASTNode_setType(old, ASTNode_getType(formula));
      if (ASTNode_getType(formula) == AST_FUNCTION) { ASTNode_setName(old, ASTNode_getName(formula)); old->type = AST_FUNCTION; }
      for (j = 0; j < ASTNode_getNumChildren(formula); j++) ASTNode_addChild(old, copyAST(ASTNode_getChild(formula, j)));
    }
  }
  List_free(names);
}
void AST_replaceFunctionDefinition(ASTNode_t *math, const char *name, const ASTNode_t *function)
{
  unsigned int i, j, n;
  List_t *names = ASTNode_getListOfNodes(math, is_function);
  n = List_size(names);
  for (i = 0; i < n; i++) {
    ASTNode_t *old = (ASTNode_t *)List_get(names, n - i - 1);   /* innermost calls first */
    const char *old_name = ASTNode_getName(old);
    ASTNode_t *fresh;
    if (!old_name || strcmp(old_name, name) != 0 || ASTNode_getType(old) != AST_FUNCTION) continue;
    fresh = copyAST(ASTNode_getChild(function, ASTNode_getNumChildren(function) - 1));   /* lambda body */
    for (j = 0; j + 1 < ASTNode_getNumChildren(function); j++)
      AST_replaceNameByFormula(fresh, ASTNode_getName(ASTNode_getChild(function, j)), ASTNode_getChild(old, j));
    if (ASTNode_isName(fresh)) { char *nm = sdup_(ASTNode_getName(fresh)); ASTNode_setType(old, AST_NAME); ASTNode_setName(old, nm); old->type = fresh->type; free(nm);
      { unsigned int k; for (k = 0; k < old->nchildren; k++) ASTNode_free(old->children[k]); old->nchildren = 0; } }
    else if (ASTNode_isInteger(fresh)) ASTNode_setInteger(old, ASTNode_getInteger(fresh));
    else if (ASTNode_isReal(fresh)) ASTNode_setReal(old, ASTNode_getReal(fresh));
    else {
      ASTNodeType_t t = ASTNode_getType(fresh);
      if (t == AST_FUNCTION) { char *nm = sdup_(ASTNode_getName(fresh)); ASTNode_setName(old, nm); free(nm); old->type = AST_FUNCTION; }
      else ASTNode_setType(old, t);
      ASTNode_swapChildren(old, fresh);
    }
    ASTNode_free(fresh);
  }
  List_free(names);
}

void AST_replaceConstants(Model_t *m, ASTNode_t *math)
{
  unsigned int i, j, nrules = Model_getNumRules(m);
  for (i = 0; i < nrules; i++) {
    Rule_t *rl = Model_getRule(m, nrules - i - 1);
    if (rl->type == SBML_ASSIGNMENT_RULE && rl->math && rl->variable) AST_replaceNameByFormula(math, rl->variable, rl->math);
  }
  for (i = 0; i < Model_getNumFunctionDefinitions(m); i++) {
    FunctionDefinition_t *f = Model_getFunctionDefinition(m, i);
    AST_replaceFunctionDefinition(math, f->id, f->math);
  }
  for (i = 0; i < Model_getNumParameters(m); i++) {
    Parameter_t *p = Model_getParameter(m, i);
    if (p->constant) AST_replaceNameByValue(math, p->id, p->value);
  }
  for (i = 0; i < Model_getNumCompartments(m); i++) {
    Compartment_t *c = Model_getCompartment(m, i);
    if (c->constant) AST_replaceNameByValue(math, c->id, c->size);
  }
  for (i = 0; i < Model_getNumSpecies(m); i++) {
    Species_t *s = Model_getSpecies(m, i);
    Compartment_t *c = Model_getCompartmentById(m, s->compartment);
    int conc = !s->hasOnlySubstanceUnits && c && c->spatialDimensions != 0;
    double v = conc ? (s->isSetInitialConcentration ? s->initialConcentration : s->initialAmount / c->size) : s->initialAmount;
    int found = 0;
    if (s->constant) AST_replaceNameByValue(math, s->id, v);
    else if (s->boundaryCondition) {
      for (j = 0; j < Model_getNumRules(m); j++) {
        Rule_t *rl = Model_getRule(m, j);
        if ((rl->type == SBML_RATE_RULE || rl->type == SBML_ASSIGNMENT_RULE) && rl->math && rl->variable && strcmp(rl->variable, s->id) == 0) ++found;
      }
      if (!found) AST_replaceNameByValue(math, s->id, v);
    }
  }
}

/* ----------------------------------------------------------- evaluateAST */
static double (*UsrDefFunc)(char *, int, double *) = NULL;
void setUserDefinedFunction(double (*udf)(char *, int, double *)) { UsrDefFunc = udf; }

/* inverse hyperbolics spelled out as the reference does (src/evaluateAST.c:17-30) */
static double a_cosh(double x) { return log(x + (sqrt(x - 1) * sqrt(x + 1))); }
static double a_sinh(double x) { return log(x + sqrt((x * x) + 1)); }
static double a_tanh(double x) { return (log(1 + x) - log(1 - x)) / 2; }

#define CH(n, i) ASTNode_getChild(n, i)
#define EV(i) evaluateAST(CH(n, i), data)

double evaluateAST(ASTNode_t *n, cvodeData_t *data)
{
  int i, j, childnum, found, truecount;
  double v1, v2, v3, result = 0;
  ASTNodeType_t type;

  if (n == NULL) { SolverError_error(FATAL_ERROR_TYPE, SOLVER_ERROR_AST_UNKNOWN_NODE_TYPE, "evaluateAST: empty Abstract Syntax Tree (AST)."); return 0; }
  childnum = (int)ASTNode_getNumChildren(n);
  type = ASTNode_getType(n);
  switch (type) {
  case AST_INTEGER: result = (double)ASTNode_getInteger(n); break;
  case AST_REAL: case AST_REAL_E: case AST_RATIONAL: result = ASTNode_getReal(n); break;
  case AST_NAME:
    found = 0;
    if (ASTNode_isSetIndex(n) && data && (int)ASTNode_getIndex(n) < data->nvalues) { result = data->value[ASTNode_getIndex(n)]; found++; }
    if (!found && data && data->model)
      for (j = 0; j < data->nvalues; j++)
        if (strcmp(ASTNode_getName(n), data->model->names[j]) == 0) { result = data->value[j]; found++; }
    if (!found) {
      SolverError_error(FATAL_ERROR_TYPE, SOLVER_ERROR_AST_EVALUATION_FAILED_MISSING_VALUE,
                        "No value found for AST_NAME %s . Defaults to Zero to avoid program crash", ASTNode_getName(n));
      result = 0;
    }
    break;
  case AST_FUNCTION_DELAY:
    SolverError_error(FATAL_ERROR_TYPE, SOLVER_ERROR_AST_EVALUATION_FAILED_DELAY,
                      "Solving ODEs with Delay is not implemented. Defaults to 0 to avoid program crash");
    result = 0.0; break;
  case AST_NAME_TIME: result = data ? (double)data->currenttime : 0.0; break;
  case AST_CONSTANT_E: result = exp(1.); break;
  case AST_CONSTANT_FALSE: result = 0.0; break;
  case AST_CONSTANT_PI: result = 4. * atan(1.); break;
  case AST_CONSTANT_TRUE: result = 1.0; break;
  case AST_PLUS: result = 0.0; for (i = 0; i < childnum; i++) result += EV(i); break;
  case AST_MINUS: result = (childnum < 2) ? -(EV(0)) : EV(0) - EV(1); break;
  case AST_TIMES:
    result = 1.0;
    for (i = 0; i < childnum; i++) { result *= EV(i); if (result == 0.0) break; }
    break;
  case AST_DIVIDE: { double a = EV(0); double b = EV(1); result = a / b; } break;
  case AST_POWER: case AST_FUNCTION_POWER: { double a = EV(0); double b = EV(1); result = pow(a, b); } break;
  case AST_LAMBDA:
    SolverError_error(FATAL_ERROR_TYPE, SOLVER_ERROR_AST_EVALUATION_FAILED_LAMBDA,
                      "Lambda can only be used in SBML function definitions. Defaults to 0 to avoid program crash");
    result = 0.0; break;
  case AST_FUNCTION:
    if (UsrDefFunc == NULL) {
      SolverError_error(FATAL_ERROR_TYPE, SOLVER_ERROR_AST_EVALUATION_FAILED_FUNCTION,
                        "The function %s() has not been defined in the SBML input model or as an externally supplied function. "
                        "Defaults to 0 to avoid program crash", ASTNode_getName(n));
      result = 0.0;
    } else {
      double *vals = (double *)calloc((size_t)childnum + 1, sizeof(double));
      for (i = 0; i < childnum; i++) vals[i] = EV(i);
      result = UsrDefFunc((char *)ASTNode_getName(n), childnum, vals);
      free(vals);
    }
    break;
  case AST_FUNCTION_ABS: result = fabs(EV(0)); break;
  case AST_FUNCTION_ARCCOS: result = acos(EV(0)); break;
  case AST_FUNCTION_ARCCOSH: result = a_cosh(EV(0)); break;
  case AST_FUNCTION_ARCCOT: result = atan(1. / EV(0)); break;
  case AST_FUNCTION_ARCCOTH: v1 = EV(0); result = log((v1 + 1) / (v1 - 1)) / 2; break;
  case AST_FUNCTION_ARCCSC: v1 = EV(0); result = atan(1 / sqrt((v1 - 1) * (v1 + 1))); break;
  case AST_FUNCTION_ARCCSCH: v1 = EV(0); result = log(1 / v1 + sqrt(1 / (v1 * v1) + 1)); break;
  case AST_FUNCTION_ARCSEC: result = acos(1 / EV(0)); break;
  case AST_FUNCTION_ARCSECH: result = a_cosh(1. / EV(0)); break;
  case AST_FUNCTION_ARCSIN: result = asin(EV(0)); break;
  case AST_FUNCTION_ARCSINH: result = a_sinh(EV(0)); break;
  case AST_FUNCTION_ARCTAN: result = atan(EV(0)); break;
  case AST_FUNCTION_ARCTANH: result = a_tanh(EV(0)); break;
  case AST_FUNCTION_CEILING: result = ceil(EV(0)); break;
  case AST_FUNCTION_COS: result = cos(EV(0)); break;
  case AST_FUNCTION_COSH: result = cosh(EV(0)); break;
  case AST_FUNCTION_COT: result = 1. / tan(EV(0)); break;
  case AST_FUNCTION_COTH: result = 1 / tanh(EV(0)); break;
  case AST_FUNCTION_CSC: result = 1. / sin(EV(0)); break;
  case AST_FUNCTION_CSCH: result = 1. / sinh(EV(0)); break;
  case AST_FUNCTION_EXP: result = exp(EV(0)); break;
  case AST_FUNCTION_FACTORIAL:
    v1 = EV(0); j = (int)floor(v1);
    if (v1 != j) SolverError_error(FATAL_ERROR_TYPE, SOLVER_ERROR_AST_EVALUATION_FAILED_FLOAT_FACTORIAL,
                                   "The factorial is only implemented for integer values. The floor value of the passed float is used for calculation!");
    for (result = 1; j > 1; --j) result *= j;
    break;
  case AST_FUNCTION_FLOOR: result = floor(EV(0)); break;
  case AST_FUNCTION_LN: result = log(EV(0)); break;
  case AST_FUNCTION_LOG: { double num = log10(EV(1)); result = num / log10(EV(0)); } break;
  case AST_FUNCTION_PIECEWISE:
    found = 0;
    for (i = 0; i < childnum - 1; i = i + 2)
      if (EV(i + 1)) { found++; result = EV(i); }
    if (i == childnum - 1 && found == 0) { found++; result = EV(i); }
    if (found == 0) SolverError_error(FATAL_ERROR_TYPE, SOLVER_ERROR_AST_EVALUATION_FAILED_PIECEWISE, "Piecewise function failed; no true piece");
    if (found > 1) SolverError_error(FATAL_ERROR_TYPE, SOLVER_ERROR_AST_EVALUATION_FAILED_PIECEWISE, "Piecewise function failed; several true pieces");
    break;
  case AST_FUNCTION_ROOT:
    v1 = EV(1); v2 = EV(0); v3 = floor(v2);
    if (v2 == v3 && (v1 < 0) && ((int)v2 % 2 != 0)) result = -pow(fabs(v1), 1. / v2);
    else result = pow(v1, 1. / v2);
    break;
  case AST_FUNCTION_SEC: result = 1. / cos(EV(0)); break;
  case AST_FUNCTION_SECH: result = 1. / cosh(EV(0)); break;
  case AST_FUNCTION_SIN: result = sin(EV(0)); break;
  case AST_FUNCTION_SINH: result = sinh(EV(0)); break;
  case AST_FUNCTION_TAN: result = tan(EV(0)); break;
  case AST_FUNCTION_TANH: result = tanh(EV(0)); break;
  case AST_LOGICAL_AND:
    truecount = 0; for (i = 0; i < childnum; i++) truecount += (int)EV(i);
    result = (truecount == childnum) ? 1.0 : 0.0; break;
  case AST_LOGICAL_NOT: result = (double)(!(EV(0))); break;
  case AST_LOGICAL_OR:
    truecount = 0; for (i = 0; i < childnum; i++) truecount += (int)EV(i);
    result = (truecount > 0) ? 1.0 : 0.0; break;
  case AST_LOGICAL_XOR:
    truecount = 0; for (i = 0; i < childnum; i++) truecount += (int)EV(i);
    result = (truecount % 2 != 0) ? 1.0 : 0.0; break;
  case AST_RELATIONAL_EQ:
    result = 1.0; if (childnum <= 1) break;
    v1 = EV(0);
    for (i = 1; i < childnum; i++) if (v1 != EV(i)) { result = 0.0; break; }
    break;
  case AST_RELATIONAL_GEQ: case AST_RELATIONAL_GT: case AST_RELATIONAL_LEQ: case AST_RELATIONAL_LT:
    result = 1.0; if (childnum <= 1) break;
    v2 = EV(0);
    for (i = 1; i < childnum; i++) {
      int bad;
      v1 = v2; v2 = EV(i);
      bad = (type == AST_RELATIONAL_GEQ) ? (v1 < v2) : (type == AST_RELATIONAL_GT) ? (v1 <= v2)
          : (type == AST_RELATIONAL_LEQ) ? (v1 > v2) : (v1 >= v2);
      if (bad) { result = 0.0; break; }
    }
    break;
  case AST_RELATIONAL_NEQ: { double a = EV(0); double b = EV(1); result = (a != b); } break;
  default:
    SolverError_error(FATAL_ERROR_TYPE, SOLVER_ERROR_AST_UNKNOWN_NODE_TYPE, "evaluateAST: unknown ASTNode type: %d", (int)type);
    result = 0; break;
  }
  return result;
}
