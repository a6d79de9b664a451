/*
 * ast.h -- abstract syntax trees for SBML math, own implementation.
 *
 * The reference (raim/SBML_odeSolver) uses libSBML's ASTNode_t everywhere
 * (src/processAST.c, src/evaluateAST.c, src/odeModel.c).  libSBML is not part
 * of this build, so this header provides the subset of the ASTNode C API the
 * hot path needs, with the same function names and the same node kinds
 * (libSBML ASTNodeType_t), plus the "index name" extension of
 * src/ASTIndexNameNode.c (an AST_NAME carrying its position in value[]).
 */
#ifndef ODES_B200_AST_H
#define ODES_B200_AST_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  AST_PLUS = '+', AST_MINUS = '-', AST_TIMES = '*', AST_DIVIDE = '/', AST_POWER = '^',
  AST_INTEGER = 256, AST_REAL, AST_REAL_E, AST_RATIONAL,
  AST_NAME, AST_NAME_TIME,
  AST_CONSTANT_E, AST_CONSTANT_FALSE, AST_CONSTANT_PI, AST_CONSTANT_TRUE,
  AST_LAMBDA,
  AST_FUNCTION,
  AST_FUNCTION_ABS, AST_FUNCTION_ARCCOS, AST_FUNCTION_ARCCOSH, AST_FUNCTION_ARCCOT,
  AST_FUNCTION_ARCCOTH, AST_FUNCTION_ARCCSC, AST_FUNCTION_ARCCSCH, AST_FUNCTION_ARCSEC,
  AST_FUNCTION_ARCSECH, AST_FUNCTION_ARCSIN, AST_FUNCTION_ARCSINH, AST_FUNCTION_ARCTAN,
  AST_FUNCTION_ARCTANH, AST_FUNCTION_CEILING, AST_FUNCTION_COS, AST_FUNCTION_COSH,
  AST_FUNCTION_COT, AST_FUNCTION_COTH, AST_FUNCTION_CSC, AST_FUNCTION_CSCH,
  AST_FUNCTION_DELAY, AST_FUNCTION_EXP, AST_FUNCTION_FACTORIAL, AST_FUNCTION_FLOOR,
  AST_FUNCTION_LN, AST_FUNCTION_LOG, AST_FUNCTION_PIECEWISE, AST_FUNCTION_POWER,
  AST_FUNCTION_ROOT, AST_FUNCTION_SEC, AST_FUNCTION_SECH, AST_FUNCTION_SIN,
  AST_FUNCTION_SINH, AST_FUNCTION_TAN, AST_FUNCTION_TANH,
  AST_LOGICAL_AND, AST_LOGICAL_NOT, AST_LOGICAL_OR, AST_LOGICAL_XOR,
  AST_RELATIONAL_EQ, AST_RELATIONAL_GEQ, AST_RELATIONAL_GT, AST_RELATIONAL_LEQ,
  AST_RELATIONAL_LT, AST_RELATIONAL_NEQ,
  AST_UNKNOWN
} ASTNodeType_t;

typedef struct ASTNode ASTNode_t;
struct ASTNode {
  ASTNodeType_t type;
  long   ival;          /* AST_INTEGER value, AST_RATIONAL numerator            */
  long   denom;         /* AST_RATIONAL denominator                             */
  double rval;          /* AST_REAL / AST_REAL_E value (mantissa*10^exponent)   */
  char  *name;          /* AST_NAME*, AST_FUNCTION                              */
  int    index;         /* >=0: position in value[] (ASTIndexNameNode.c), else -1 */
  int    isData;        /* "data" flag of an index name (observation data)      */
  unsigned int nchildren, cap;
  ASTNode_t **children;
};

typedef int (*ASTNodePredicate)(const ASTNode_t *);

/* simple pointer list, stands in for libSBML's List_t where the reference uses it */
typedef struct { void **items; unsigned int size, cap; } List_t;
List_t *List_create(void);
void List_add(List_t *, void *);
void *List_get(const List_t *, unsigned int);
void *List_remove(List_t *, unsigned int);
unsigned int List_size(const List_t *);
void List_free(List_t *);

ASTNode_t *ASTNode_create(void);
ASTNode_t *ASTNode_createWithType(ASTNodeType_t);
ASTNode_t *ASTNode_createIndexName(void);
void ASTNode_free(ASTNode_t *);
ASTNode_t *ASTNode_deepCopy(const ASTNode_t *);
void ASTNode_addChild(ASTNode_t *, ASTNode_t *);
void ASTNode_prependChild(ASTNode_t *, ASTNode_t *);
ASTNode_t *ASTNode_getChild(const ASTNode_t *, unsigned int);
ASTNode_t *ASTNode_getLeftChild(const ASTNode_t *);
ASTNode_t *ASTNode_getRightChild(const ASTNode_t *);
unsigned int ASTNode_getNumChildren(const ASTNode_t *);
void ASTNode_swapChildren(ASTNode_t *, ASTNode_t *);
void ASTNode_reduceToBinary(ASTNode_t *);

ASTNodeType_t ASTNode_getType(const ASTNode_t *);
void ASTNode_setType(ASTNode_t *, ASTNodeType_t);
void ASTNode_setCharacter(ASTNode_t *, char);
void ASTNode_setName(ASTNode_t *, const char *);
void ASTNode_setInteger(ASTNode_t *, long);
void ASTNode_setReal(ASTNode_t *, double);
void ASTNode_setRealWithExponent(ASTNode_t *, double mantissa, long exponent);
void ASTNode_setRational(ASTNode_t *, long num, long den);
const char *ASTNode_getName(const ASTNode_t *);
long ASTNode_getInteger(const ASTNode_t *);
double ASTNode_getReal(const ASTNode_t *);
int ASTNode_getPrecedence(const ASTNode_t *);

int ASTNode_isInteger(const ASTNode_t *);
int ASTNode_isReal(const ASTNode_t *);
int ASTNode_isNumber(const ASTNode_t *);
int ASTNode_isName(const ASTNode_t *);
int ASTNode_isOperator(const ASTNode_t *);
int ASTNode_isFunction(const ASTNode_t *);
int ASTNode_isLogical(const ASTNode_t *);
int ASTNode_isRelational(const ASTNode_t *);
int ASTNode_isConstant(const ASTNode_t *);
int ASTNode_isUMinus(const ASTNode_t *);
int ASTNode_isUnknown(const ASTNode_t *);
int ASTNode_isLog10(const ASTNode_t *);
int ASTNode_isSqrt(const ASTNode_t *);
int ASTNode_isLambda(const ASTNode_t *);

/* index-name extension, cf. reference src/ASTIndexNameNode.c */
int ASTNode_isIndexName(const ASTNode_t *);
int ASTNode_isSetIndex(const ASTNode_t *);
unsigned int ASTNode_getIndex(const ASTNode_t *);
void ASTNode_setIndex(ASTNode_t *, unsigned int);
int ASTNode_isSetData(const ASTNode_t *);
void ASTNode_setData(ASTNode_t *);

/* pre-order list of all nodes satisfying the predicate (libSBML order) */
List_t *ASTNode_getListOfNodes(const ASTNode_t *, ASTNodePredicate);

/* infix formulas, SBML Level 1 syntax (libSBML SBML_parseFormula / SBML_formulaToString) */
ASTNode_t *SBML_parseFormula(const char *formula);
char *SBML_formulaToString(const ASTNode_t *);

/* MathML <math> fragment reader (libSBML readMathMLFromString); NULL on error */
ASTNode_t *readMathMLFromString(const char *xml);
/* MathML writer; caller frees */
char *writeMathMLToString(const ASTNode_t *);

#ifdef __cplusplus
}
#endif
#endif
