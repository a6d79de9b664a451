/*
 * processAST.h -- symbolic operations on ASTs: copy, index, differentiate,
 * simplify, evaluate, name/function replacement.
 * Mirrors the reference interfaces src/sbmlsolver/processAST.h:56-71,
 * src/sbmlsolver/evaluateAST.h, src/sbmlsolver/modelSimplify.h:44-58.
 */
#ifndef ODES_B200_PROCESSAST_H
#define ODES_B200_PROCESSAST_H
#include "sbmlsolver/ast.h"
#include "sbmlsolver/sbmlModel.h"
#ifdef __cplusplus
extern "C" {
#endif
struct cvodeData;

ASTNode_t *copyAST(const ASTNode_t *f);
ASTNode_t *indexAST(const ASTNode_t *f, int nvalues, char **names);
ASTNode_t *simplifyAST(const ASTNode_t *f);
/* reference processAST.h:62: determinant of an N x N matrix of expressions (cofactor expansion, simplified); caller frees */
ASTNode_t *determinantNAST(ASTNode_t ***A, int N);
ASTNode_t *differentiateAST(ASTNode_t *f, char *x);
void AST_dump(const char *context, ASTNode_t *node);
int ASTNode_getIndices(const ASTNode_t *node, List_t *indices);
int *ASTNode_getIndexArray(const ASTNode_t *node, int nvalues);
int ASTNode_containsTime(const ASTNode_t *node);
int ASTNode_containsPiecewise(const ASTNode_t *node);

/* reference: src/evaluateAST.c:106 -- interpreter over data->value[] */
double evaluateAST(ASTNode_t *n, struct cvodeData *data);
void setUserDefinedFunction(double (*udf)(char *, int, double *));

/* reference: src/modelSimplify.c:66-378 */
void AST_replaceNameByName(ASTNode_t *math, const char *name, const char *newname);
void AST_replaceNameByValue(ASTNode_t *math, const char *name, double x);
void AST_replaceNameByParameters(ASTNode_t *math, ListOf_t *lp);
void AST_replaceNameByFormula(ASTNode_t *math, const char *name, const ASTNode_t *formula);
void AST_replaceFunctionDefinition(ASTNode_t *math, const char *name, const ASTNode_t *function);
void AST_replaceConstants(Model_t *m, ASTNode_t *math);

#ifdef __cplusplus
}
#endif
#endif
