/*
 * odeModel.h -- odeModel_t: the indexed ODE system derived from an SBML
 * model.  Struct fields and functions follow the reference's
 * src/sbmlsolver/odeModel.h:63-364, src/sbmlsolver/variableIndex.h:61-66,
 * src/sbmlsolver/odeConstruct.h:48-51 and src/sbmlsolver/compiler.h:63-107.
 * What differs: "compiled code" is a CUDA module for sm_100a (NVRTC) instead
 * of a dlopen'ed shared object, and the generated functions are device
 * functions inlined into the batched BDF kernel.
 */
#ifndef ODES_B200_ODEMODEL_H
#define ODES_B200_ODEMODEL_H
#include "sbmlsolver/ast.h"
#include "sbmlsolver/sbmlModel.h"
#include "sbmlsolver/integratorSettings.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct odeModel odeModel_t;
typedef struct nonzeroElem nonzeroElem_t;
typedef struct odeSense odeSense_t;
typedef struct variableIndex variableIndex_t;
typedef struct compiled_code compiled_code_t;

typedef enum variableType { ODE_VARIABLE, ASSIGNMENT_VARIABLE, ALGEBRAIC_VARIABLE, CONSTANT } variableType_t;
struct variableIndex { variableType_t type; int type_index; int index; };

struct nonzeroElem { int i, j; ASTNode_t *ij; };

/* reference compiler.h:63-80: handle of generated+compiled code.  Here: a cubin. */
struct compiled_code {
  void *cubin; unsigned long cubinSize;   /* sm_100a machine code image            */
  void *module;                           /* CUmodule once loaded on a device     */
  char *log;                              /* NVRTC compile log                    */
  char *source;                           /* generated CUDA source (kept for ncu) */
};

struct odeModel {
  SBMLDocument_t *d; Model_t *m; Model_t *simple; double *values;
  char **names; int **dependencyMatrix; int hasCycle;
  int nconst;
  int nass; ASTNode_t **assignment;
  nonzeroElem_t **assignmentOrder;
  int nassbeforeodes; nonzeroElem_t **assignmentsBeforeODEs;
  int neq; ASTNode_t **ode;
  ASTNode_t ***jacob; nonzeroElem_t **jacobSparse; int sparsesize; int jacobian; int jacobianFailed;
  int npiecewise;
  int *indexInit; int ninitAss; int *initIndex; ASTNode_t **initAssignment;
  int nevents; ASTNode_t **event; int *neventAss; int **eventIndex; ASTNode_t ***eventAssignment;
  nonzeroElem_t **initAssignmentOrder;
  int nassbeforeevents; nonzeroElem_t **assignmentsBeforeEvents;
  int nalg; ASTNode_t **algebraic;
  /* compiled-code cache: one entry per (varied-index set, storage plan), see batchIntegrator.h */
  void *kernelCache;
  /* kept for source compatibility with callers that test these flags */
  int discrete_observation_data; int compute_vector_v;
};

/* reference src/sbmlsolver/odeModel.h:214-247: the forward-sensitivity structures derived from an odeModel_t.
   sens[i][l] = d(ode_i)/d(parameter l) with the assignment rules inlined, differentiated, simplified and indexed
   (src/odeModel.c:1963-2100); variables (initial-condition sensitivities) have index_sensP == -1 and no column. */
struct odeSense {
  odeModel_t *om;
  int neq, nsens;
  int *index_sens;          /* [nsens] index into names[] / value[] */
  int nsensP; int *index_sensP;   /* [nsens] column of sens[][] or -1 */
  ASTNode_t ***sens;        /* [neq][nsensP] */
  int sensitivity;          /* 1: matrix constructed */
  int **sensLogic;          /* [neq][nsensP] structurally non-zero */
  nonzeroElem_t **sensSparse; int sparsesize;
};

/* construction */
odeModel_t *ODEModel_createFromFile(const char *);
odeModel_t *ODEModel_createFromSBML2(SBMLDocument_t *);
odeModel_t *ODEModel_create(Model_t *);
odeModel_t *ODEModel_createFromODEs(ASTNode_t **, int neq, int nass, int nconst, char **, double *, Model_t *);
void ODEModel_free(odeModel_t *);
Model_t *Model_reduceToOdes(Model_t *m);
ASTNode_t *Species_odeFromReactions(Species_t *s, Model_t *m);
double Model_getValueById(Model_t *m, const char *id);
int Model_setValue(Model_t *m, const char *id, const char *rid, double value);
SBMLDocument_t *parseModel(const char *file, int printMessage, int validate);

/* variables */
const Model_t *ODEModel_getModel(odeModel_t *);
const Model_t *ODEModel_getOdeSBML(odeModel_t *);
int ODEModel_getNumValues(const odeModel_t *);
int ODEModel_getNeq(const odeModel_t *);
int ODEModel_getNalg(const odeModel_t *);
int ODEModel_getNumAssignments(const odeModel_t *);
int ODEModel_getNumConstants(const odeModel_t *);
int ODEModel_hasVariable(const odeModel_t *, const char *);
variableIndex_t *ODEModel_getVariableIndexByNum(const odeModel_t *, int);
variableIndex_t *ODEModel_getOdeVariableIndex(const odeModel_t *, int);
variableIndex_t *ODEModel_getAssignedVariableIndex(const odeModel_t *, int);
variableIndex_t *ODEModel_getConstantIndex(const odeModel_t *, int);
variableIndex_t *ODEModel_getVariableIndex(odeModel_t *, const char *);
int ODEModel_getVariableIndexFields(const odeModel_t *, const char *SBML_ID);
void ODEModel_dumpNames(odeModel_t *);
const char *ODEModel_getVariableName(const odeModel_t *, const variableIndex_t *);
const char *VariableIndex_getName(const variableIndex_t *, const odeModel_t *);
int VariableIndex_getIndex(const variableIndex_t *);
void VariableIndex_free(variableIndex_t *);

/* rule ordering */
int ODEModel_hasCycle(const odeModel_t *);
int ODEModel_getNumAssignmentsBeforeODEs(const odeModel_t *);
int ODEModel_getNumAssignmentsBeforeEvents(const odeModel_t *);
int ODEModel_getNumJacobiElements(const odeModel_t *);
const nonzeroElem_t *ODEModel_getAssignmentOrder(odeModel_t *, int);
const nonzeroElem_t *ODEModel_getAssignmentBeforeODEs(odeModel_t *, int);
const nonzeroElem_t *ODEModel_getAssignmentBeforeEvents(odeModel_t *, int);
const nonzeroElem_t *ODEModel_getJacobiElement(const odeModel_t *, int);
const ASTNode_t *NonzeroElement_getEquation(nonzeroElem_t *);
const char *NonzeroElement_getVariableName(nonzeroElem_t *, odeModel_t *);
variableIndex_t *NonzeroElement_getVariableIndex(nonzeroElem_t *, odeModel_t *);
variableIndex_t *NonzeroElement_getVariable2Index(nonzeroElem_t *, odeModel_t *);
const Model_t *ODEModel_getInputSBML(odeModel_t *);
int ODEModel_getNumPiecewise(odeModel_t *);
int ODEModel_getNumEvents(odeModel_t *);
const ASTNode_t *ODEModel_getEventTrigger(odeModel_t *, int);
const ASTNode_t *ODEModel_getEventAssignment(odeModel_t *, int, int);
ASTNode_t *ASTNode_indexAST(const ASTNode_t *f, odeModel_t *om);
const char *NonzeroElement_getVariable2Name(nonzeroElem_t *, odeModel_t *);
List_t *topoSort(int **matrix, int n, int *changed, int *required);

/* equations */
const ASTNode_t *ODEModel_getOde(const odeModel_t *, const variableIndex_t *);
const ASTNode_t *ODEModel_getAssignment(const odeModel_t *, const variableIndex_t *);

/* Jacobian */
int ODEModel_constructJacobian(odeModel_t *);
void ODEModel_freeJacobian(odeModel_t *);
ASTNode_t *ODEModel_constructDeterminant(const odeModel_t *);   /* Laplace expansion of the symbolic Jacobian; caller frees */
const ASTNode_t *ODEModel_getJacobianIJEntry(const odeModel_t *, int i, int j);
const ASTNode_t *ODEModel_getJacobianEntry(const odeModel_t *, const variableIndex_t *, const variableIndex_t *);

/* forward sensitivities (reference odeModel.h:346-353) */
odeSense_t *ODEModel_constructSensitivity(odeModel_t *);
odeSense_t *ODESense_create(odeModel_t *, cvodeSettings_t *);
void ODESense_free(odeSense_t *);
variableIndex_t *ODESense_getSensParamIndexByNum(const odeSense_t *, int);
int ODESense_getNeq(const odeSense_t *);
int ODESense_getNsens(const odeSense_t *);
const ASTNode_t *ODESense_getSensIJEntry(const odeSense_t *, int i, int j);
const ASTNode_t *ODESense_getSensEntry(const odeSense_t *, const variableIndex_t *, const variableIndex_t *);

/* code generation seam (reference: src/processAST.c:2057-2311 generateAST/generateMacros,
   src/odeModel.c:2600-2918 function generators, src/compiler.c:428-509 Compiler_*) */
typedef struct charBuffer charBuffer_t;
charBuffer_t *CharBuffer_create(void);
void CharBuffer_free(charBuffer_t *);
void CharBuffer_append(charBuffer_t *, const char *);
void CharBuffer_appendInt(charBuffer_t *, long);
void CharBuffer_appendDouble(charBuffer_t *, double);   /* "%.17g", not the reference's "%g" */
const char *CharBuffer_getBuffer(charBuffer_t *);
void generateAST(charBuffer_t *expressionStream, const ASTNode_t *node);
void generateMacros(charBuffer_t *buffer);

/* Emits the model-specific CUDA translation unit: device functions ode_f / jacobi_f /
   rules / event_f plus the kernel specialisation macros.  vary[] are value[] indices that
   are per-trajectory inputs (everything else is uniform).  Caller frees. */
char *ODEModel_generateCUDASource(odeModel_t *, const cvodeSettings_t *, int nvary, const int *vary,
                                  int nstore, const int *store, const double *base /* values after reset, or NULL */);
/* the same with the forward-sensitivity right-hand side (os may be NULL) */
char *ODEModel_generateCUDASourceSens(odeModel_t *, const cvodeSettings_t *, int nvary, const int *vary,
                                      int nstore, const int *store, const double *base, const odeSense_t *os);
/* Host-C flavour of the same generated functions (the reference's compiled mode), used by
   tests and by the CPU baseline harness.  Caller frees. */
/* text of flux_row: the indexed kinetic laws of SBMLResults_fluxASTs as one device function over a stored row */
char *ODEModel_generateCUDAFluxSection(int nflux, ASTNode_t *const *kls);
char *ODEModel_generateCSource(odeModel_t *);
/* ... plus sense_f (reference src/odeModel.c:2994-3100) when os is given */
char *ODEModel_generateCSourceSens(odeModel_t *, const odeSense_t *os);

compiled_code_t *Compiler_compile(const char *sourceCode);
void *CompiledCode_getFunction(compiled_code_t *, const char *symbol);
void CompiledCode_free(compiled_code_t *);
/* reference odeModel.h:356-363, the callback seam: compile the model's right-hand side / Jacobian / event functions (here:
   the model's sm_100a kernel module under default settings, kept in the model's plan cache) and hand out the compiled
   callbacks.  The reference returns host function pointers (CVRhsFn, CVDlsDenseJacFn, CVSensRhs1Fn); here the handle is the
   CUfunction of the module's kernel -- opaque to the caller, NULL + SolverError when the Jacobian (130501) or the
   parametric matrix (130505) is missing or no device can load the module. */
int ODEModel_compileCVODEFunctions(odeModel_t *);
int ODESense_compileCVODESenseFunctions(odeSense_t *);
void *ODEModel_getCompiledCVODERHSFunction(odeModel_t *);
void *ODEModel_getCompiledCVODEJacobianFunction(odeModel_t *);
void *ODESense_getCompiledCVODESenseFunction(odeSense_t *);

#ifdef __cplusplus
}
#endif
#endif
