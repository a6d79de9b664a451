/*
 * sbml_model.h -- minimal SBML object model (Level 1 and Level 2 core).
 *
 * Replaces the part of libSBML's C API that the reference front-end uses
 * (src/sbml.c:68-118, src/odeConstruct.c:85-866, src/odeSolver.c:354-440):
 * documents, models, compartments, species, parameters, reactions with
 * kinetic laws, rules, function definitions, initial assignments, events.
 * Attribute defaults follow the SBML specifications as libSBML applies them.
 */
#ifndef ODES_B200_SBML_MODEL_H
#define ODES_B200_SBML_MODEL_H
#include "sbmlsolver/ast.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef enum { SBML_RATE_RULE = 1, SBML_ASSIGNMENT_RULE, SBML_ALGEBRAIC_RULE } SBMLTypeCode_t;

typedef struct { void **items; unsigned int size, cap; } ListOf_t;

typedef struct Compartment {
  char *id; double size; int isSetSize; int spatialDimensions; int constant;
} Compartment_t;

typedef struct Species {
  char *id; char *compartment;
  double initialAmount, initialConcentration;
  int isSetInitialAmount, isSetInitialConcentration;
  int hasOnlySubstanceUnits, boundaryCondition, constant;
} Species_t;

typedef struct Parameter { char *id; double value; int isSetValue; int constant; } Parameter_t;

typedef struct SpeciesReference {
  char *species; double stoichiometry; ASTNode_t *stoichiometryMath;
} SpeciesReference_t;

typedef struct KineticLaw { ASTNode_t *math; ListOf_t parameters; char *formula; /* text of math, made on request */ } KineticLaw_t;

typedef struct Reaction {
  char *id; int reversible; int fast;
  ListOf_t reactants, products, modifiers;   /* SpeciesReference_t */
  KineticLaw_t *kineticLaw;
} Reaction_t;

typedef struct Rule { SBMLTypeCode_t type; char *variable; ASTNode_t *math; char *formula; /* text of math, made on request */ } Rule_t;
typedef struct FunctionDefinition { char *id; ASTNode_t *math; } FunctionDefinition_t;
typedef struct InitialAssignment { char *symbol; ASTNode_t *math; } InitialAssignment_t;
typedef struct EventAssignment { char *variable; ASTNode_t *math; } EventAssignment_t;
typedef struct Event {
  char *id; ASTNode_t *trigger; ASTNode_t *delay; ListOf_t assignments;
} Event_t;

typedef struct Model {
  unsigned int level, version;
  char *id, *name, *notes;
  ListOf_t functionDefinitions, compartments, species, parameters,
           initialAssignments, rules, reactions, events;
} Model_t;

typedef struct SBMLDocument { unsigned int level, version; Model_t *model; } SBMLDocument_t;

/* generic list */
void ListOf_append(ListOf_t *, void *);
unsigned int ListOf_size(const ListOf_t *);
void *ListOf_get(const ListOf_t *, unsigned int);

/* documents */
SBMLDocument_t *readSBML(const char *filename);
SBMLDocument_t *readSBMLFromString(const char *xml);
char *writeSBMLToString(const SBMLDocument_t *);       /* Level 2 Version 1 text; caller frees */
int writeSBML(const SBMLDocument_t *, const char *filename);
SBMLDocument_t *SBMLDocument_create(void);
void SBMLDocument_free(SBMLDocument_t *);
/* libSBML's reader object, as the reference's examples use it (examples/integrate.c:71-73, batch_integrate.c:111-113) */
typedef struct SBMLReader { int unused; } SBMLReader_t;
SBMLReader_t *SBMLReader_create(void);
SBMLDocument_t *SBMLReader_readSBML(SBMLReader_t *, const char *filename);
SBMLDocument_t *SBMLReader_readSBMLFromString(SBMLReader_t *, const char *xml);
void SBMLReader_free(SBMLReader_t *);
Model_t *SBMLDocument_getModel(SBMLDocument_t *);
unsigned int SBMLDocument_getLevel(const SBMLDocument_t *);

/* model construction (builder API) and access */
Model_t *Model_create(unsigned int level, unsigned int version);
Model_t *Model_clone(const Model_t *);
void Model_free(Model_t *);
void Model_setId(Model_t *, const char *);
const char *Model_getId(const Model_t *);

Compartment_t *Model_createCompartment(Model_t *, const char *id, double size);
Species_t *Model_createSpecies(Model_t *, const char *id, const char *compartment, double initialConcentration);
Parameter_t *Model_createParameter(Model_t *, const char *id, double value);
Reaction_t *Model_createReaction(Model_t *, const char *id);
SpeciesReference_t *Reaction_createReactant(Reaction_t *, const char *species, double stoichiometry);
SpeciesReference_t *Reaction_createProduct(Reaction_t *, const char *species, double stoichiometry);
SpeciesReference_t *Reaction_createModifier(Reaction_t *, const char *species);
KineticLaw_t *Reaction_createKineticLaw(Reaction_t *, const char *formula);
Parameter_t *KineticLaw_createParameter(KineticLaw_t *, const char *id, double value);
Rule_t *Model_createRule(Model_t *, SBMLTypeCode_t type, const char *variable, ASTNode_t *math /* adopted */);
Rule_t *Model_createRuleFromFormula(Model_t *, SBMLTypeCode_t type, const char *variable, const char *formula);
Event_t *Model_createEvent(Model_t *, const char *id, ASTNode_t *trigger /* adopted */);
EventAssignment_t *Event_createEventAssignment(Event_t *, const char *variable, ASTNode_t *math /* adopted */);
InitialAssignment_t *Model_createInitialAssignment(Model_t *, const char *symbol, ASTNode_t *math /* adopted */);
FunctionDefinition_t *Model_createFunctionDefinition(Model_t *, const char *id, ASTNode_t *lambda /* adopted */);

unsigned int Model_getNumCompartments(const Model_t *);
unsigned int Model_getNumSpecies(const Model_t *);
unsigned int Model_getNumParameters(const Model_t *);
unsigned int Model_getNumReactions(const Model_t *);
unsigned int Model_getNumRules(const Model_t *);
unsigned int Model_getNumEvents(const Model_t *);
unsigned int Model_getNumInitialAssignments(const Model_t *);
unsigned int Model_getNumFunctionDefinitions(const Model_t *);
Compartment_t *Model_getCompartment(const Model_t *, unsigned int);
Species_t *Model_getSpecies(const Model_t *, unsigned int);
Parameter_t *Model_getParameter(const Model_t *, unsigned int);
Reaction_t *Model_getReaction(const Model_t *, unsigned int);
Rule_t *Model_getRule(const Model_t *, unsigned int);
Event_t *Model_getEvent(const Model_t *, unsigned int);
InitialAssignment_t *Model_getInitialAssignment(const Model_t *, unsigned int);
FunctionDefinition_t *Model_getFunctionDefinition(const Model_t *, unsigned int);
Compartment_t *Model_getCompartmentById(const Model_t *, const char *);
Species_t *Model_getSpeciesById(const Model_t *, const char *);
Parameter_t *Model_getParameterById(const Model_t *, const char *);
Reaction_t *Model_getReactionById(const Model_t *, const char *);

Compartment_t *Compartment_clone(const Compartment_t *);
Species_t *Species_clone(const Species_t *);
Parameter_t *Parameter_clone(const Parameter_t *);
Event_t *Event_clone(const Event_t *);
void Parameter_setId(Parameter_t *, const char *);

/* further libSBML-named accessors the reference's own unit tests and examples call on these structures */
typedef void SBase_t;
const char *Compartment_getId(const Compartment_t *);
const char *Species_getId(const Species_t *);
const char *Parameter_getId(const Parameter_t *);
double Parameter_getValue(const Parameter_t *);
void Parameter_setValue(Parameter_t *, double);
void Parameter_setName(Parameter_t *, const char *);
void Parameter_setConstant(Parameter_t *, int);
void Compartment_setConstant(Compartment_t *, int);
Parameter_t *Parameter_create(void);
Parameter_t *Parameter_createWith(const char *id, const char *name);
void Parameter_free(Parameter_t *);
const char *Reaction_getId(const Reaction_t *);
KineticLaw_t *Reaction_getKineticLaw(const Reaction_t *);
const char *KineticLaw_getFormula(const KineticLaw_t *);       /* infix text of the math; owned by the kinetic law */
const ASTNode_t *KineticLaw_getMath(const KineticLaw_t *);
const char *Rule_getFormula(const Rule_t *);
const char *Rule_getVariable(const Rule_t *);
unsigned int Model_getNumUnitDefinitions(const Model_t *);                 /* not kept by this front end: 0 */
unsigned int Model_getNumSpeciesWithBoundaryCondition(const Model_t *);
unsigned int Model_getNumSpeciesTypes(const Model_t *);
unsigned int Model_getNumCompartmentTypes(const Model_t *);
unsigned int Model_getNumConstraints(const Model_t *);
ListOf_t *ListOf_create(void);
void ListOf_appendAndOwn(ListOf_t *, SBase_t *);
void ListOf_free(ListOf_t *);                                              /* the list, not its items */

#ifdef __cplusplus
}
#endif
#endif
