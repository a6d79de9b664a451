/*
 * ode_model.c -- from an SBML reaction network to the indexed ODE system.
 *
 * Restates, on the object model of sbmlModel.h, what the reference does in
 *   src/odeConstruct.c:85-147   Model_reduceToOdes  (steps C.1-C.8)
 *   src/odeConstruct.c:253-386  Model_createOdes    (species ODEs + one assignment rule per reaction)
 *   src/odeConstruct.c:405-584  Species_odeFromReactions
 *   src/odeModel.c:109-150      ODEModel_create
 *   src/odeModel.c:253-418      topoSort
 *   src/odeModel.c:440-801      ODEModel_topologicalRuleSort
 *   src/odeModel.c:809-1122     ODEModel_fillStructures / setDiscontinuities
 *   src/odeModel.c:1728-1864    ODEModel_constructJacobian
 * The order of names[] (= index of every species/parameter in value[]), the
 * text of every equation and the order of the non-zero Jacobian entries are
 * part of the contract (unittest/test_odeModel.c:155-358, :564-719).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sbmlsolver/odeModel.h"
#include "sbmlsolver/processAST.h"
#include "sbmlsolver/solverError.h"

static char *sdup(const char *s) { char *r; if (!s) return NULL; r = (char *)malloc(strlen(s) + 1); strcpy(r, s); return r; }

SBMLDocument_t *parseModel(const char *file, int printMessage, int validate)
{
  SBMLDocument_t *d = readSBML(file);
  (void)printMessage; (void)validate;
  if (!d) SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_CANNOT_OPEN_FILE, "Can't read SBML file %s", file);
  return d;
}

/* ------------------------------------------------------ ODE construction */
ASTNode_t *Species_odeFromReactions(Species_t *s, Model_t *m)
{
  unsigned int j, k, errors = 0;
  ASTNode_t *ode = NULL, *simple;
  Compartment_t *c;

  for (j = 0; j < Model_getNumReactions(m); j++) {
    Reaction_t *r = Model_getReaction(m, j);
    KineticLaw_t *kl = r->kineticLaw;
    if (!kl) {
      SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_NO_KINETIC_LAW_FOUND_FOR_REACTION,
                        "The model has no kinetic law for reaction %s", r->id);
      ++errors;
    }
    for (k = 0; k < r->reactants.size; k++) {
      SpeciesReference_t *sref = (SpeciesReference_t *)r->reactants.items[k];
      ASTNode_t *term, *sym;
      if (strcmp(sref->species, s->id) != 0) continue;
      sym = ASTNode_create(); ASTNode_setName(sym, r->id);
      if (sref->stoichiometryMath) {
        term = ASTNode_createWithType(AST_TIMES);
        ASTNode_addChild(term, copyAST(sref->stoichiometryMath)); ASTNode_addChild(term, sym);
      } else if (sref->stoichiometry == 1.) term = sym;
      else {
        ASTNode_t *st = ASTNode_create(); ASTNode_setReal(st, sref->stoichiometry);
        term = ASTNode_createWithType(AST_TIMES);
        ASTNode_addChild(term, st); ASTNode_addChild(term, sym);
      }
      if (kl) AST_replaceNameByParameters(term, &kl->parameters);
      {                                         /* consumed: ode = (ode) - term, or -term first */
        ASTNode_t *n = ASTNode_createWithType(AST_MINUS);
        if (ode) ASTNode_addChild(n, ode);
        ASTNode_addChild(n, term);
        ode = n;
      }
    }
    for (k = 0; k < r->products.size; k++) {
      SpeciesReference_t *sref = (SpeciesReference_t *)r->products.items[k];
      ASTNode_t *term, *sym, *st;
      if (strcmp(sref->species, s->id) != 0) continue;
      sym = ASTNode_create(); ASTNode_setName(sym, r->id);
      term = ASTNode_createWithType(AST_TIMES);                  /* produced: always stoich * reaction */
      if (sref->stoichiometryMath) st = copyAST(sref->stoichiometryMath);
      else { st = ASTNode_create(); ASTNode_setReal(st, sref->stoichiometry); }
      ASTNode_addChild(term, st); ASTNode_addChild(term, sym);
      if (kl) AST_replaceNameByParameters(term, &kl->parameters);
      if (!ode) ode = term;
      else { ASTNode_t *n = ASTNode_createWithType(AST_PLUS); ASTNode_addChild(n, ode); ASTNode_addChild(n, term); ode = n; }
    }
  }

  c = Model_getCompartmentById(m, s->compartment);
  if (ode) {
    if (!s->hasOnlySubstanceUnits && c && c->spatialDimensions != 0) {
      ASTNode_t *n = ASTNode_createWithType(AST_DIVIDE), *cn = ASTNode_create();
      ASTNode_setName(cn, s->compartment);
      ASTNode_addChild(n, ode); ASTNode_addChild(n, cn);
      ode = n;
    }
  } else { ode = ASTNode_create(); ASTNode_setInteger(ode, 0); }

  simple = simplifyAST(ode);
  ASTNode_free(ode);
  if (errors > 0) { ASTNode_free(simple); return NULL; }
  return simple;
}

static void replace_function_definitions(Model_t *m)
{
  unsigned int i, j, k;
  for (i = 0; i < Model_getNumFunctionDefinitions(m); i++) {
    FunctionDefinition_t *f = Model_getFunctionDefinition(m, i);
    for (j = 0; j < Model_getNumRules(m); j++) {
      Rule_t *rl = Model_getRule(m, j);
      if (rl->math) AST_replaceFunctionDefinition(rl->math, f->id, f->math);
    }
    for (j = 0; j < Model_getNumEvents(m); j++) {
      Event_t *e = Model_getEvent(m, j);
      for (k = 0; k < e->assignments.size; k++) {
        EventAssignment_t *ea = (EventAssignment_t *)e->assignments.items[k];
        if (ea->math) AST_replaceFunctionDefinition(ea->math, f->id, f->math);
      }
      if (e->trigger) AST_replaceFunctionDefinition(e->trigger, f->id, f->math);
    }
  }
}

Model_t *Model_reduceToOdes(Model_t *m)
{
  unsigned int i, j;
  int errors = 0, nalg = 0;
  Model_t *ode = Model_create(m->level, m->version);

  /* C.1 inits: compartments, parameters, species (amounts -> concentrations), function definitions */
  if (m->id) Model_setId(ode, m->id); else if (m->name) Model_setId(ode, m->name);
  for (i = 0; i < Model_getNumCompartments(m); i++) ListOf_append(&ode->compartments, Compartment_clone(Model_getCompartment(m, i)));
  for (i = 0; i < Model_getNumParameters(m); i++) ListOf_append(&ode->parameters, Parameter_clone(Model_getParameter(m, i)));
  for (i = 0; i < Model_getNumSpecies(m); i++) {
    Species_t *s = Species_clone(Model_getSpecies(m, i));
    if (s->isSetInitialAmount && !s->hasOnlySubstanceUnits) {
      Compartment_t *c = Model_getCompartmentById(ode, s->compartment);
      if (c && c->spatialDimensions != 0) {
        s->initialConcentration = s->initialAmount / c->size; s->isSetInitialConcentration = 1;
        s->isSetInitialAmount = 0;
      }
    }
    ListOf_append(&ode->species, s);
  }
  for (i = 0; i < Model_getNumFunctionDefinitions(m); i++) {
    FunctionDefinition_t *f = Model_getFunctionDefinition(m, i);
    Model_createFunctionDefinition(ode, f->id, copyAST(f->math));
  }
  /* C.2 predefined rate rules */
  for (j = 0; j < Model_getNumRules(m); j++) {
    Rule_t *rl = Model_getRule(m, j);
    if (rl->type == SBML_RATE_RULE && rl->math && rl->variable) Model_createRule(ode, SBML_RATE_RULE, rl->variable, copyAST(rl->math));
  }
  /* C.3 assignment rules */
  for (j = 0; j < Model_getNumRules(m); j++) {
    Rule_t *rl = Model_getRule(m, j);
    if (rl->type == SBML_ASSIGNMENT_RULE && rl->math && rl->variable) Model_createRule(ode, SBML_ASSIGNMENT_RULE, rl->variable, copyAST(rl->math));
  }
  /* C.4 initial assignments */
  for (i = 0; i < Model_getNumInitialAssignments(m); i++) {
    InitialAssignment_t *a = Model_getInitialAssignment(m, i);
    Model_createInitialAssignment(ode, a->symbol, copyAST(a->math));
  }
  /* C.5 ODEs from reactions: species without rule that are neither constant nor boundary */
  for (i = 0; i < Model_getNumSpecies(m); i++) {
    Species_t *s = Model_getSpecies(m, i);
    int found = 0;
    for (j = 0; j < Model_getNumRules(m); j++) {
      Rule_t *rule = Model_getRule(m, j);
      if ((rule->type == SBML_RATE_RULE || rule->type == SBML_ASSIGNMENT_RULE) && rule->variable && strcmp(s->id, rule->variable) == 0) found = 1;
    }
    if (!found && !s->constant && !s->boundaryCondition) {
      ASTNode_t *math = Species_odeFromReactions(s, m);
      if (!math) {
        errors++;
        SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_ODE_COULD_NOT_BE_CONSTRUCTED_FOR_SPECIES, "ODE could not be constructed for species %s", s->id);
      } else if ((ASTNode_getType(math) == AST_REAL && ASTNode_getReal(math) == 0.0) ||
                 (ASTNode_getType(math) == AST_INTEGER && ASTNode_getInteger(math) == 0))
        ASTNode_free(math);                   /* identically zero: stays a constant */
      else Model_createRule(ode, SBML_RATE_RULE, s->id, math);
    }
  }
  /* one non-constant parameter + assignment rule per reaction: <reactionId> = kinetic law */
  for (i = 0; i != Model_getNumReactions(m); i++) {
    Reaction_t *r = Model_getReaction(m, i);
    Parameter_t *p = Model_createParameter(ode, r->id, 0.0);
    p->isSetValue = 0; p->constant = 0;
    if (r->kineticLaw) {
      ASTNode_t *math = copyAST(r->kineticLaw->math);
      AST_replaceNameByParameters(math, &r->kineticLaw->parameters);
      Model_createRule(ode, SBML_ASSIGNMENT_RULE, r->id, math);
    }
  }
  if (errors > 0) {
    SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_ODE_MODEL_COULD_NOT_BE_CONSTRUCTED, "ODE construction failed for %d variables.", errors);
    Model_free(ode);
    return NULL;
  }
  /* C.6a events */
  for (i = 0; i < Model_getNumEvents(m); i++) {
    ListOf_append(&ode->events, Event_clone(Model_getEvent(m, i)));
    if (!i) SolverError_error(WARNING_ERROR_TYPE, SOLVER_ERROR_THE_MODEL_CONTAINS_EVENTS,
                              "The model contains events. The SBML_odeSolver implementation of events is not fully SBML conformant. "
                              "Results will depend on the simulation duration and the number of output steps.");
  }
  /* C.6b algebraic rules: copied, flagged, unsupported */
  for (i = 0; i < Model_getNumRules(m); i++) {
    Rule_t *rl = Model_getRule(m, i);
    if (rl->type == SBML_ALGEBRAIC_RULE && rl->math) {
      Model_createRule(ode, SBML_ALGEBRAIC_RULE, NULL, copyAST(rl->math));
      if (!nalg) SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_THE_MODEL_CONTAINS_ALGEBRAIC_RULES,
                                   "The model contains Algebraic Rules. SBML_odeSolver is unable to solve models of this type.");
      nalg++;
    }
  }
  if (nalg > 0) {
    SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_ODE_MODEL_COULD_NOT_BE_CONSTRUCTED, "Model contains %d algebraic rules.", nalg);
    free(ode->notes); ode->notes = sdup("DAE model");
  }
  /* C.8 inline function definitions */
  replace_function_definitions(ode);
  return ode;
}

double Model_getValueById(Model_t *m, const char *id)
{
  Species_t *s; Parameter_t *p; Compartment_t *c;
  if ((p = Model_getParameterById(m, id)) != NULL && p->isSetValue) return p->value;
  if ((c = Model_getCompartmentById(m, id)) != NULL && c->isSetSize) return c->size;
  if ((s = Model_getSpeciesById(m, id)) != NULL) {
    if (s->isSetInitialConcentration) return s->initialConcentration;
    else if (s->isSetInitialAmount) {
      c = Model_getCompartmentById(m, s->compartment);
      if (c && c->spatialDimensions != 0 && !s->hasOnlySubstanceUnits) return s->initialAmount / c->size;
      return s->initialAmount;
    }
  }
  SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_REQUESTED_PARAMETER_NOT_FOUND,
                    "SBML Model doesn't provide a value for SBML ID %s, value defaults to 0!", id);
  return 0.0;
}

int Model_setValue(Model_t *m, const char *id, const char *rid, double value)
{
  unsigned int i; Compartment_t *c; Species_t *s; Parameter_t *p; Reaction_t *r;
  if (rid && (r = Model_getReactionById(m, rid)) != NULL && r->kineticLaw)
    for (i = 0; i < r->kineticLaw->parameters.size; i++) {
      p = (Parameter_t *)r->kineticLaw->parameters.items[i];
      if (strcmp(id, p->id) == 0) { p->value = value; p->isSetValue = 1; return 1; }
    }
  if ((c = Model_getCompartmentById(m, id)) != NULL) { c->size = value; c->isSetSize = 1; return 1; }
  if ((s = Model_getSpeciesById(m, id)) != NULL) {
    if (s->isSetInitialAmount) s->initialAmount = value;
    else { s->initialConcentration = value; s->isSetInitialConcentration = 1; }
    return 1;
  }
  if ((p = Model_getParameterById(m, id)) != NULL) { p->value = value; p->isSetValue = 1; return 1; }
  return 0;
}

/* ---------------------------------------------------------- topo sorting */
static void searchPath(int n, int **matrix, int start, int *required)
{
  int i;
  for (i = 0; i < n; i++)
    if (matrix[start][i] && !required[i]) { searchPath(n, matrix, i, required); required[i] = 1; }
}

/* Kahn's algorithm over "row i depends on column j"; a node enters the result only when it is
   both required and changed; the changed flag is inherited downstream and -- as in the
   reference -- written into the caller's array. */
List_t *topoSort(int **inMatrix, int n, int *changed, int *required)
{
  int i, j, ins, allchanged = 0, allrequired = 0;
  int **matrix = (int **)calloc((size_t)(n > 0 ? n : 1), sizeof(int *));
  List_t *sorted = List_create(), *noincome = List_create();

  for (i = 0; i < n; i++) { matrix[i] = (int *)malloc((size_t)n * sizeof(int)); memcpy(matrix[i], inMatrix[i], (size_t)n * sizeof(int)); }
  if (!changed) { allchanged = 1; changed = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int)); for (i = 0; i < n; i++) changed[i] = 1; }
  if (!required) { allrequired = 1; required = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int)); for (i = 0; i < n; i++) required[i] = 1; }

  for (i = 0; i < n; i++) {
    ins = 0; for (j = 0; j < n; j++) ins += matrix[i][j];
    if (!ins) { int *idx = (int *)malloc(sizeof(int)); *idx = i; List_add(noincome, idx); }
  }
  while (List_size(noincome)) {
    int *idx = (int *)List_remove(noincome, 0);
    int current = *idx;
    if (required[current] && changed[current]) List_add(sorted, idx); else free(idx);
    for (i = 0; i < n; i++)
      if (matrix[i][current]) {
        matrix[i][current] = 0;
        if (changed[current]) changed[i] = 1;
        ins = 0; for (j = 0; j < n; j++) ins += matrix[i][j];
        if (!ins) { int *id2 = (int *)malloc(sizeof(int)); *id2 = i; List_add(noincome, id2); }
      }
  }
  List_free(noincome);
  ins = 0;
  for (i = 0; i < n; i++) for (j = 0; j < n; j++) ins += matrix[i][j];
  if (ins) {
    SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_ODE_MODEL_CYCLIC_DEPENDENCY_IN_RULES, "Cyclic dependency found in topological sorting.");
    while (List_size(sorted)) free(List_remove(sorted, 0));
    List_free(sorted); sorted = NULL;
  }
  for (i = 0; i < n; i++) free(matrix[i]);
  free(matrix);
  if (allchanged) free(changed);
  if (allrequired) free(required);
  return sorted;
}

static void free_int_list(List_t *l) { while (List_size(l)) free(List_remove(l, 0)); List_free(l); }
static nonzeroElem_t *elem_new(int i, int j, ASTNode_t *ij)
{ nonzeroElem_t *e = (nonzeroElem_t *)malloc(sizeof *e); e->i = i; e->j = j; e->ij = ij; return e; }

static nonzeroElem_t **assignment_subset(odeModel_t *om, List_t *dep, int *count)
{
  unsigned int n; int k = 0; nonzeroElem_t **out;
  for (n = 0; n < List_size(dep); n++) { int idx = *(int *)List_get(dep, n); if (idx >= om->neq && idx < om->neq + om->nass) k++; }
  *count = k;
  out = (nonzeroElem_t **)calloc((size_t)(k > 0 ? k : 1), sizeof *out);
  k = 0;
  for (n = 0; n < List_size(dep); n++) {
    int idx = *(int *)List_get(dep, n);
    if (idx >= om->neq && idx < om->neq + om->nass) out[k++] = elem_new(idx, -1, om->assignment[idx - om->neq]);
  }
  return out;
}
static void mark_required(odeModel_t *om, const ASTNode_t *eq, int nvalues, int *required)
{
  int j, *tmp = ASTNode_getIndexArray(eq, nvalues);
  for (j = 0; j < nvalues; j++)
    if (!required[j] && tmp[j]) { searchPath(nvalues, om->dependencyMatrix, j, required); required[j] = 1; }
  free(tmp);
}

static int ODEModel_topologicalRuleSort(odeModel_t *om)
{
  int i, j, k, l, nvalues = om->neq + om->nass + om->nconst;
  int **matrix, *changedBySolver, *requiredForODEs, *requiredForEvents;
  unsigned int n; List_t *dep;

  om->initAssignmentOrder = NULL; om->assignmentOrder = NULL;
  om->assignmentsBeforeODEs = NULL; om->assignmentsBeforeEvents = NULL;

  /* 1: row i lists the values equation i reads (assignment rules; initial assignments for others) */
  matrix = (int **)calloc((size_t)(nvalues > 0 ? nvalues : 1), sizeof(int *));
  for (i = 0; i < nvalues; i++) {
    if (i >= om->neq && i < om->neq + om->nass) matrix[i] = ASTNode_getIndexArray(om->assignment[i - om->neq], nvalues);
    else matrix[i] = ASTNode_getIndexArray(om->indexInit[i] != -1 ? om->initAssignment[om->indexInit[i]] : NULL, nvalues);
  }
  om->dependencyMatrix = matrix;

  /* 2: complete ordering (initial assignments + assignment rules) */
  dep = topoSort(matrix, nvalues, NULL, NULL);
  if (!dep) {
    SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_ODE_MODEL_RULE_SORTING_FAILED,
                      "Topological sorting failed for complete rule set (initial assignments, assignments and kinetic laws) "
                      "Found cyclic dependency in rules. ");
    return 1;
  }
  om->assignmentOrder = (nonzeroElem_t **)calloc((size_t)(om->nass > 0 ? om->nass : 1), sizeof(nonzeroElem_t *));
  om->initAssignmentOrder = (nonzeroElem_t **)calloc((size_t)(om->nass + om->ninitAss > 0 ? om->nass + om->ninitAss : 1), sizeof(nonzeroElem_t *));
  k = l = 0;
  for (n = 0; n < List_size(dep); n++) {
    int idx = *(int *)List_get(dep, n);
    if (idx >= om->neq && idx < om->neq + om->nass) {
      om->initAssignmentOrder[k++] = elem_new(idx, -1, om->assignment[idx - om->neq]);
      om->assignmentOrder[l++] = elem_new(idx, -1, om->assignment[idx - om->neq]);
    } else if (om->indexInit[idx] != -1)
      om->initAssignmentOrder[k++] = elem_new(-1, idx, om->initAssignment[om->indexInit[idx]]);
  }
  free_int_list(dep);

  /* 3: evaluation-stage subsets; ODE variables (and time-dependent rules) count as changed */
  changedBySolver = (int *)calloc((size_t)(nvalues > 0 ? nvalues : 1), sizeof(int));
  for (i = 0; i < om->neq; i++) changedBySolver[i] = 1;
  for (i = 0; i < om->nass; i++) if (ASTNode_containsTime(om->assignment[i])) changedBySolver[om->neq + i] = 1;

  requiredForODEs = (int *)calloc((size_t)(nvalues > 0 ? nvalues : 1), sizeof(int));
  for (i = 0; i < om->neq; i++) mark_required(om, om->ode[i], nvalues, requiredForODEs);
  dep = topoSort(matrix, nvalues, changedBySolver, requiredForODEs);
  if (!dep) { free(changedBySolver); free(requiredForODEs); return 1; }
  om->assignmentsBeforeODEs = assignment_subset(om, dep, &om->nassbeforeodes);
  free_int_list(dep);

  requiredForEvents = (int *)calloc((size_t)(nvalues > 0 ? nvalues : 1), sizeof(int));
  for (i = 0; i < om->nevents; i++) {
    mark_required(om, om->event[i], nvalues, requiredForEvents);
    for (j = 0; j < om->neventAss[i]; j++) mark_required(om, om->eventAssignment[i][j], nvalues, requiredForEvents);
  }
  dep = topoSort(matrix, nvalues, changedBySolver, requiredForEvents);   /* changedBySolver as left by the previous sort */
  if (!dep) { free(changedBySolver); free(requiredForODEs); free(requiredForEvents); return 1; }
  om->assignmentsBeforeEvents = assignment_subset(om, dep, &om->nassbeforeevents);
  free_int_list(dep);

  free(changedBySolver); free(requiredForODEs); free(requiredForEvents);
  return 0;
}

/* --------------------------------------------------------- fill structures */
int ODEModel_getVariableIndexFields(const odeModel_t *om, const char *id)
{
  int i, nvalues = om->neq + om->nass + om->nconst + om->nalg;
  for (i = 0; i < nvalues && om->names[i]; i++) if (strcmp(id, om->names[i]) == 0) return i;
  return -1;
}

static odeModel_t *ODEModel_allocate(int neq, int nconst, int nass, int nalg)
{
  odeModel_t *om = (odeModel_t *)calloc(1, sizeof *om);
  int nvalues = neq + nalg + nass + nconst, n1 = nvalues > 0 ? nvalues : 1;
  om->names = (char **)calloc((size_t)n1, sizeof(char *));
  om->values = (double *)calloc((size_t)n1, sizeof(double));
  om->ode = (ASTNode_t **)calloc((size_t)(neq > 0 ? neq : 1), sizeof(ASTNode_t *));
  om->assignment = (ASTNode_t **)calloc((size_t)(nass > 0 ? nass : 1), sizeof(ASTNode_t *));
  om->algebraic = (ASTNode_t **)calloc((size_t)(nalg > 0 ? nalg : 1), sizeof(ASTNode_t *));
  om->neq = neq; om->nconst = nconst; om->nass = nass; om->nalg = nalg;
  return om;
}

static void set_discontinuities(odeModel_t *om, Model_t *ode)
{
  int i, j, nvalues = om->neq + om->nass + om->nconst;
  int nevents = ode ? (int)Model_getNumEvents(ode) : 0, ninit = ode ? (int)Model_getNumInitialAssignments(ode) : 0;
  om->ninitAss = ninit; om->nevents = nevents;
  om->indexInit = (int *)calloc((size_t)(nvalues > 0 ? nvalues : 1), sizeof(int));
  om->initIndex = (int *)calloc((size_t)(ninit > 0 ? ninit : 1), sizeof(int));
  om->initAssignment = (ASTNode_t **)calloc((size_t)(ninit > 0 ? ninit : 1), sizeof(ASTNode_t *));
  om->event = (ASTNode_t **)calloc((size_t)(nevents > 0 ? nevents : 1), sizeof(ASTNode_t *));
  om->neventAss = (int *)calloc((size_t)(nevents > 0 ? nevents : 1), sizeof(int));
  om->eventIndex = (int **)calloc((size_t)(nevents > 0 ? nevents : 1), sizeof(int *));
  om->eventAssignment = (ASTNode_t ***)calloc((size_t)(nevents > 0 ? nevents : 1), sizeof(ASTNode_t **));
  for (i = 0; i < nvalues; i++) om->indexInit[i] = -1;
  for (i = 0; i < ninit; i++) {
    InitialAssignment_t *init = Model_getInitialAssignment(ode, (unsigned int)i);
    int idx = ODEModel_getVariableIndexFields(om, init->symbol);
    om->initIndex[i] = idx;
    if (idx >= 0) om->indexInit[idx] = i;
    om->initAssignment[i] = indexAST(init->math, nvalues, om->names);
  }
  for (i = 0; i < nevents; i++) {
    Event_t *e = Model_getEvent(ode, (unsigned int)i);
    int nea = (int)e->assignments.size;
    om->event[i] = indexAST(e->trigger, nvalues, om->names);
    om->neventAss[i] = nea;
    om->eventIndex[i] = (int *)calloc((size_t)(nea > 0 ? nea : 1), sizeof(int));
    om->eventAssignment[i] = (ASTNode_t **)calloc((size_t)(nea > 0 ? nea : 1), sizeof(ASTNode_t *));
    for (j = 0; j < nea; j++) {
      EventAssignment_t *ea = (EventAssignment_t *)e->assignments.items[j];
      om->eventIndex[i][j] = ODEModel_getVariableIndexFields(om, ea->variable);
      om->eventAssignment[i][j] = indexAST(ea->math, nvalues, om->names);
    }
  }
}

static odeModel_t *ODEModel_fillStructures(Model_t *ode)
{
  unsigned int i, j;
  int neq = 0, nalg = 0, nass = 0, nconst, nvalues, found, npiecewise = 0;
  odeModel_t *om;

  for (j = 0; j < Model_getNumRules(ode); j++) {
    Rule_t *rl = Model_getRule(ode, j);
    if (rl->type == SBML_RATE_RULE) neq++;
    if (rl->type == SBML_ALGEBRAIC_RULE) nalg++;
    if (rl->type == SBML_ASSIGNMENT_RULE) nass++;
  }
  nvalues = (int)(Model_getNumCompartments(ode) + Model_getNumSpecies(ode) + Model_getNumParameters(ode));
  nconst = nvalues - nass - neq - nalg;
  om = ODEModel_allocate(neq, nconst, nass, nalg);

  /* names: [rate-rule variables | assigned variables | constants (compartments, species, parameters)] */
  neq = nass = 0;
  for (j = 0; j < Model_getNumRules(ode); j++) {
    Rule_t *rl = Model_getRule(ode, j);
    if (rl->type == SBML_RATE_RULE) om->names[neq++] = sdup(rl->variable);
    else if (rl->type == SBML_ASSIGNMENT_RULE) om->names[om->neq + nass++] = sdup(rl->variable);
  }
  nconst = 0;
#define ADD_CONST(ID) do { found = 0; for (j = 0; j < (unsigned int)(neq + nass); j++) if (strcmp((ID), om->names[j]) == 0) found++; \
    if (!found && neq + nass + nconst < nvalues) om->names[neq + nass + nconst++] = sdup(ID); } while (0)
  for (i = 0; i < Model_getNumCompartments(ode); i++) ADD_CONST(Model_getCompartment(ode, i)->id);
  for (i = 0; i < Model_getNumSpecies(ode); i++) ADD_CONST(Model_getSpecies(ode, i)->id);
  for (i = 0; i < Model_getNumParameters(ode); i++) ADD_CONST(Model_getParameter(ode, i)->id);
#undef ADD_CONST

  neq = nass = nalg = 0;
  for (j = 0; j < Model_getNumRules(ode); j++) {
    Rule_t *rl = Model_getRule(ode, j);
    ASTNode_t *math = indexAST(rl->math, nvalues, om->names);
    npiecewise += ASTNode_containsPiecewise(math);
    if (rl->type == SBML_RATE_RULE) om->ode[neq++] = math;
    else if (rl->type == SBML_ASSIGNMENT_RULE) om->assignment[nass++] = math;
    else om->algebraic[nalg++] = math;
  }
  om->simple = ode;
  om->npiecewise = npiecewise;

  /* initial conditions and parameter values from the (reduced) SBML model */
  for (i = 0; i < (unsigned int)om->neq; i++) om->values[i] = Model_getValueById(ode, om->names[i]);
  for (i = (unsigned int)(om->neq + om->nass); i < (unsigned int)nvalues; i++) om->values[i] = Model_getValueById(ode, om->names[i]);

  set_discontinuities(om, ode);
  return om;
}

odeModel_t *ODEModel_create(Model_t *m)
{
  Model_t *ode; odeModel_t *om;
  if (!m) return NULL;
  ode = Model_reduceToOdes(m);
  if (!ode) return NULL;
  if (ode->notes && strcmp(ode->notes, "DAE model") == 0) { Model_free(ode); return NULL; }
  om = ODEModel_fillStructures(ode);
  if (!om) return NULL;
  om->m = m; om->d = NULL;
  om->hasCycle = ODEModel_topologicalRuleSort(om);
  return om;
}
odeModel_t *ODEModel_createFromSBML2(SBMLDocument_t *d)
{
  if (!d || !d->model) return NULL;
  return ODEModel_create(d->model);
}
odeModel_t *ODEModel_createFromFile(const char *file)
{
  SBMLDocument_t *d = parseModel(file, 0, 0);
  odeModel_t *om;
  if (!d) return NULL;
  om = ODEModel_createFromSBML2(d);
  if (!om) { SBMLDocument_free(d); return NULL; }
  om->d = d;                                   /* owned: freed with the odeModel */
  return om;
}

/* reference src/odeModel.c:1330-1420: model straight from ODE ASTs, names and values */
odeModel_t *ODEModel_createFromODEs(ASTNode_t **f, int neq, int nass, int nconst, char **names, double *values, Model_t *events)
{
  int i, nvalues = neq + nass + nconst;
  odeModel_t *om = ODEModel_allocate(neq, nconst, nass, 0);
  for (i = 0; i < nvalues; i++) { om->names[i] = sdup(names[i]); om->values[i] = values ? values[i] : 0.0; }
  for (i = 0; i < neq; i++) { om->ode[i] = indexAST(f[i], nvalues, om->names); om->npiecewise += ASTNode_containsPiecewise(om->ode[i]); }
  for (i = 0; i < nass; i++) { om->assignment[i] = indexAST(f[neq + i], nvalues, om->names); om->npiecewise += ASTNode_containsPiecewise(om->assignment[i]); }
  om->simple = NULL; om->m = NULL; om->d = NULL;
  set_discontinuities(om, events);
  om->hasCycle = ODEModel_topologicalRuleSort(om);
  return om;
}

extern void ODEModel_freeKernelCache(odeModel_t *om);   /* batch.c */

static void free_elems(nonzeroElem_t **e, int n) { int i; if (!e) return; for (i = 0; i < n; i++) free(e[i]); free(e); }
void ODEModel_free(odeModel_t *om)
{
  int i, j, nvalues;
  if (!om) return;
  nvalues = om->neq + om->nass + om->nconst + om->nalg;
  ODEModel_freeKernelCache(om);
  ODEModel_freeJacobian(om);
  for (i = 0; i < nvalues; i++) free(om->names[i]);
  free(om->names); free(om->values);
  for (i = 0; i < om->neq; i++) ASTNode_free(om->ode[i]);
  for (i = 0; i < om->nass; i++) ASTNode_free(om->assignment[i]);
  for (i = 0; i < om->nalg; i++) ASTNode_free(om->algebraic[i]);
  free(om->ode); free(om->assignment); free(om->algebraic);
  if (om->dependencyMatrix) { for (i = 0; i < om->neq + om->nass + om->nconst; i++) free(om->dependencyMatrix[i]); free(om->dependencyMatrix); }
  free_elems(om->assignmentOrder, om->assignmentOrder ? om->nass : 0);
  free_elems(om->initAssignmentOrder, om->initAssignmentOrder ? om->nass + om->ninitAss : 0);
  free_elems(om->assignmentsBeforeODEs, om->nassbeforeodes);
  free_elems(om->assignmentsBeforeEvents, om->nassbeforeevents);
  for (i = 0; i < om->ninitAss; i++) ASTNode_free(om->initAssignment[i]);
  free(om->initAssignment); free(om->indexInit); free(om->initIndex);
  for (i = 0; i < om->nevents; i++) {
    ASTNode_free(om->event[i]);
    for (j = 0; j < om->neventAss[i]; j++) ASTNode_free(om->eventAssignment[i][j]);
    free(om->eventIndex[i]); free(om->eventAssignment[i]);
  }
  free(om->event); free(om->neventAss); free(om->eventIndex); free(om->eventAssignment);
  if (om->simple) Model_free(om->simple);
  if (om->d) SBMLDocument_free(om->d);
  free(om);
}

/* ---------------------------------------------------------------- Jacobian */
int ODEModel_constructJacobian(odeModel_t *om)
{
  int i, j, failed = 0, nvalues, cap = 16;
  nonzeroElem_t **sparse;
  if (!om) return 0;
  nvalues = om->neq + om->nass + om->nconst;
  om->jacob = (ASTNode_t ***)calloc((size_t)(om->neq > 0 ? om->neq : 1), sizeof(ASTNode_t **));
  for (i = 0; i < om->neq; i++) om->jacob[i] = (ASTNode_t **)calloc((size_t)om->neq, sizeof(ASTNode_t *));
  sparse = (nonzeroElem_t **)malloc((size_t)cap * sizeof *sparse);
  om->sparsesize = 0;

  for (i = 0; i < om->neq; i++) {
    ASTNode_t *ode = copyAST(om->ode[i]);
    /* inline assignment rules, last rule first, so rules may refer to earlier assigned variables */
    for (j = om->nass - 1; j >= 0; j--) AST_replaceNameByFormula(ode, om->names[om->neq + j], om->assignment[j]);
    for (j = 0; j < om->neq; j++) {
      ASTNode_t *fprime = differentiateAST(ode, om->names[j]);
      ASTNode_t *simple = simplifyAST(fprime);
      ASTNode_t *index = indexAST(simple, nvalues, om->names);
      double val = 1;
      unsigned int k; List_t *names;
      ASTNode_free(fprime); ASTNode_free(simple);
      om->jacob[i][j] = index;
      if (ASTNode_isInteger(index)) val = (double)ASTNode_getInteger(index);
      if (ASTNode_isReal(index)) val = ASTNode_getReal(index);
      if (val != 0.0) {                         /* row-major list of structurally non-zero entries */
        if (om->sparsesize == cap) { cap *= 2; sparse = (nonzeroElem_t **)realloc(sparse, (size_t)cap * sizeof *sparse); }
        sparse[om->sparsesize++] = elem_new(i, j, index);
      }
      names = ASTNode_getListOfNodes(index, (ASTNodePredicate)ASTNode_isName);
      for (k = 0; k < List_size(names); k++) {
        const char *nm = ASTNode_getName((ASTNode_t *)List_get(names, k));
        if (nm && strcmp(nm, "differentiation_failed") == 0) failed++;
      }
      List_free(names);
    }
    ASTNode_free(ode);
  }
  if (failed != 0) {
    SolverError_error(WARNING_ERROR_TYPE, SOLVER_ERROR_ENTRIES_OF_THE_JACOBIAN_MATRIX_COULD_NOT_BE_CONSTRUCTED,
                      "%d entries of the Jacobian matrix could not be constructed, due to failure of differentiation. "
                      "Cvode will use internal approximation of the Jacobian instead.", failed);
    om->jacobian = 0;
  } else om->jacobian = 1;
  om->jacobianFailed = failed;
  om->jacobSparse = sparse;
  return om->jacobian;
}

/* ------------------------------------------------- forward sensitivities */
/* src/odeModel.c:1963-2100 ODESense_constructMatrix */
static int ODESense_constructMatrix(odeSense_t *os, odeModel_t *om)
{
  int i, j, l, failed = 0, nvalues = om->neq + om->nass + om->nconst, cap = 16;
  nonzeroElem_t **sparse = (nonzeroElem_t **)malloc((size_t)cap * sizeof *sparse);
  os->sens = (ASTNode_t ***)calloc((size_t)(om->neq > 0 ? om->neq : 1), sizeof(ASTNode_t **));
  os->sensLogic = (int **)calloc((size_t)(om->neq > 0 ? om->neq : 1), sizeof(int *));
  os->sparsesize = 0;
  for (i = 0; i < om->neq; i++) {
    ASTNode_t *ode = copyAST(om->ode[i]);
    os->sens[i] = (ASTNode_t **)calloc((size_t)(os->nsensP > 0 ? os->nsensP : 1), sizeof(ASTNode_t *));
    os->sensLogic[i] = (int *)calloc((size_t)(os->nsensP > 0 ? os->nsensP : 1), sizeof(int));
    for (j = om->nass - 1; j >= 0; j--) AST_replaceNameByFormula(ode, om->names[om->neq + j], om->assignment[j]);
    l = 0;
    for (j = 0; j < os->nsens; j++) {
      ASTNode_t *fprime, *simple, *index; double val = 1; unsigned int k; List_t *names;
      if (os->index_sens[j] < om->neq) continue;                /* variables have no column */
      fprime = differentiateAST(ode, om->names[os->index_sens[j]]);
      simple = simplifyAST(fprime);
      index = indexAST(simple, nvalues, om->names);
      ASTNode_free(fprime); ASTNode_free(simple);
      os->sens[i][l] = index;
      if (ASTNode_isInteger(index)) val = (double)ASTNode_getInteger(index);
      if (ASTNode_isReal(index)) val = ASTNode_getReal(index);
      if (val != 0.0) {
        if (os->sparsesize == cap) { cap *= 2; sparse = (nonzeroElem_t **)realloc(sparse, (size_t)cap * sizeof *sparse); }
        sparse[os->sparsesize++] = elem_new(i, j, index);
        os->sensLogic[i][l] = 1;
      }
      l++;
      names = ASTNode_getListOfNodes(index, (ASTNodePredicate)ASTNode_isName);
      for (k = 0; k < List_size(names); k++) {
        const char *nm = ASTNode_getName((ASTNode_t *)List_get(names, k));
        if (nm && strcmp(nm, "differentiation_failed") == 0) failed++;
      }
      List_free(names);
    }
    ASTNode_free(ode);
  }
  os->sensSparse = sparse;
  if (failed != 0) {
    SolverError_error(WARNING_ERROR_TYPE, SOLVER_ERROR_ENTRIES_OF_THE_PARAMETRIC_MATRIX_COULD_NOT_BE_CONSTRUCTED,
                      "%d entries of the parametric `Jacobian' matrix could not be constructed, due to failure of "
                      "differentiation. Cvode will use internal approximation instead.", failed);
    return 0;
  }
  return 1;
}
/* src/odeModel.c:2188-2278 */
odeSense_t *ODESense_create(odeModel_t *om, cvodeSettings_t *opt)
{
  int i, k, all, construct = 0;
  odeSense_t *os;
  if (!om) return NULL;
  os = (odeSense_t *)calloc(1, sizeof *os);
  if (!opt) { all = 1; construct = 1; }
  else { all = (opt->sensIDs == NULL); if (opt->DoAdjoint || om->jacobian) construct = 1; }
  os->om = om; os->neq = om->neq;
  os->nsens = all ? om->nconst : opt->nsens;
  os->index_sens = (int *)calloc((size_t)(os->nsens > 0 ? os->nsens : 1), sizeof(int));
  os->index_sensP = (int *)calloc((size_t)(os->nsens > 0 ? os->nsens : 1), sizeof(int));
  if (all) {
    for (i = 0; i < os->nsens; i++) { os->index_sens[i] = om->neq + om->nass + i; os->index_sensP[i] = i; }
    os->nsensP = om->nconst;
  } else {
    k = 0;
    for (i = 0; i < os->nsens; i++) {
      os->index_sens[i] = ODEModel_getVariableIndexFields(om, opt->sensIDs[i]);
      if (os->index_sens[i] < 0 || (os->index_sens[i] >= om->neq && os->index_sens[i] < om->neq + om->nass)) {
        SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_SYMBOL_IS_NOT_IN_MODEL, "sensitivity requested for %s, which is not a variable or constant of the model", opt->sensIDs[i]);
        ODESense_free(os);
        return NULL;
      }
      if (os->index_sens[i] < om->neq) os->index_sensP[i] = -1;
      else os->index_sensP[i] = k++;
    }
    os->nsensP = k;
  }
  os->sensitivity = construct ? ODESense_constructMatrix(os, om) : 0;
  return os;
}
odeSense_t *ODEModel_constructSensitivity(odeModel_t *om) { return ODESense_create(om, NULL); }
void ODESense_free(odeSense_t *os)
{
  int i, j;
  if (!os) return;
  if (os->sens) for (i = 0; i < os->neq; i++) { if (os->sens[i]) for (j = 0; j < os->nsensP; j++) ASTNode_free(os->sens[i][j]); free(os->sens[i]); }
  if (os->sensLogic) for (i = 0; i < os->neq; i++) free(os->sensLogic[i]);
  for (i = 0; i < os->sparsesize; i++) free(os->sensSparse[i]);
  free(os->sens); free(os->sensLogic); free(os->sensSparse); free(os->index_sens); free(os->index_sensP); free(os);
}
variableIndex_t *ODESense_getSensParamIndexByNum(const odeSense_t *os, int i)
{ return (i >= 0 && i < os->nsens) ? ODEModel_getVariableIndexByNum(os->om, os->index_sens[i]) : NULL; }
int ODESense_getNeq(const odeSense_t *os) { return os->neq; }
int ODESense_getNsens(const odeSense_t *os) { return os->nsens; }
const ASTNode_t *ODESense_getSensIJEntry(const odeSense_t *os, int i, int j)
{
  if (!os->sens || i < 0 || i >= os->neq || j < 0 || j >= os->nsens || os->index_sensP[j] < 0) return NULL;
  return os->sens[i][os->index_sensP[j]];
}
const ASTNode_t *ODESense_getSensEntry(const odeSense_t *os, const variableIndex_t *vi1, const variableIndex_t *vi2)
{
  int j;
  for (j = 0; j < os->nsens; j++) if (os->index_sens[j] == vi2->index) return ODESense_getSensIJEntry(os, vi1->index, j);
  return NULL;
}

/* ------------------------------------------------------------ determinant of the Jacobian
   Reference: ODEModel_constructDeterminant -> determinantNAST (src/odeModel.c:1950-1956).  Laplace expansion along the
   first row of the symbolic Jacobian, structurally zero entries skipped; the result is a well-formed expression tree
   (the reference's printout of its own tree has empty operands, unittest/test_odeModel.c:734-800: that text is not
   reproduced). */
static int entry_is_zero(const ASTNode_t *e)
{
  if (ASTNode_isInteger(e)) return ASTNode_getInteger(e) == 0;
  if (ASTNode_isReal(e)) return ASTNode_getReal(e) == 0.0;
  return 0;
}
static ASTNode_t *minor_det(ASTNode_t ***jac, const int *rows, const int *cols, int n)
{
  ASTNode_t *sum = NULL; int j, k, sign = 1;
  if (n == 1) return copyAST(jac[rows[0]][cols[0]]);
  for (j = 0; j < n; j++, sign = -sign) {
    int *sub; ASTNode_t *m, *term;
    if (entry_is_zero(jac[rows[0]][cols[j]])) continue;
    sub = (int *)malloc((size_t)(n - 1) * sizeof(int));
    for (k = 0; k < n - 1; k++) sub[k] = cols[k < j ? k : k + 1];
    m = minor_det(jac, rows + 1, sub, n - 1);
    free(sub);
    if (!m) continue;
    term = ASTNode_createWithType(AST_TIMES);
    ASTNode_addChild(term, copyAST(jac[rows[0]][cols[j]])); ASTNode_addChild(term, m);
    if (!sum) {
      if (sign < 0) { ASTNode_t *neg = ASTNode_createWithType(AST_MINUS); ASTNode_addChild(neg, term); sum = neg; } else sum = term;
    } else {
      ASTNode_t *op = ASTNode_createWithType(sign < 0 ? AST_MINUS : AST_PLUS);
      ASTNode_addChild(op, sum); ASTNode_addChild(op, term); sum = op;
    }
  }
  return sum;                                     /* NULL: the minor vanishes structurally */
}
ASTNode_t *ODEModel_constructDeterminant(const odeModel_t *om)
{
  int i, *idx; ASTNode_t *det;
  if (!om || !om->jacob || om->jacobian != 1 || om->neq < 1) return NULL;
  idx = (int *)malloc((size_t)om->neq * sizeof(int));
  for (i = 0; i < om->neq; i++) idx[i] = i;
  det = minor_det(om->jacob, idx, idx, om->neq);
  free(idx);
  if (!det) { det = ASTNode_create(); ASTNode_setInteger(det, 0); }
  return det;
}

/* reference src/processAST.c:1152-1251 (processAST.h:62, "experimental"): the determinant of an N x N matrix of expression
   trees, as a simplified tree; the caller frees it.  Same cofactor expansion as above over a caller-supplied matrix. */
ASTNode_t *determinantNAST(ASTNode_t ***A, int N)
{
  int i, *idx; ASTNode_t *det, *simple;
  if (!A || N < 1) return NULL;
  idx = (int *)malloc((size_t)N * sizeof(int));
  for (i = 0; i < N; i++) idx[i] = i;
  det = minor_det(A, idx, idx, N);
  free(idx);
  if (!det) { det = ASTNode_create(); ASTNode_setInteger(det, 0); }
  simple = simplifyAST(det);
  ASTNode_free(det);
  return simple;
}

void ODEModel_freeJacobian(odeModel_t *om)
{
  int i, j;
  if (!om || !om->jacob) return;
  for (i = 0; i < om->neq; i++) { for (j = 0; j < om->neq; j++) ASTNode_free(om->jacob[i][j]); free(om->jacob[i]); }
  free(om->jacob); om->jacob = NULL;
  for (i = 0; i < om->sparsesize; i++) free(om->jacobSparse[i]);
  free(om->jacobSparse); om->jacobSparse = NULL; om->sparsesize = 0; om->jacobian = 0;
}
const ASTNode_t *ODEModel_getJacobianIJEntry(const odeModel_t *om, int i, int j) { return om->jacob ? om->jacob[i][j] : NULL; }
const ASTNode_t *ODEModel_getJacobianEntry(const odeModel_t *om, const variableIndex_t *vi1, const variableIndex_t *vi2)
{ return ODEModel_getJacobianIJEntry(om, vi1->type_index, vi2->type_index); }

/* ------------------------------------------------------- accessor functions */
const Model_t *ODEModel_getModel(odeModel_t *om) { return om->m; }
const Model_t *ODEModel_getOdeSBML(odeModel_t *om) { return om->simple; }
int ODEModel_getNumValues(const odeModel_t *om) { return om->neq + om->nass + om->nconst + om->nalg; }
int ODEModel_getNeq(const odeModel_t *om) { return om->neq; }
int ODEModel_getNalg(const odeModel_t *om) { return om->nalg; }
int ODEModel_getNumAssignments(const odeModel_t *om) { return om->nass; }
int ODEModel_getNumConstants(const odeModel_t *om) { return om->nconst; }
int ODEModel_hasVariable(const odeModel_t *om, const char *id) { return ODEModel_getVariableIndexFields(om, id) != -1; }
int ODEModel_hasCycle(const odeModel_t *om) { return om->hasCycle; }
int ODEModel_getNumAssignmentsBeforeODEs(const odeModel_t *om) { return om->nassbeforeodes; }
int ODEModel_getNumAssignmentsBeforeEvents(const odeModel_t *om) { return om->nassbeforeevents; }
int ODEModel_getNumJacobiElements(const odeModel_t *om) { return om->sparsesize; }
const nonzeroElem_t *ODEModel_getAssignmentOrder(odeModel_t *om, int i) { return i < om->nass ? om->assignmentOrder[i] : NULL; }
const nonzeroElem_t *ODEModel_getAssignmentBeforeODEs(odeModel_t *om, int i) { return i < om->nassbeforeodes ? om->assignmentsBeforeODEs[i] : NULL; }
const nonzeroElem_t *ODEModel_getAssignmentBeforeEvents(odeModel_t *om, int i) { return i < om->nassbeforeevents ? om->assignmentsBeforeEvents[i] : NULL; }
const nonzeroElem_t *ODEModel_getJacobiElement(const odeModel_t *om, int i) { return (i < 0 || i >= om->sparsesize) ? NULL : om->jacobSparse[i]; }
const ASTNode_t *NonzeroElement_getEquation(nonzeroElem_t *e) { return e->ij; }
const char *NonzeroElement_getVariableName(nonzeroElem_t *e, odeModel_t *om) { return e->i != -1 ? om->names[e->i] : om->names[e->j]; }
const char *NonzeroElement_getVariable2Name(nonzeroElem_t *e, odeModel_t *om) { return (e->i == -1 || e->j == -1) ? NULL : om->names[e->j]; }

variableIndex_t *ODEModel_getVariableIndexByNum(const odeModel_t *om, int i)
{
  variableIndex_t *vi;
  if (i < 0 || i >= ODEModel_getNumValues(om)) {
    SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_SYMBOL_IS_NOT_IN_MODEL, "No such variable in the model");
    return NULL;
  }
  vi = (variableIndex_t *)malloc(sizeof *vi);
  vi->index = i;
  if (i < om->neq) { vi->type = ODE_VARIABLE; vi->type_index = i; }
  else if (i < om->neq + om->nass) { vi->type = ASSIGNMENT_VARIABLE; vi->type_index = i - om->neq; }
  else if (i < om->neq + om->nass + om->nconst) { vi->type = CONSTANT; vi->type_index = i - om->neq - om->nass; }
  else { vi->type = ALGEBRAIC_VARIABLE; vi->type_index = i - om->neq - om->nass - om->nconst; }
  return vi;
}
variableIndex_t *ODEModel_getOdeVariableIndex(const odeModel_t *om, int i) { return (0 <= i && i < om->neq) ? ODEModel_getVariableIndexByNum(om, i) : NULL; }
variableIndex_t *ODEModel_getAssignedVariableIndex(const odeModel_t *om, int i) { return (0 <= i && i < om->nass) ? ODEModel_getVariableIndexByNum(om, i + om->neq) : NULL; }
variableIndex_t *ODEModel_getConstantIndex(const odeModel_t *om, int i) { return (0 <= i && i < om->nconst) ? ODEModel_getVariableIndexByNum(om, i + om->neq + om->nass) : NULL; }
variableIndex_t *ODEModel_getVariableIndex(odeModel_t *om, const char *symbol)
{
  int index = symbol ? ODEModel_getVariableIndexFields(om, symbol) : -1;
  if (index == -1) {
    SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_SYMBOL_IS_NOT_IN_MODEL, "Symbol %s is not in the model", symbol ? symbol : "(null)");
    return NULL;
  }
  return ODEModel_getVariableIndexByNum(om, index);
}
void ODEModel_dumpNames(odeModel_t *om)
{
  int i;
  for (i = 0; i < ODEModel_getNumValues(om); i++) printf("%s ", om->names[i]);
  printf("\n");
}
const char *ODEModel_getVariableName(const odeModel_t *om, const variableIndex_t *vi) { return om->names[vi->index]; }
const char *VariableIndex_getName(const variableIndex_t *vi, const odeModel_t *om) { return om->names[vi->index]; }
int VariableIndex_getIndex(const variableIndex_t *vi) { return vi->index; }
void VariableIndex_free(variableIndex_t *vi) { free(vi); }
const ASTNode_t *ODEModel_getOde(const odeModel_t *om, const variableIndex_t *vi) { return (0 <= vi->index && vi->index < om->neq) ? om->ode[vi->index] : NULL; }
const ASTNode_t *ODEModel_getAssignment(const odeModel_t *om, const variableIndex_t *vi)
{ return (om->neq <= vi->index && vi->index < om->neq + om->nass) ? om->assignment[vi->type_index] : NULL; }

/* ---- smaller accessors of the reference's odeModel.h (src/odeModel.c:1241-1300, :1609-1720) ---- */
const Model_t *ODEModel_getInputSBML(odeModel_t *om) { return om->m; }
int ODEModel_getNumPiecewise(odeModel_t *om) { return om->npiecewise; }
int ODEModel_getNumEvents(odeModel_t *om) { return om->nevents; }
const ASTNode_t *ODEModel_getEventTrigger(odeModel_t *om, int i) { return (i >= 0 && i < om->nevents) ? om->event[i] : NULL; }
const ASTNode_t *ODEModel_getEventAssignment(odeModel_t *om, int i, int j)
{ return (i >= 0 && i < om->nevents && j >= 0 && j < om->neventAss[i]) ? om->eventAssignment[i][j] : NULL; }
variableIndex_t *NonzeroElement_getVariableIndex(nonzeroElem_t *nz, odeModel_t *om) { return ODEModel_getVariableIndexByNum(om, nz->i != -1 ? nz->i : nz->j); }
variableIndex_t *NonzeroElement_getVariable2Index(nonzeroElem_t *nz, odeModel_t *om) { return (nz->i == -1 || nz->j == -1) ? NULL : ODEModel_getVariableIndexByNum(om, nz->j); }
ASTNode_t *ASTNode_indexAST(const ASTNode_t *f, odeModel_t *om) { return indexAST(f, om->neq + om->nass + om->nconst, om->names); }
