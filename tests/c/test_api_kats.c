/*
 * test_api_kats.c -- the reference's own unit tests for the small host-side API pieces on the path, restated against
 * this library (plain asserts instead of libcheck; same calls, same expected values):
 *   unittest/test_odeSolver.c:111-290     VarySettings_*  (fixture: VarySettings_allocate(3, 4))
 *   unittest/test_solverError.c:18-148    SolverError_*   (codes, order, dump format)
 *   unittest/test_sbmlResults.c:27-260    SBMLResults_* / TimeCourse_* accessors on a results object
 *   unittest/test_modelSimplify.c:40-240  AST_replaceNameBy{Formula,Name,Value,Parameters}, AST_replaceFunctionDefinition,
 *                                         AST_replaceConstants
 * Built and run by tests/test_c_api_kats.py (CPU tier; no device needed).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sbmlsolver/odeSolver.h"
#include "sbmlsolver/processAST.h"
#include "sbmlsolver/solverError.h"

static int failures = 0;
#define CK(cond) do { if (!(cond)) { fprintf(stderr, "%s:%d: failed: %s\n", __FILE__, __LINE__, #cond); failures++; } } while (0)
#define CK_STR(a, b) CK((a) != NULL && strcmp((a), (b)) == 0)

static varySettings_t *vs;
static void setup_vs(void) { vs = VarySettings_allocate(3, 4); }
static void teardown_vs(void) { VarySettings_free(vs); }
static void name_all(void) { VarySettings_setName(vs, 0, "foo", NULL); VarySettings_setName(vs, 1, "foo", "local"); VarySettings_setName(vs, 2, "bar", "local"); }

static void test_vary_settings(void)
{
  static const double p0[] = {0.0, 2.5, 4.0}, p1[] = {0.1, -0.1, 4.2}, p2[] = {0.2, -1.4, 5.9}, p3[] = {0.3, -2.9, 5.1}, p4[] = {0.4, -5.0, 4.1};
  varySettings_t *v = VarySettings_allocate(7, 77);
  CK(v != NULL); CK(v->nrparams == 7); CK(v->nrdesignpoints == 77);
  VarySettings_free(v);
  VarySettings_free(NULL);                                   /* freeing NULL is safe */
  setup_vs();
  CK(VarySettings_addDesignPoint(vs, p0) == 1); CK(VarySettings_addDesignPoint(vs, p1) == 2);
  CK(VarySettings_addDesignPoint(vs, p2) == 3); CK(VarySettings_addDesignPoint(vs, p3) == 4);
  CK(VarySettings_addDesignPoint(vs, p4) == 0);              /* full */
  teardown_vs();
  setup_vs();
  CK(VarySettings_addParameter(vs, "foo", NULL) == 1); CK_STR(vs->id[0], "foo"); CK(vs->rid[0] == NULL);
  CK(VarySettings_addParameter(vs, "foo", "local") == 2); CK_STR(vs->id[1], "foo"); CK_STR(vs->rid[1], "local");
  CK(VarySettings_addParameter(vs, "bar", "local") == 3); CK_STR(vs->id[2], "bar"); CK_STR(vs->rid[2], "local");
  CK(VarySettings_addParameter(vs, "failure", "full") == 0);
  teardown_vs();
  setup_vs();
  CK(VarySettings_setName(vs, 0, "foo", NULL) == 1); CK(VarySettings_setName(vs, 1, "foo", "local") == 1);
  CK(VarySettings_setName(vs, 2, "bar", "local") == 1); CK(VarySettings_setName(vs, 3, "failure", "fencepost") == 0);
  CK_STR(vs->id[0], "foo"); CK(vs->rid[0] == NULL); CK_STR(vs->id[1], "foo"); CK_STR(vs->rid[1], "local"); CK_STR(vs->id[2], "bar"); CK_STR(vs->rid[2], "local");
  teardown_vs();
  setup_vs();
  CK(VarySettings_setValue(vs, 0, 0, 7.7) == 1); CK(VarySettings_setValue(vs, 2, 2, 2.3) == 1);
  CK(VarySettings_setValue(vs, 4, 1, 0.1) == 0); CK(VarySettings_setValue(vs, 2, 3, 0.2) == 0);
  CK(vs->params[0][0] == 7.7); CK(vs->params[2][2] == 2.3);
  CK(VarySettings_getValue(vs, 0, 0) == 7.7); CK(VarySettings_getValue(vs, 2, 2) == 2.3);
  teardown_vs();
  setup_vs();
  name_all();
  CK(VarySettings_setValueByID(vs, 0, "foo", NULL, 1.2) == 1); CK(VarySettings_setValueByID(vs, 1, "foo", NULL, 2.3) == 1);
  CK(VarySettings_setValueByID(vs, 1, "foo", "local", 3.4) == 1); CK(VarySettings_setValueByID(vs, 3, "bar", "local", 4.5) == 1);
  CK(VarySettings_getValue(vs, 0, 0) == 1.2); CK(VarySettings_getValue(vs, 1, 0) == 2.3);
  CK(VarySettings_getValue(vs, 1, 1) == 3.4); CK(VarySettings_getValue(vs, 3, 2) == 4.5);
  CK(VarySettings_getValueByID(vs, 0, "foo", NULL) == 1.2); CK(VarySettings_getValueByID(vs, 1, "foo", NULL) == 2.3);
  CK(VarySettings_getValueByID(vs, 1, "foo", "local") == 3.4); CK(VarySettings_getValueByID(vs, 3, "bar", "local") == 4.5);
  CK_STR(VarySettings_getName(vs, 0), "foo"); CK_STR(VarySettings_getName(vs, 1), "foo"); CK_STR(VarySettings_getName(vs, 2), "bar");
  CK(VarySettings_getName(vs, 3) == NULL);
  CK(VarySettings_getReactionName(vs, 0) == NULL); CK_STR(VarySettings_getReactionName(vs, 1), "local");
  CK_STR(VarySettings_getReactionName(vs, 2), "local"); CK(VarySettings_getReactionName(vs, 3) == NULL);
  teardown_vs();
}

static void add_some_errors(void)
{
  SolverError_error(FATAL_ERROR_TYPE, SOLVER_ERROR_ODE_COULD_NOT_BE_CONSTRUCTED_FOR_SPECIES, "fatal0");
  SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_THE_MODEL_CONTAINS_EVENTS, "error0");
  SolverError_error(WARNING_ERROR_TYPE, SOLVER_ERROR_THE_MODEL_CONTAINS_ALGEBRAIC_RULES, "warning0");
  SolverError_error(MESSAGE_ERROR_TYPE, SOLVER_ERROR_ODE_MODEL_COULD_NOT_BE_CONSTRUCTED, "message0");
  SolverError_error(FATAL_ERROR_TYPE, SOLVER_ERROR_NO_KINETIC_LAW_FOUND_FOR_REACTION, "fatal1");
  SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_ENTRIES_OF_THE_JACOBIAN_MATRIX_COULD_NOT_BE_CONSTRUCTED, "error1");
}
static void test_solver_error(void)
{
  char *s;
  SolverError_clear();
  CK(SolverError_getNum(FATAL_ERROR_TYPE) == 0); CK(SolverError_getNum(ERROR_ERROR_TYPE) == 0);
  CK(SolverError_getNum(WARNING_ERROR_TYPE) == 0); CK(SolverError_getNum(MESSAGE_ERROR_TYPE) == 0);
  SolverError_haltOnErrors();                                /* returns: nothing stored */
  add_some_errors();
  CK(SolverError_getNum(FATAL_ERROR_TYPE) == 2); CK(SolverError_getNum(ERROR_ERROR_TYPE) == 2);
  CK(SolverError_getNum(WARNING_ERROR_TYPE) == 1); CK(SolverError_getNum(MESSAGE_ERROR_TYPE) == 1);
  CK_STR(SolverError_getMessage(FATAL_ERROR_TYPE, 0), "fatal0"); CK_STR(SolverError_getMessage(FATAL_ERROR_TYPE, 1), "fatal1");
  CK_STR(SolverError_getMessage(ERROR_ERROR_TYPE, 0), "error0"); CK_STR(SolverError_getMessage(ERROR_ERROR_TYPE, 1), "error1");
  CK_STR(SolverError_getMessage(WARNING_ERROR_TYPE, 0), "warning0"); CK_STR(SolverError_getMessage(MESSAGE_ERROR_TYPE, 0), "message0");
  CK(SolverError_getCode(FATAL_ERROR_TYPE, 0) == SOLVER_ERROR_ODE_COULD_NOT_BE_CONSTRUCTED_FOR_SPECIES);
  CK(SolverError_getCode(FATAL_ERROR_TYPE, 1) == SOLVER_ERROR_NO_KINETIC_LAW_FOUND_FOR_REACTION);
  CK(SolverError_getCode(ERROR_ERROR_TYPE, 0) == SOLVER_ERROR_THE_MODEL_CONTAINS_EVENTS);
  CK(SolverError_getCode(ERROR_ERROR_TYPE, 1) == SOLVER_ERROR_ENTRIES_OF_THE_JACOBIAN_MATRIX_COULD_NOT_BE_CONSTRUCTED);
  CK(SolverError_getCode(WARNING_ERROR_TYPE, 0) == SOLVER_ERROR_THE_MODEL_CONTAINS_ALGEBRAIC_RULES);
  CK(SolverError_getCode(MESSAGE_ERROR_TYPE, 0) == SOLVER_ERROR_ODE_MODEL_COULD_NOT_BE_CONSTRUCTED);
  CK(SolverError_getLastCode(FATAL_ERROR_TYPE) == SOLVER_ERROR_NO_KINETIC_LAW_FOUND_FOR_REACTION);
  CK(SolverError_getLastCode(ERROR_ERROR_TYPE) == SOLVER_ERROR_ENTRIES_OF_THE_JACOBIAN_MATRIX_COULD_NOT_BE_CONSTRUCTED);
  CK(SolverError_getLastCode(WARNING_ERROR_TYPE) == SOLVER_ERROR_THE_MODEL_CONTAINS_ALGEBRAIC_RULES);
  CK(SolverError_getLastCode(MESSAGE_ERROR_TYPE) == SOLVER_ERROR_ODE_MODEL_COULD_NOT_BE_CONSTRUCTED);
  s = SolverError_dumpToString();
  CK_STR(s, "Fatal Error\t100000\tfatal0\n" "Fatal Error\t100004\tfatal1\n" "      Error\t100001\terror0\n"
            "      Error\t100005\terror1\n" "    Warning\t100002\twarning0\n" "    Message\t100003\tmessage0\n");
  SolverError_freeDumpString(s);
  SolverError_clear();
  CK(SolverError_getNum(FATAL_ERROR_TYPE) == 0); CK(SolverError_getNum(ERROR_ERROR_TYPE) == 0);
  CK(SolverError_getNum(WARNING_ERROR_TYPE) == 0); CK(SolverError_getNum(MESSAGE_ERROR_TYPE) == 0);
}

/* results object as the reference's fixture builds it: containers for MAPK with 3 time points */
static void test_sbml_results(const char *mapk)
{
  SBMLDocument_t *d = readSBML(mapk); Model_t *m; SBMLResults_t *r; timeCourse_t *tc; SBMLResultsArray_t *a;
  CK(d != NULL); if (!d) return;
  m = SBMLDocument_getModel(d);
  r = SBMLResults_create(m, 3);
  CK(SBMLResults_getTime(r) == r->time); CK(SBMLResults_getNout(r) == 3); CK(SBMLResults_getNumSens(r) == 0);
  CK(SBMLResults_getTimeCourse(r, "no_such_name") == NULL);
  CK(r->species->num_val == 8); CK(r->compartments->num_val == 0); CK(r->parameters->num_val == 0); CK(r->fluxes->num_val == 10);
  tc = Species_getTimeCourse(Model_getSpecies(m, 1), r);
  CK(tc != NULL && tc == r->species->tc[1]); CK_STR(TimeCourse_getName(tc), "MKKK_P"); CK(TimeCourse_getNumValues(tc) == 3);
  tc->values[2] = 4.25; CK(TimeCourse_getValue(tc, 2) == 4.25);
  CK(Compartment_getTimeCourse(Model_getCompartment(m, 0), r) == NULL);     /* constant compartment */
  CK(Parameter_getTimeCourse(Model_getParameter(m, 0), r) == NULL);         /* constant parameter */
  tc = SBMLResults_getTimeCourse(r, "J3"); CK(tc != NULL && tc == r->fluxes->tc[3]);
  a = SBMLResultsArray_allocate(2);
  CK(SBMLResultsArray_getNumResults(a) == 2);
  a->results[0] = r; a->results[1] = SBMLResults_create(m, 1);
  CK(SBMLResultsArray_getResults(a, 0) == r); CK(SBMLResultsArray_getResults(a, 2) == NULL);
  SBMLResultsArray_free(a);
  SBMLResults_free(NULL); SBMLResultsArray_free(NULL);
  SBMLDocument_free(d);
}

/* fixture of unittest/test_modelSimplify.c: 1 + (x * x) */
static ASTNode_t *sum_of_one_and_square(void) { ASTNode_t *n = SBML_parseFormula("1 + (x * x)"); CK(n != NULL); return n; }
static void check_one_plus_product(ASTNode_t *node, ASTNode_t **factor0, ASTNode_t **factor1)
{
  ASTNode_t *c;
  CK(ASTNode_getType(node) == AST_PLUS); CK(ASTNode_getNumChildren(node) == 2);
  c = ASTNode_getChild(node, 0);
  CK(ASTNode_getType(c) == AST_INTEGER); CK(ASTNode_getInteger(c) == 1);
  c = ASTNode_getChild(node, 1);
  CK(ASTNode_getType(c) == AST_TIMES); CK(ASTNode_getNumChildren(c) == 2);
  *factor0 = ASTNode_getChild(c, 0); *factor1 = ASTNode_getChild(c, 1);
}
static void check_constants(const char *dir, const char *file, const char *formula, const char *expected)
{
  char path[4096]; SBMLDocument_t *d; ASTNode_t *n; char *s;
  snprintf(path, sizeof path, "%s/%s", dir, file);
  d = readSBML(path);
  CK(d != NULL); if (!d) return;
  n = SBML_parseFormula(formula);
  AST_replaceConstants(SBMLDocument_getModel(d), n);
  s = SBML_formulaToString(n);
  CK_STR(s, expected);
  free(s); ASTNode_free(n); SBMLDocument_free(d);
}
static void test_model_simplify(const char *dir)
{
  ASTNode_t *node, *n, *f[2], *c; int i;
  /* replaceNameByFormula: x -> y/z */
  node = sum_of_one_and_square(); n = SBML_parseFormula("y/z");
  AST_replaceNameByFormula(node, "x", n);
  check_one_plus_product(node, &f[0], &f[1]);
  for (i = 0; i < 2; i++) {
    CK(ASTNode_getType(f[i]) == AST_DIVIDE); CK(ASTNode_getNumChildren(f[i]) == 2);
    c = ASTNode_getChild(f[i], 0); CK(ASTNode_getType(c) == AST_NAME); CK_STR(ASTNode_getName(c), "y");
    c = ASTNode_getChild(f[i], 1); CK(ASTNode_getType(c) == AST_NAME); CK_STR(ASTNode_getName(c), "z");
  }
  ASTNode_free(n); ASTNode_free(node);
  /* replaceNameByName */
  node = sum_of_one_and_square();
  AST_replaceNameByName(node, "x", "y");
  check_one_plus_product(node, &f[0], &f[1]);
  for (i = 0; i < 2; i++) { CK(ASTNode_getType(f[i]) == AST_NAME); CK_STR(ASTNode_getName(f[i]), "y"); }
  ASTNode_free(node);
  /* replaceNameByValue */
  node = sum_of_one_and_square();
  AST_replaceNameByValue(node, "x", 42.0);
  check_one_plus_product(node, &f[0], &f[1]);
  for (i = 0; i < 2; i++) { CK(ASTNode_getType(f[i]) == AST_REAL); CK(ASTNode_getReal(f[i]) == 42.0); }
  ASTNode_free(node);
  /* replaceNameByParameters: a list holding parameter x = 2.5 */
  {
    Model_t *m = Model_create(2, 4); Reaction_t *r = Model_createReaction(m, "r"); KineticLaw_t *kl = Reaction_createKineticLaw(r, "x");
    KineticLaw_createParameter(kl, "x", 2.5);
    node = sum_of_one_and_square();
    AST_replaceNameByParameters(node, &kl->parameters);
    check_one_plus_product(node, &f[0], &f[1]);
    for (i = 0; i < 2; i++) { CK(ASTNode_getType(f[i]) == AST_REAL); CK(ASTNode_getReal(f[i]) == 2.5); }
    ASTNode_free(node); Model_free(m);
  }
  /* replaceFunctionDefinition: f(1, f(x, y)) with f = lambda(a, b, a - b) */
  node = SBML_parseFormula("f(1, f(x, y))"); n = SBML_parseFormula("lambda(a, b, a - b)");
  CK(node != NULL && n != NULL);
  AST_replaceFunctionDefinition(node, "f", n);
  CK(ASTNode_getType(node) == AST_MINUS); CK(ASTNode_getNumChildren(node) == 2);
  c = ASTNode_getChild(node, 0); CK(ASTNode_getType(c) == AST_INTEGER); CK(ASTNode_getInteger(c) == 1);
  c = ASTNode_getChild(node, 1); CK(ASTNode_getType(c) == AST_MINUS); CK(ASTNode_getNumChildren(c) == 2);
  CK(ASTNode_getType(ASTNode_getChild(c, 0)) == AST_NAME); CK_STR(ASTNode_getName(ASTNode_getChild(c, 0)), "x");
  CK(ASTNode_getType(ASTNode_getChild(c, 1)) == AST_NAME); CK_STR(ASTNode_getName(ASTNode_getChild(c, 1)), "y");
  ASTNode_free(n); ASTNode_free(node);
  /* replaceConstants on the example models */
  check_constants(dir, "mapk_cascade.xml", "V1 + Ki + K1", "2.5 + 9 + 10");
  check_constants(dir, "basic_forward.xml", "c + k_1 + k_2", "1 + k_1 + k_2");
  check_constants(dir, "basic_reversible.xml", "c + k_1 + k_2", "1 + 1 + k_2");
}

int main(int argc, char **argv)
{
  test_vary_settings();
  test_solver_error();
  if (argc > 1) {
    char path[4096];
    snprintf(path, sizeof path, "%s/mapk_cascade.xml", argv[1]);
    test_sbml_results(path);
    test_model_simplify(argv[1]);
  }
  if (failures) { fprintf(stderr, "%d check(s) failed\n", failures); return EXIT_FAILURE; }
  printf("ok\n");
  return EXIT_SUCCESS;
}
