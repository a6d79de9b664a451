/*
 * test_sens_kats.c -- the reference's sensitivity unit tests restated as a plain C99 caller of libodes_b200.so:
 *   unittest/test_cvodeData.c:131-215   CvodeResults_allocateSens / getSensitivityByNum / getSensitivity / computeDirectional
 *   unittest/test_sensSolver.c:14-31    the fixture: MAPK.xml, sensitivity switched on, simultaneous method, parameters
 *                                       MAPK, MAPK_P, K1, Ki -- stepped with IntegratorInstance_cvodeOneStep until the
 *                                       time course is complete (:56-62)
 * usage: test_sens_kats MAPK.xml [cpu]   ("cpu": only the parts that need no device)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sbmlsolver/integratorInstance.h"
#include "sbmlsolver/odeSolver.h"
#include "sbmlsolver/solverError.h"

static int failures = 0;
#define CK(cond) do { if (!(cond)) { fprintf(stderr, "%s:%d: failed: %s\n", __FILE__, __LINE__, #cond); failures++; } } while (0)

static const char *params[] = { "MAPK", "MAPK_P", "K1", "Ki" };

static void results_kats(odeModel_t *model)
{
  cvodeData_t *cd = CvodeData_create(model);
  cvodeResults_t *cr = CvodeResults_create(cd, 3);
  variableIndex_t y, s;
  static const double dp[] = { 1.0, 0.5, 2.0 };
  int i, j, k, n = 1;
  cr->nout = 3;
  CvodeResults_allocateSens(cr, 2, 3, cr->nout);
  cr->sensitivity[1][2][3] = .7;
  CK(CvodeResults_getSensitivityByNum(cr, 1, 2, 3) == .7);
  CK(CvodeResults_getSensitivityByNum(cr, 2, 2, 3) == 0);
  CK(CvodeResults_getSensitivityByNum(cr, 1, 3, 3) == 0);
  CK(CvodeResults_getSensitivityByNum(cr, 1, 2, 4) == 0);
  CvodeResults_allocateSens(cr, 2, 3, cr->nout);
  cr->index_sens[0] = 0; cr->index_sens[1] = 2; cr->index_sens[2] = 17;
  cr->sensitivity[0][1][3] = .7; cr->sensitivity[1][2][3] = .8;
  memset(&y, 0, sizeof y); memset(&s, 0, sizeof s);
  y.index = 0; s.index = 16;
  CK(CvodeResults_getSensitivity(cr, &y, &s, 3) == 0);
  y.index = 1; s.index = 17;
  CK(CvodeResults_getSensitivity(cr, &y, &s, 3) == .8);
  CvodeResults_allocateSens(cr, 2, 3, cr->nout);
  for (i = 0; i < 2; i++) for (j = 0; j < 3; j++) for (k = 0; k < 4; k++) cr->sensitivity[i][j][k] = n++;
  CvodeResults_computeDirectional(cr, dp);
  CK(cr->directional[0][0] == 21.5); CK(cr->directional[0][1] == 25); CK(cr->directional[0][2] == 28.5); CK(cr->directional[0][3] == 32);
  CK(cr->directional[1][0] == 63.5); CK(cr->directional[1][1] == 67); CK(cr->directional[1][2] == 70.5); CK(cr->directional[1][3] == 74);
  CvodeResults_free(cr);
  CvodeData_free(cd);
}

int main(int argc, char **argv)
{
  odeModel_t *model; cvodeSettings_t *cs; integratorInstance_t *ii; const cvodeResults_t *cr; SBMLDocument_t *d; SBMLResults_t *sr;
  variableIndex_t *vy, *vp; timeCourse_t *tc; int k;
  if (argc < 2) return EXIT_FAILURE;
  model = ODEModel_createFromFile(argv[1]);
  CK(model != NULL);
  results_kats(model);
  /* settings round trip */
  cs = CvodeSettings_create();
  CvodeSettings_setSensitivity(cs, 1);
  CvodeSettings_setSensMethod(cs, 0);
  CvodeSettings_setSensParams(cs, (char **)params, 4);
  CK(CvodeSettings_getSensitivity(cs) == 1 && cs->nsens == 4 && strcmp(cs->sensIDs[2], "K1") == 0);
  CK(strcmp(CvodeSettings_getSensMethod(cs), "simultaneous") == 0);
  if (argc > 2 && strcmp(argv[2], "cpu") == 0) goto done;

  /* the fixture of unittest/test_sensSolver.c */
  ii = IntegratorInstance_create(model, cs);
  CK(ii != NULL);
  if (!ii) { SolverError_dump(); return EXIT_FAILURE; }
  CK(IntegratorInstance_getNsens(ii) == 4);
  CK(IntegratorInstance_getSensitivityModel(ii) != NULL);
  CK(strcmp(IntegratorInstance_getSensVariableName(ii, 3), "Ki") == 0);
  /* initial sensitivities: d MAPK / d MAPK(0) = 1, everything else 0 (src/cvodeData.c:453-461) */
  CK(IntegratorInstance_getSensitivityByNum(ii, 5, 0) == 1.0 && IntegratorInstance_getSensitivityByNum(ii, 6, 1) == 1.0);
  CK(IntegratorInstance_getSensitivityByNum(ii, 5, 1) == 0.0 && IntegratorInstance_getSensitivityByNum(ii, 0, 2) == 0.0);
  while (!IntegratorInstance_timeCourseCompleted(ii))
    if (!IntegratorInstance_cvodeOneStep(ii)) { fprintf(stderr, "failed to advance one step\n"); SolverError_dump(); failures++; break; }
  cr = IntegratorInstance_getResults(ii);
  CK(cr->nsens == 4 && cr->neq == 8 && cr->sensitivity != NULL);
  vy = ODEModel_getVariableIndex(model, "MKKK"); vp = ODEModel_getVariableIndex(model, "K1");
  CK(IntegratorInstance_getSensitivity(ii, vy, vp) == CvodeResults_getSensitivity((cvodeResults_t *)cr, vy, vp, cr->nout));
  CK(IntegratorInstance_getSensitivity(ii, vy, vp) != 0.0);
  /* MKKK + MKKK_P is conserved: the pair's sensitivities cancel for every parameter at every stored time */
  for (k = 0; k <= cr->nout; k++) {
    CK(fabs(cr->sensitivity[0][2][k] + cr->sensitivity[1][2][k]) <= 1e-9 * (1.0 + fabs(cr->sensitivity[0][2][k])));
    CK(fabs(cr->sensitivity[0][3][k] + cr->sensitivity[1][3][k]) <= 1e-9 * (1.0 + fabs(cr->sensitivity[0][3][k])));
  }
  VariableIndex_free(vy); VariableIndex_free(vp);
  IntegratorInstance_free(ii);

  /* the same through SBML_odeSolver: SBMLResults_createSens (src/odeSolver.c:545-575) */
  d = parseModel(argv[1], 0, 0);
  sr = SBML_odeSolver(d, cs);
  CK(sr != NULL);
  if (sr) {
    CK(SBMLResults_getNumSens(sr) == 4 && strcmp(SBMLResults_getSensParam(sr, 2), "K1") == 0);
    tc = SBMLResults_getTimeCourse(sr, "MAPK");
    CK(tc != NULL && tc->sensitivity != NULL && TimeCourse_getSensitivity(tc, 0, 0) == 1.0);
    tc = SBMLResults_getTimeCourse(sr, "MKKK");
    CK(tc != NULL && TimeCourse_getSensitivity(tc, 2, TimeCourse_getNumValues(tc) - 1) != 0.0);
    SBMLResults_free(sr);
  }
  SBMLDocument_free(d);
done:
  CvodeSettings_free(cs);
  ODEModel_free(model);
  if (failures) { fprintf(stderr, "%d check(s) failed\n", failures); return EXIT_FAILURE; }
  printf("ok\n");
  return EXIT_SUCCESS;
}
