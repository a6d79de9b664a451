/*
 * test_integrator_kats.c -- unittest/test_integratorInstance.c:15-150 restated against this library (plain asserts):
 * the single-instance API with the default settings (BDF, Newton, atol 1e-18, rtol 1e-10, t = 0..1 in 10 steps) on
 * MAPK.  Needs a CUDA device: the integration runs on the GPU (batch of one).  Run by tests/test_c_api_kats.py (GPU tier).
 */
#include <math.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include "sbmlsolver/odeSolver.h"
#include "sbmlsolver/solverError.h"

static int failures = 0;
#define CK(cond) do { if (!(cond)) { fprintf(stderr, "%s:%d: failed: %s\n", __FILE__, __LINE__, #cond); failures++; } } while (0)
#define CK_TOL(d, e) CK(fabs((d) - (e)) <= DBL_EPSILON)

int main(int argc, char **argv)
{
  odeModel_t *model; cvodeSettings_t *cs; integratorInstance_t *ii; const cvodeResults_t *cr; cvodeResults_t *copy; FILE *fp; int i;
  if (argc < 2) return EXIT_FAILURE;
  model = ODEModel_createFromFile(argv[1]);
  CK(model != NULL);
  cs = CvodeSettings_create();
  /* create */
  ii = IntegratorInstance_create(model, cs);
  CK(ii != NULL);
  CK(IntegratorInstance_getSettings(ii) == cs);
  CK_TOL(IntegratorInstance_getTime(ii), 0.0);
  CK(IntegratorInstance_timeCourseCompleted(ii) == 0);
  CK(IntegratorInstance_getIntegrationTime(ii) == 0.0);
  /* integrate */
  CK(IntegratorInstance_integrate(ii) == 1);
  CK_TOL(IntegratorInstance_getTime(ii), 1.0);
  CK(IntegratorInstance_timeCourseCompleted(ii) == 1);
  /* results */
  cr = IntegratorInstance_getResults(ii);
  CK(ii->results == cr);
  CK(CvodeResults_getNout(cr) == 11);                       /* number of time points, src/cvodeData.c:208-211 */
  for (i = 0; i <= 10; i++) CK_TOL(CvodeResults_getTime(cr, i), i * 0.1);
  CK(cr->value[0][0] == 90.0 && cr->value[0][10] < 90.0 && cr->value[0][10] > 89.0);      /* MKKK decays slowly from 90 */
  CK(fabs(cr->value[0][10] + cr->value[1][10] - 100.0) < 1e-8);                          /* MKKK + MKKK_P is conserved */
  copy = IntegratorInstance_createResults(ii);
  CK(copy != NULL && copy != ii->results);
  CK(copy->value[3][7] == cr->value[3][7]);
  CvodeResults_free(copy);
  fp = tmpfile();
  CK(fp != NULL);
  IntegratorInstance_printResults(ii, fp);
  CK(ftell(fp) > 100);
  IntegratorInstance_printStatistics(ii, fp);
  fclose(fp);
  /* the value list contains the reaction symbols J0..J9, which are not SBML species, compartments or parameters */
  CK(IntegratorInstance_updateModel(ii) == 0);
  /* a second run after reset reproduces the first bit for bit */
  {
    double last = cr->value[7][10];
    CK(IntegratorInstance_reset(ii) == 1);
    CK_TOL(IntegratorInstance_getTime(ii), 0.0);
    CK(IntegratorInstance_integrate(ii) == 1);
    CK(IntegratorInstance_getResults(ii)->value[7][10] == last);
  }
  IntegratorInstance_free(ii);
  IntegratorInstance_free(NULL);                             /* freeing NULL is safe */
  CvodeSettings_free(cs);
  ODEModel_free(model);
  if (failures || SolverError_getNum(ERROR_ERROR_TYPE) || SolverError_getNum(FATAL_ERROR_TYPE)) { SolverError_dumpAndClearErrors(); fprintf(stderr, "%d check(s) failed\n", failures); return EXIT_FAILURE; }
  printf("ok\n");
  return EXIT_SUCCESS;
}
