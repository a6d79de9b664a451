/*
 * integrate_timecourse.c -- the reference's examples/integrate.c use case through the kept C API:
 *     integrate_timecourse <model.xml> <end time> <print steps> [atol rtol mxstep]
 * reads an SBML model, integrates it once over [0, end] with BDF/Newton and prints the species time course in the
 * "#time id id ..." / %g format of the reference's result dumps (examples/testrun.out:27-129 is this program run on
 * MAPK.xml with 1000 100).  The stepping runs on the GPU (a batch of one).
 * Build: make -C examples      Run: examples/_build/integrate_timecourse tests/golden/models/mapk_cascade.xml 1000 100
 */
#include <stdio.h>
#include <stdlib.h>
#include <odes_b200.h>

int main(int argc, char *argv[])
{
  odeModel_t *om; cvodeSettings_t *set; integratorInstance_t *ii; const cvodeResults_t *res;
  double end, atol = 1e-9, rtol = 1e-4; int steps, mxstep = 1000, i, k, neq;
  if (argc < 4) { fprintf(stderr, "usage %s sbml-model-file simulation-time time-steps [atol rtol mxstep]\n", argv[0]); return EXIT_FAILURE; }
  end = atof(argv[2]); steps = atoi(argv[3]);
  if (argc > 4) atol = atof(argv[4]);
  if (argc > 5) rtol = atof(argv[5]);
  if (argc > 6) mxstep = atoi(argv[6]);
  om = ODEModel_createFromFile(argv[1]);
  if (!om) { SolverError_dumpAndClearErrors(); return EXIT_FAILURE; }
  set = CvodeSettings_createWith(end, steps, atol, rtol, mxstep, 0, 0, 1, 0, 0, 0, 1, 0, 0);
  ii = IntegratorInstance_create(om, set);
  if (!ii || !IntegratorInstance_integrate(ii)) {
    fprintf(stderr, "Integration not sucessful!\n");
    SolverError_dumpAndClearErrors();
    return EXIT_FAILURE;
  }
  res = IntegratorInstance_getResults(ii);
  neq = ODEModel_getNeq(om);
  printf("#time");
  for (k = 0; k < neq; k++) { variableIndex_t *vi = ODEModel_getVariableIndexByNum(om, k); printf(" %s", ODEModel_getVariableName(om, vi)); VariableIndex_free(vi); }
  printf("\n");
  for (i = 0; i < CvodeResults_getNout(res); i++) {
    printf("%g", CvodeResults_getTime(res, i));
    for (k = 0; k < neq; k++) { variableIndex_t *vi = ODEModel_getVariableIndexByNum(om, k); printf(" %g", CvodeResults_getValue((cvodeResults_t *)res, vi, i)); VariableIndex_free(vi); }
    printf("\n");
  }
  IntegratorInstance_free(ii);
  CvodeSettings_free(set);
  ODEModel_free(om);
  return EXIT_SUCCESS;
}
