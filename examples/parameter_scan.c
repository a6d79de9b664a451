/*
 * parameter_scan.c -- the reference's examples/ParameterScanner.c / batch_integrate.c use case:
 *     parameter_scan <model.xml> <end time> <print steps> <parameter id> <reaction id or -> <start> <end> <points> [species id]
 * scans one (global, or kinetic-law local) parameter over `points` values and prints, per design point, the value
 * of every ODE variable (or of one species) at the end time.  Where the reference loops
 *     for each value: IntegratorInstance_setVariableValue; while (!timeCourseCompleted) integrateOneStep; reset
 * (examples/ParameterScanner.c:136-163, src/odeSolver.c:313-330) this is ONE call of Model_odeSolverBatchFlat (or Model_odeSolverBatch for per-point SBMLResults_t):
 * all design points are integrated by one kernel launch.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <odes_b200.h>

int main(int argc, char *argv[])
{
  SBMLDocument_t *d; cvodeSettings_t *set; varySettings_t *vs; batchResults_t *res;
  const char *par, *rid, *only; double end, lo, hi; int steps, points, i, k, failed = 0;
  if (argc < 9) { fprintf(stderr, "usage %s sbml-model-file time steps parameter reaction|- start end points [species]\n", argv[0]); return EXIT_FAILURE; }
  end = atof(argv[2]); steps = atoi(argv[3]); par = argv[4]; rid = strcmp(argv[5], "-") ? argv[5] : NULL;
  lo = atof(argv[6]); hi = atof(argv[7]); points = atoi(argv[8]); only = argc > 9 ? argv[9] : NULL;
  d = readSBML(argv[1]);
  if (!d) { fprintf(stderr, "cannot read %s\n", argv[1]); return EXIT_FAILURE; }
  set = CvodeSettings_createWith(end, steps, 1e-9, 1e-6, 10000, 0, 0, 1, 0, 0, 0, 1, 0, 0);
  vs = VarySettings_allocate(1, points);
  VarySettings_addParameter(vs, par, rid);
  for (i = 0; i < points; i++) { double v = points > 1 ? lo + (hi - lo) * i / (points - 1) : lo; VarySettings_addDesignPoint(vs, &v); }
  res = SBML_odeSolverBatchFlat(d, set, vs);
  if (!res) { SolverError_dumpAndClearErrors(); return EXIT_FAILURE; }
  printf("#%s", par);
  for (k = 0; k < res->nvalues; k++) if (only ? !strcmp(only, res->names[k]) : 1) printf(" %s", res->names[k]);
  printf(" status\n");
  for (i = 0; i < res->nsets; i++) {
    printf("%g", VarySettings_getValue(vs, i, 0));
    for (k = 0; k < res->nvalues; k++) if (only ? !strcmp(only, res->names[k]) : 1) printf(" %.10g", BatchResults_getValue(res, i, res->names[k], res->nout - 1));
    printf(" %d\n", res->status[i]);
    if (res->status[i] < 0) failed++;
  }
  BatchResults_free(res);
  VarySettings_free(vs);
  CvodeSettings_free(set);
  SBMLDocument_free(d);
  return failed ? EXIT_FAILURE : EXIT_SUCCESS;
}
