/*
 * batch_scan_results.c -- a two-parameter grid scan that consumes the results the way callers of the reference do:
 * one SBMLResults_t per design point out of SBML_odeSolverBatch, printed with SBMLResults_dumpSpecies.
 * Same call sequence as the reference's examples/batch_integrate.c:121-190 (VarySettings_allocate /
 * addParameter / addDesignPoint, SBML_odeSolverBatch, SBMLResultsArray_getResults); underneath, all design points
 * are one batched device run.
 *
 * usage: batch_scan_results model.xml endtime printstep  id1 from1 to1 n1  id2 rid2|- from2 to2 n2
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sbmlsolver/odeSolver.h"
#include "sbmlsolver/solverError.h"

int main(int argc, char **argv)
{
  SBMLDocument_t *d; cvodeSettings_t *set; varySettings_t *vs; SBMLResultsArray_t *all;
  int n1, n2, i, j; double a1, b1, a2, b2, point[2];
  if (argc < 13) { fprintf(stderr, "usage: %s model.xml endtime printstep id1 from1 to1 n1 id2 rid2|- from2 to2 n2\n", argv[0]); return EXIT_FAILURE; }
  a1 = atof(argv[5]); b1 = atof(argv[6]); n1 = atoi(argv[7]);
  a2 = atof(argv[10]); b2 = atof(argv[11]); n2 = atoi(argv[12]);
  d = readSBML(argv[1]);
  if (!d) { fprintf(stderr, "cannot read %s\n", argv[1]); return EXIT_FAILURE; }
  set = CvodeSettings_create();
  CvodeSettings_setTime(set, atof(argv[2]), atoi(argv[3]));
  CvodeSettings_setErrors(set, 1e-9, 1e-6, 10000);
  vs = VarySettings_allocate(2, n1 * n2);
  VarySettings_addParameter(vs, argv[4], NULL);
  VarySettings_addParameter(vs, argv[8], strcmp(argv[9], "-") ? argv[9] : NULL);
  for (i = 0; i < n1; i++)
    for (j = 0; j < n2; j++) {
      point[0] = a1 + i * (b1 - a1) / n1;
      point[1] = a2 + j * (b2 - a2) / n2;
      VarySettings_addDesignPoint(vs, point);
    }
  all = SBML_odeSolverBatch(d, set, vs);
  if (!all) {
    printf("### Parameter variation not succesful!\n");
    SolverError_dumpAndClearErrors();
    CvodeSettings_free(set); SBMLDocument_free(d); VarySettings_free(vs);
    return EXIT_FAILURE;
  }
  for (i = 0; i < SBMLResultsArray_getNumResults(all); i++) {
    printf("### Parameters: %s=%f, %s=%f, \n", VarySettings_getName(vs, 0), VarySettings_getValue(vs, i, 0), VarySettings_getName(vs, 1), VarySettings_getValue(vs, i, 1));
    SBMLResults_dumpSpecies(SBMLResultsArray_getResults(all, i));
    printf("\n");
  }
  SBMLResultsArray_free(all);
  VarySettings_free(vs); CvodeSettings_free(set); SBMLDocument_free(d);
  return EXIT_SUCCESS;
}
