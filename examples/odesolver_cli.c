/*
 * odesolver_cli.c -- the reference's command-line driver for one time course, reduced to the options that touch the
 * integration path, with the reference's text formats.
 *   odeSolver/printModel.c:847-947  "#t <names>" / "##CONCENTRATIONS" blocks (values printed %g: the format of the
 *                                   golden files examples/MAPK_{10,100,1000}pt.dat)
 *   odeSolver/printModel.c:674-720  "##ODE VALUES"      (-y: right-hand sides at the stored points)
 *   odeSolver/printModel.c:737-800  "##REACTION RATES"  (-k: kinetic laws at the stored points)
 *   odeSolver/options.c:91-98       defaults: --error 1e-9, --rerror 1e-4, --mxstep 10000, --time 1, --printstep 50
 * The time course is computed by SBML_odeSolver-style calls of the kept API (IntegratorInstance_*), i.e. on the device.
 *
 * usage: odesolver_cli model.xml [--time T] [--printstep N] [--error A] [--rerror R] [--mxstep M] [-a] [-y] [-k] [-n]
 *   -a print all values (variables, assigned values, constants) instead of the ODE variables; -n do not halt on events
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sbmlsolver/odeSolver.h"
#include "sbmlsolver/processAST.h"
#include "sbmlsolver/solverError.h"

static void header(FILE *f, char **names, int n, const char *block, int top)
{
  int i;
  if (!top) fprintf(f, "##%s\n", block);
  fprintf(f, "#t ");
  for (i = 0; i < n; i++) fprintf(f, "%s ", names[i]);
  fprintf(f, "\n");
  if (top) fprintf(f, "##%s\n", block);
}

int main(int argc, char **argv)
{
  double time = 1.0, aerr = 1e-9, rerr = 1e-4; int printstep = 50, mxstep = 10000, all = 0, rates = 0, reactions = 0, halt = 1, i, j, k;
  const char *file = NULL; SBMLDocument_t *d; Model_t *m; odeModel_t *om; cvodeSettings_t *set; integratorInstance_t *ii;
  const cvodeResults_t *res; cvodeData_t *data; int nvalues, ncols;
  for (i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--time") && i + 1 < argc) time = atof(argv[++i]);
    else if (!strcmp(argv[i], "--printstep") && i + 1 < argc) printstep = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--error") && i + 1 < argc) aerr = atof(argv[++i]);
    else if (!strcmp(argv[i], "--rerror") && i + 1 < argc) rerr = atof(argv[++i]);
    else if (!strcmp(argv[i], "--mxstep") && i + 1 < argc) mxstep = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-a")) all = 1;
    else if (!strcmp(argv[i], "-y")) rates = 1;
    else if (!strcmp(argv[i], "-k")) reactions = 1;
    else if (!strcmp(argv[i], "-n")) halt = 0;
    else if (argv[i][0] != '-') file = argv[i];
    else { fprintf(stderr, "unknown option %s\n", argv[i]); return EXIT_FAILURE; }
  }
  if (!file) { fprintf(stderr, "usage: %s model.xml [--time T] [--printstep N] [--error A] [--rerror R] [--mxstep M] [-a] [-y] [-k] [-n]\n", argv[0]); return EXIT_FAILURE; }
  d = readSBML(file);
  if (!d || !(m = SBMLDocument_getModel(d))) { fprintf(stderr, "cannot read %s\n", file); return EXIT_FAILURE; }
  om = ODEModel_create(m);
  if (!om) { SolverError_dumpAndClearErrors(); return EXIT_FAILURE; }
  set = CvodeSettings_createWith(time, printstep, aerr, rerr, mxstep, 0, 0, 1, 0, halt, 0, 1, 0, 0);
  ii = IntegratorInstance_create(om, set);
  if (!ii) { SolverError_dumpAndClearErrors(); return EXIT_FAILURE; }
  if (!IntegratorInstance_integrate(ii)) { SolverError_dumpAndClearErrors(); fprintf(stderr, "Integration not sucessful!\n"); }
  res = IntegratorInstance_getResults(ii);
  data = IntegratorInstance_getData(ii);
  nvalues = ODEModel_getNumValues(om);
  if (rates) {
    /* right-hand sides evaluated on the stored values of every output time */
    header(stdout, om->names, om->neq, "ODE VALUES", 1);
    for (i = 0; i <= res->nout; i++) {
      printf("%g ", res->time[i]);
      data->currenttime = (float)res->time[i];
      for (j = 0; j < nvalues; j++) data->value[j] = res->value[j][i];
      for (j = 0; j < om->neq; j++) printf("%g ", evaluateAST(om->ode[j], data));
      printf("\n");
    }
    header(stdout, om->names, om->neq, "ODE VALUES", 0);
  } else if (reactions) {
    /* kinetic laws with their local parameters substituted, evaluated on the stored values */
    unsigned int nr = Model_getNumReactions(m), r;
    ASTNode_t **kls = (ASTNode_t **)calloc(nr ? nr : 1, sizeof *kls);
    char **ids = (char **)calloc(nr ? nr : 1, sizeof *ids);
    for (r = 0; r < nr; r++) {
      Reaction_t *rx = Model_getReaction(m, r);
      ids[r] = rx->id;
      kls[r] = copyAST(rx->kineticLaw->math);
      AST_replaceNameByParameters(kls[r], &rx->kineticLaw->parameters);
      AST_replaceConstants(m, kls[r]);
    }
    header(stdout, ids, (int)nr, "REACTION RATES", 1);
    for (k = 0; k <= res->nout; k++) {
      printf("%g ", res->time[k]);
      data->currenttime = (float)res->time[k];
      for (j = 0; j < nvalues; j++) data->value[j] = res->value[j][k];
      for (r = 0; r < nr; r++) printf("%g ", evaluateAST(kls[r], data));
      printf("\n");
    }
    header(stdout, ids, (int)nr, "REACTION RATES", 0);
    for (r = 0; r < nr; r++) ASTNode_free(kls[r]);
    free(kls); free(ids);
  } else {
    ncols = all ? nvalues : om->neq;
    header(stdout, om->names, ncols, "CONCENTRATIONS", 1);
    for (i = 0; i <= res->nout; i++) {
      printf("%g ", res->time[i]);
      for (j = 0; j < ncols; j++) printf("%g ", res->value[j][i]);
      printf("\n");
    }
    header(stdout, om->names, ncols, "CONCENTRATIONS", 0);
  }
  printf("\n\n");
  IntegratorInstance_free(ii);
  CvodeSettings_free(set);
  om->m = NULL; ODEModel_free(om);
  SBMLDocument_free(d);
  return SolverError_getNum(ERROR_ERROR_TYPE) || SolverError_getNum(FATAL_ERROR_TYPE) ? EXIT_FAILURE : EXIT_SUCCESS;
}
