/*
 * cvode_data.c -- cvodeData_t / cvodeResults_t: current values, trigger flags
 * and stored time courses of ONE integrator instance.
 * Reference: src/cvodeData.c:133-190 (create, initializeValues),
 * :327-402 (initialize), :596-622 (results), :208-231 (accessors).
 */
#include <stdlib.h>
#include <string.h>
#include "sbmlsolver/cvodeData.h"
#include "sbmlsolver/processAST.h"
#include "sbmlsolver/solverError.h"

static void free_result_sens(cvodeResults_t *r);
cvodeData_t *CvodeData_create(odeModel_t *om)
{
  cvodeData_t *data = (cvodeData_t *)calloc(1, sizeof *data);
  int nvalues = om->neq + om->nconst + om->nass;
  if (!data) return NULL;
  data->nvalues = nvalues; data->neq = om->neq; data->nevents = om->nevents;
  data->value = (double *)calloc((size_t)(nvalues > 0 ? nvalues : 1), sizeof(double));
  data->trigger = (int *)calloc((size_t)(om->nevents > 0 ? om->nevents : 1), sizeof(int));
  data->model = om;
  data->allRulesUpdated = 0;
  return data;
}

/* model values, then the complete ordered rule set: initial assignments and assignment rules */
void CvodeData_initializeValues(cvodeData_t *data)
{
  int i; odeModel_t *om = data->model;
  for (i = 0; i < data->nvalues; i++) data->value[i] = om->values[i];
  data->currenttime = 0.0;
  for (i = 0; i < om->nass + om->ninitAss; i++) {
    nonzeroElem_t *ordered = om->initAssignmentOrder[i];
    int idx = ordered->i == -1 ? ordered->j : ordered->i;
    data->value[idx] = evaluateAST(ordered->ij, data);
  }
  data->allRulesUpdated = 1;
}

int CvodeData_initialize(cvodeData_t *data, cvodeSettings_t *opt, odeModel_t *om, int keepValues)
{
  int i;
  data->opt = opt;
  if (!keepValues) CvodeData_initializeValues(data);
  data->steadystate = 0;
  data->currenttime = opt->Indefinitely ? 0 : (float)opt->TimePoints[0];
  if (data->currenttime != 0.0) {
    for (i = 0; i < om->nass; i++) { nonzeroElem_t *o = om->assignmentOrder[i]; data->value[o->i] = evaluateAST(o->ij, data); }
    data->allRulesUpdated = 1;
  }
  for (i = 0; i < om->neq; i++) evaluateAST(om->ode[i], data);          /* dry run: all symbols resolvable? */
  for (i = 0; i < data->nevents; i++) data->trigger[i] = (int)evaluateAST(om->event[i], data);
  if (data->results) { CvodeResults_free(data->results); data->results = NULL; }
  opt->StoreResults = !opt->Indefinitely && opt->StoreResults;
  if (opt->StoreResults) {
    data->results = CvodeResults_create(data, opt->PrintStep);
    if (!data->results) return 0;
  }
  return 1;
}

static void free_current_sens(cvodeData_t *data)
{
  int i;
  if (data->sensitivity) { for (i = 0; i < data->neq; i++) free(data->sensitivity[i]); free(data->sensitivity); }
  data->sensitivity = NULL; data->nsens = 0;
}
/* src/cvodeData.c:405-480: current sensitivities 0, or 1 for a variable with respect to its own initial value; row 0 of
   the stored time courses */
int CvodeData_initializeSensitivities(cvodeData_t *data, cvodeSettings_t *opt, odeModel_t *om, odeSense_t *os)
{
  int i, j;
  free_current_sens(data);
  data->os = os;
  if (!os) return 1;
  data->nsens = os->nsens;
  data->sensitivity = (double **)calloc((size_t)(om->neq > 0 ? om->neq : 1), sizeof(double *));
  for (i = 0; i < om->neq; i++) {
    data->sensitivity[i] = (double *)calloc((size_t)(os->nsens > 0 ? os->nsens : 1), sizeof(double));
    for (j = 0; j < os->nsens; j++) data->sensitivity[i][j] = (os->index_sensP[j] == -1 && os->index_sens[j] == i) ? 1.0 : 0.0;
  }
  if (data->results && opt->StoreResults) {
    if (!CvodeResults_allocateSens(data->results, om->neq, os->nsens, opt->PrintStep)) return 0;
    for (j = 0; j < os->nsens; j++) {
      data->results->index_sens[j] = os->index_sens[j];
      for (i = 0; i < om->neq; i++) data->results->sensitivity[i][j][0] = data->sensitivity[i][j];
    }
  }
  return 1;
}

void CvodeData_free(cvodeData_t *data)
{
  if (!data) return;
  free_current_sens(data);
  free(data->value); free(data->trigger);
  CvodeResults_free(data->results);
  free(data);
}

cvodeResults_t *CvodeResults_create(cvodeData_t *data, int nout)
{
  int i; cvodeResults_t *r = (cvodeResults_t *)calloc(1, sizeof *r);
  if (!r) return NULL;
  r->time = (double *)calloc((size_t)nout + 1, sizeof(double));
  r->value = (double **)calloc((size_t)(data->nvalues > 0 ? data->nvalues : 1), sizeof(double *));
  r->nvalues = data->nvalues;
  for (i = 0; i < data->nvalues; i++) r->value[i] = (double *)calloc((size_t)nout + 1, sizeof(double));
  r->nout = 0;
  return r;
}
void CvodeResults_free(cvodeResults_t *r)
{
  int i;
  if (!r) return;
  free_result_sens(r);
  for (i = 0; i < r->nvalues; i++) free(r->value[i]);
  free(r->value); free(r->time); free(r);
}
double CvodeResults_getTime(const cvodeResults_t *r, int n) { return r->time[n]; }
double CvodeResults_getValue(cvodeResults_t *r, variableIndex_t *vi, int n) { return r->value[vi->index][n]; }
int CvodeResults_getNout(const cvodeResults_t *r) { return r->nout + 1; }

static void free_result_sens(cvodeResults_t *r)
{
  int i, j;
  if (r->sensitivity) {
    for (i = 0; i < r->neq; i++) { for (j = 0; j < r->nsens; j++) free(r->sensitivity[i][j]); free(r->sensitivity[i]); }
    free(r->sensitivity); free(r->index_sens);
    if (r->directional) { for (i = 0; i < r->neq; i++) free(r->directional[i]); free(r->directional); }
  }
  r->sensitivity = NULL; r->index_sens = NULL; r->directional = NULL; r->nsens = 0;
}
/* src/cvodeData.c:555-590 */
int CvodeResults_allocateSens(cvodeResults_t *r, int neq, int nsens, int nout)
{
  int i, j;
  free_result_sens(r);
  r->index_sens = (int *)calloc((size_t)(nsens > 0 ? nsens : 1), sizeof(int));
  r->sensitivity = (double ***)calloc((size_t)(neq > 0 ? neq : 1), sizeof(double **));
  for (i = 0; i < neq; i++) {
    r->sensitivity[i] = (double **)calloc((size_t)(nsens > 0 ? nsens : 1), sizeof(double *));
    for (j = 0; j < nsens; j++) r->sensitivity[i][j] = (double *)calloc((size_t)(nout + 1), sizeof(double));
  }
  r->directional = (double **)calloc((size_t)(neq > 0 ? neq : 1), sizeof(double *));
  for (i = 0; i < neq; i++) r->directional[i] = (double *)calloc((size_t)(nout + 1), sizeof(double));
  r->nsens = nsens; r->neq = neq;
  return 1;
}
/* src/cvodeData.c:275-287 */
void CvodeResults_computeDirectional(cvodeResults_t *r, const double *dp)
{
  int i, j, k;
  for (i = 0; i < r->neq; i++)
    for (j = 0; j < r->nout + 1; j++) {
      r->directional[i][j] = 0;
      for (k = 0; k < r->nsens; k++) r->directional[i][j] += r->sensitivity[i][k][j] * dp[k];
    }
}
/* src/cvodeData.c:233-262: 0 outside the stored ranges */
double CvodeResults_getSensitivityByNum(cvodeResults_t *r, int value, int parameter, int timestep)
{
  if (!r || !r->sensitivity || value < 0 || value >= r->neq || parameter < 0 || parameter >= r->nsens || timestep < 0 || timestep > r->nout) return 0;
  return r->sensitivity[value][parameter][timestep];
}
double CvodeResults_getSensitivity(cvodeResults_t *r, variableIndex_t *y, variableIndex_t *p, int timestep)
{
  int i;
  if (!r || !r->sensitivity || !y || !p) return 0;
  for (i = 0; i < r->nsens; i++) if (r->index_sens[i] == p->index) return CvodeResults_getSensitivityByNum(r, y->index, i, timestep);
  return 0;
}
