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

void CvodeData_free(cvodeData_t *data)
{
  if (!data) return;
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
  for (i = 0; i < r->nvalues; i++) free(r->value[i]);
  free(r->value); free(r->time); free(r);
}
double CvodeResults_getTime(const cvodeResults_t *r, int n) { return r->time[n]; }
double CvodeResults_getValue(cvodeResults_t *r, variableIndex_t *vi, int n) { return r->value[vi->index][n]; }
int CvodeResults_getNout(const cvodeResults_t *r) { return r->nout + 1; }
