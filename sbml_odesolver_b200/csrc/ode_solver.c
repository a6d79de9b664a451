/*
 * ode_solver.c -- varySettings_t and the high-level batch call.
 * Reference: src/odeSolver.c:235-352 (Model_odeSolverBatch), :354-440
 * (globalizeParameter / localizeParameter), :570-860 (VarySettings_*).
 * The design-point loop of the reference is replaced by one batched device
 * run; results come back as one batchResults_t in batch-fastest layout.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sbmlsolver/odeSolver.h"
#include "sbmlsolver/processAST.h"
#include "sbmlsolver/solverError.h"

static char *sdup(const char *s) { char *r; if (!s) return NULL; r = (char *)malloc(strlen(s) + 1); strcpy(r, s); return r; }

varySettings_t *VarySettings_allocate(int nrparams, int nrdesignpoints)
{
  int i; varySettings_t *vs = (varySettings_t *)calloc(1, sizeof *vs);
  vs->id = (char **)calloc((size_t)(nrparams > 0 ? nrparams : 1), sizeof(char *));
  vs->rid = (char **)calloc((size_t)(nrparams > 0 ? nrparams : 1), sizeof(char *));
  vs->params = (double **)calloc((size_t)(nrdesignpoints > 0 ? nrdesignpoints : 1), sizeof(double *));
  for (i = 0; i < nrdesignpoints; i++) vs->params[i] = (double *)calloc((size_t)(nrparams > 0 ? nrparams : 1), sizeof(double));
  vs->nrdesignpoints = nrdesignpoints; vs->nrparams = nrparams;
  return vs;
}
int VarySettings_setName(varySettings_t *vs, int i, const char *id, const char *rid)
{
  if (i >= vs->nrparams) { SolverError_error(WARNING_ERROR_TYPE, SOLVER_ERROR_VARY_SETTINGS, "VarySettings_setName:\tRequested Value %d not found in varySettings # parameters: %d", i, vs->nrparams); return 0; }
  free(vs->id[i]); free(vs->rid[i]);
  vs->id[i] = sdup(id);
  vs->rid[i] = (rid && strlen(rid) > 0) ? sdup(rid) : NULL;
  return 1;
}
int VarySettings_addParameter(varySettings_t *vs, const char *id, const char *rid)
{
  if (vs->cnt_params >= vs->nrparams) { SolverError_error(WARNING_ERROR_TYPE, SOLVER_ERROR_VARY_SETTINGS, "VarySettings_addParameter:\tAllocated parameter array already full, #%d parameters", vs->cnt_params); return 0; }
  VarySettings_setName(vs, vs->cnt_params, id, rid);
  return ++vs->cnt_params;
}
int VarySettings_addDesignPoint(varySettings_t *vs, const double *params)
{
  int i;
  if (vs->cnt_points >= vs->nrdesignpoints) { SolverError_error(WARNING_ERROR_TYPE, SOLVER_ERROR_VARY_SETTINGS, "VarySettings_addDesignPoint:\tAllocated design point array already full, #%d design points", vs->cnt_points); return 0; }
  for (i = 0; i < vs->nrparams; i++) vs->params[vs->cnt_points][i] = params[i];
  return ++vs->cnt_points;
}
const char *VarySettings_getName(const varySettings_t *vs, int i) { return i < vs->nrparams ? vs->id[i] : NULL; }
const char *VarySettings_getReactionName(const varySettings_t *vs, int i) { return i < vs->nrparams ? vs->rid[i] : NULL; }
int VarySettings_setValue(varySettings_t *vs, int i, int j, double value)
{ if (i >= vs->nrdesignpoints || j >= vs->nrparams) return 0; vs->params[i][j] = value; return 1; }
double VarySettings_getValue(varySettings_t *vs, int i, int j) { return (i < vs->nrdesignpoints && j < vs->nrparams) ? vs->params[i][j] : 0; }
static int find_param(varySettings_t *vs, const char *id, const char *rid)
{
  int j;
  for (j = 0; j < vs->nrparams; j++)
    if (vs->id[j] && strcmp(id, vs->id[j]) == 0 && ((!rid && !vs->rid[j]) || (rid && vs->rid[j] && strcmp(rid, vs->rid[j]) == 0))) return j;
  return -1;
}
int VarySettings_setValueByID(varySettings_t *vs, int i, const char *id, const char *rid, double value)
{ int j = find_param(vs, id, rid); return j < 0 ? 0 : VarySettings_setValue(vs, i, j, value); }
double VarySettings_getValueByID(varySettings_t *vs, int i, const char *id, const char *rid)
{ int j = find_param(vs, id, rid); return j < 0 ? 0 : VarySettings_getValue(vs, i, j); }
void VarySettings_dump(varySettings_t *vs)
{
  int i, j;
  printf("\nRequested Parameter Variation: Design Points\n#Nr.\t");
  for (j = 0; j < vs->nrparams; j++) printf("%s\t", vs->id[j]);
  printf("\n");
  for (i = 0; i < vs->nrdesignpoints; i++) { printf("#%d:\t", i); for (j = 0; j < vs->nrparams; j++) printf("%.3f\t", vs->params[i][j]); printf("\n"); }
}
void VarySettings_free(varySettings_t *vs)
{
  int i;
  if (!vs) return;
  for (i = 0; i < vs->nrparams; i++) { free(vs->id[i]); free(vs->rid[i]); }
  free(vs->id); free(vs->rid);
  for (i = 0; i < vs->nrdesignpoints; i++) free(vs->params[i]);
  free(vs->params); free(vs);
}

/* local kinetic-law parameter -> global parameter r_<rid>_<id> (src/odeSolver.c:354-391) */
int Model_globalizeParameter(Model_t *m, const char *id, const char *rid)
{
  unsigned int i; int found = 0; Reaction_t *r = Model_getReactionById(m, rid); char *newname;
  if (!r || !r->kineticLaw) return 0;
  newname = (char *)malloc(strlen(id) + strlen(rid) + 4);
  sprintf(newname, "r_%s_%s", rid, id);
  AST_replaceNameByName(r->kineticLaw->math, id, newname);
  for (i = 0; i < r->kineticLaw->parameters.size; i++) {
    Parameter_t *p = (Parameter_t *)r->kineticLaw->parameters.items[i];
    if (strcmp(p->id, id) == 0) { Parameter_t *g = Parameter_clone(p); Parameter_setId(g, newname); ListOf_append(&m->parameters, g); found = 1; }
  }
  free(newname);
  return found;
}
/* undo: the kinetic law reads the local name again; the global copy stays (as in the reference) */
int Model_localizeParameter(Model_t *m, const char *id, const char *rid)
{
  Reaction_t *r = Model_getReactionById(m, rid); char *newname;
  if (!r || !r->kineticLaw) return 0;
  newname = (char *)malloc(strlen(id) + strlen(rid) + 4);
  sprintf(newname, "r_%s_%s", rid, id);
  AST_replaceNameByName(r->kineticLaw->math, newname, id);
  free(newname);
  return 1;
}

void BatchResults_free(batchResults_t *r)
{
  int i;
  if (!r) return;
  for (i = 0; i < r->nvalues; i++) free(r->names[i]);
  for (i = 0; i < r->nsens; i++) free(r->sensParam[i]);
  free(r->sensParam); free(r->sens);
  for (i = 0; i < r->nflux; i++) free(r->fluxNames[i]);
  free(r->fluxNames); free(r->flux);
  free(r->names); free(r->time); free(r->value); free(r->status); free(r);
}
double BatchResults_getValue(const batchResults_t *r, int set, const char *name, int timestep)
{
  int i;
  for (i = 0; i < r->nvalues; i++)
    if (strcmp(r->names[i], name) == 0) return r->value[((size_t)set * (size_t)r->nout + (size_t)timestep) * (size_t)r->nvalues + (size_t)i];
  SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_SYMBOL_IS_NOT_IN_MODEL, "Symbol %s is not in the model", name);
  return 0.0;
}

double BatchResults_getFlux(const batchResults_t *r, int set, const char *reaction, int timestep)
{
  int i;
  for (i = 0; i < r->nflux; i++)
    if (strcmp(r->fluxNames[i], reaction) == 0) return r->flux[((size_t)set * (size_t)r->nout + (size_t)timestep) * (size_t)r->nflux + (size_t)i];
  SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_SYMBOL_IS_NOT_IN_MODEL, "Reaction %s has no flux course in these results", reaction);
  return 0.0;
}

/* one batched device run over all design points; when mapped is given, the per-point SBMLResults_t are built while
   the scanned kinetic-law parameters are still globalised (their names appear in the kinetic laws) */
static batchResults_t *run_design_points(Model_t *m, cvodeSettings_t *set, varySettings_t *vs, SBMLResultsArray_t **mapped)
{
  int i, j, nvalues, *idx; odeModel_t *om; batchResults_t *res = NULL; double *rows;
  for (i = 0; i < vs->nrparams; i++) if (vs->rid[i] && strlen(vs->rid[i]) > 0) Model_globalizeParameter(m, vs->id[i], vs->rid[i]);
  om = ODEModel_create(m);
  if (om) {
    idx = (int *)malloc((size_t)(vs->nrparams > 0 ? vs->nrparams : 1) * sizeof(int));
    for (j = 0; j < vs->nrparams; j++) {
      if (vs->rid[j] && strlen(vs->rid[j]) > 0) {
        char *nm = (char *)malloc(strlen(vs->id[j]) + strlen(vs->rid[j]) + 4);
        sprintf(nm, "r_%s_%s", vs->rid[j], vs->id[j]);
        idx[j] = ODEModel_getVariableIndexFields(om, nm);
        free(nm);
      } else idx[j] = ODEModel_getVariableIndexFields(om, vs->id[j]);
      if (idx[j] < 0) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_SYMBOL_IS_NOT_IN_MODEL, "Symbol %s is not in the model", vs->id[j]); free(idx); idx = NULL; break; }
    }
    if (idx) {
      nvalues = ODEModel_getNumValues(om);
      rows = (double *)malloc((size_t)vs->nrdesignpoints * (size_t)(vs->nrparams > 0 ? vs->nrparams : 1) * sizeof(double));
      for (i = 0; i < vs->nrdesignpoints; i++) for (j = 0; j < vs->nrparams; j++) rows[(size_t)i * (size_t)vs->nrparams + (size_t)j] = vs->params[i][j];
      res = (batchResults_t *)calloc(1, sizeof *res);
      res->nsets = vs->nrdesignpoints; res->nout = set->PrintStep + 1; res->nvalues = nvalues;
      res->names = (char **)calloc((size_t)nvalues, sizeof(char *));
      for (i = 0; i < nvalues; i++) res->names[i] = sdup(om->names[i]);
      res->time = (double *)malloc((size_t)res->nout * sizeof(double));
      for (i = 0; i < res->nout; i++) res->time[i] = set->TimePoints[i];
      res->value = (double *)malloc((size_t)res->nout * (size_t)nvalues * (size_t)res->nsets * sizeof(double));
      res->status = (int *)calloc((size_t)(res->nsets > 0 ? res->nsets : 1), sizeof(int));
      /* all design points in one batched run, spread over the GPUs of the box when there are enough of them */
      if (set->Sensitivity) {           /* per design point what SBMLResults_createSens maps (src/odeSolver.c:545-575) */
        odeSense_t *os;
        if (set->UseJacobian && !om->jacob && !om->jacobianFailed) ODEModel_constructJacobian(om);
        os = ODESense_create(om, set);
        if (os && os->nsens > 0) {
          res->nsens = os->nsens; res->neq = om->neq;
          res->sensParam = (char **)calloc((size_t)os->nsens, sizeof(char *));
          for (i = 0; i < os->nsens; i++) res->sensParam[i] = sdup(om->names[os->index_sens[i]]);
          res->sens = (double *)malloc((size_t)res->nsets * (size_t)res->nout * (size_t)os->nsens * (size_t)om->neq * sizeof(double));
        }
        ODESense_free(os);
      }
      /* reaction fluxes: evaluated on the device from the stored rows (the reference walks every kinetic law at every
         time point of every design point on the host, src/odeSolver.c:519-531); ODES_HOST_FLUX=1 keeps that path */
      if (Model_getNumReactions(m) > 0 && !(getenv("ODES_HOST_FLUX") && atoi(getenv("ODES_HOST_FLUX"))) && !(getenv("ODES_FLUX") && !atoi(getenv("ODES_FLUX")))) {
        res->nflux = (int)Model_getNumReactions(m);
        res->fluxNames = (char **)calloc((size_t)res->nflux, sizeof(char *));
        for (i = 0; i < res->nflux; i++) res->fluxNames[i] = sdup(Model_getReaction(m, (unsigned int)i)->id);
        res->flux = (double *)malloc((size_t)res->nsets * (size_t)res->nout * (size_t)res->nflux * sizeof(double) + 8);
      }
      if (!ODESBatch_runHostAllDevicesEx(om, set, vs->nrparams, idx, 0, NULL, res->nsets, rows, res->value, res->sens, res->flux, res->status, 0)) { BatchResults_free(res); res = NULL; }
      free(rows);
      if (res && mapped) *mapped = SBMLResultsArray_fromBatch(m, om, res);
    }
    free(idx);
    om->m = NULL;                    /* the SBML model stays with the caller */
    ODEModel_free(om);
  }
  for (i = 0; i < vs->nrparams; i++) if (vs->rid[i] && strlen(vs->rid[i]) > 0) Model_localizeParameter(m, vs->id[i], vs->rid[i]);
  return res;
}
batchResults_t *Model_odeSolverBatchFlat(Model_t *m, cvodeSettings_t *set, varySettings_t *vs) { return run_design_points(m, set, vs, NULL); }
batchResults_t *SBML_odeSolverBatchFlat(SBMLDocument_t *d, cvodeSettings_t *set, varySettings_t *vs)
{ return (d && d->model) ? Model_odeSolverBatchFlat(d->model, set, vs) : NULL; }
/* the reference's call (src/odeSolver.c:235-352): one SBMLResults_t per design point */
SBMLResultsArray_t *Model_odeSolverBatch(Model_t *m, cvodeSettings_t *set, varySettings_t *vs)
{
  SBMLResultsArray_t *a = NULL;
  batchResults_t *flat = run_design_points(m, set, vs, &a);
  BatchResults_free(flat);
  return a;
}
SBMLResultsArray_t *SBML_odeSolverBatch(SBMLDocument_t *d, cvodeSettings_t *set, varySettings_t *vs)
{ return (d && d->model) ? Model_odeSolverBatch(d->model, set, vs) : NULL; }
