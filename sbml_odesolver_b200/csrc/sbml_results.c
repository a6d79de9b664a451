/*
 * sbml_results.c -- results mapped back onto SBML structures, and the high-level one-call drivers built on them.
 * Reference: src/sbmlResults.c:66-560 (timeCourse_t / SBMLResults_t containers, accessors, dumps),
 * src/odeSolver.c:80-232 (SBML_odeSolver, Model_odeSolver), :441-543 (SBMLResults_fromIntegrator).
 *
 * Reaction fluxes follow the reference: the kinetic law of every reaction, with its local parameters and the
 * model's constants substituted as numbers (AST_replaceNameByParameters, AST_replaceConstants), evaluated on the
 * stored values of every output time.  Consequence kept on purpose: in a parameter scan the substituted constants
 * are the MODEL's values, not the design point's (src/odeSolver.c:463-471 reads them from Model_t) -- the per-point
 * fluxes at the varied values are the assigned variables named after the reactions in the flat results
 * (batchResults_t, Model_odeSolverBatchFlat).
 * Deviation: the reference indexes the compartment time courses by the compartment's position in the model
 * (src/sbmlResults.c:387), which overruns the array when a constant compartment precedes a variable one; a running
 * index is used here.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sbmlsolver/odeSolver.h"
#include "sbmlsolver/processAST.h"
#include "sbmlsolver/solverError.h"

static char *sdup(const char *s) { char *r; if (!s) s = ""; r = (char *)malloc(strlen(s) + 1); strcpy(r, s); return r; }

static timeCourse_t *TimeCourse_create(int timepoints, const char *name)
{
  timeCourse_t *tc = (timeCourse_t *)calloc(1, sizeof *tc);
  tc->timepoints = timepoints;
  tc->values = (double *)calloc((size_t)(timepoints > 0 ? timepoints : 1), sizeof(double));
  tc->name = sdup(name);
  return tc;
}
static void TimeCourse_free(timeCourse_t *tc) { if (!tc) return; free(tc->name); free(tc->values); free(tc); }
static timeCourseArray_t *TimeCourseArray_create(int num_val)
{
  timeCourseArray_t *a = (timeCourseArray_t *)calloc(1, sizeof *a);
  a->num_val = num_val;
  a->tc = (timeCourse_t **)calloc((size_t)(num_val > 0 ? num_val : 1), sizeof(timeCourse_t *));
  return a;
}
static void TimeCourseArray_free(timeCourseArray_t *a)
{
  int i;
  if (!a) return;
  for (i = 0; i < a->num_val; i++) TimeCourse_free(a->tc[i]);
  free(a->tc); free(a);
}
static timeCourse_t *TimeCourseArray_getTimeCourse(const char *id, timeCourseArray_t *a)
{
  int i;
  if (!a || !id) return NULL;
  for (i = 0; i < a->num_val; i++) if (strcmp(id, a->tc[i]->name) == 0) return a->tc[i];
  return NULL;
}
static void TimeCourseArray_dump(const timeCourseArray_t *a, const timeCourse_t *time)
{
  int i, j;
  if (!a || a->num_val == 0) { printf("## No Values.\n"); return; }
  printf("#time ");
  for (j = 0; j < a->num_val; j++) printf("%s ", a->tc[j]->name);
  printf("\n");
  for (i = 0; i < time->timepoints; i++) {
    printf("%g ", time->values[i]);
    for (j = 0; j < a->num_val; j++) printf("%g ", a->tc[j]->values[i]);
    printf("\n");
  }
}

timeCourse_t *SBMLResults_getTime(SBMLResults_t *r) { return r->time; }
int SBMLResults_getNout(const SBMLResults_t *r) { return r->time->timepoints; }
int SBMLResults_getNumSens(const SBMLResults_t *r) { return r->nsens; }
const char *SBMLResults_getSensParam(const SBMLResults_t *r, int i) { return (r->param && i >= 0 && i < r->nsens) ? r->param[i] : NULL; }
timeCourse_t *Compartment_getTimeCourse(Compartment_t *c, SBMLResults_t *r) { return TimeCourseArray_getTimeCourse(c->id, r->compartments); }
timeCourse_t *Species_getTimeCourse(Species_t *s, SBMLResults_t *r) { return TimeCourseArray_getTimeCourse(s->id, r->species); }
timeCourse_t *Parameter_getTimeCourse(Parameter_t *p, SBMLResults_t *r) { return TimeCourseArray_getTimeCourse(p->id, r->parameters); }
timeCourse_t *SBMLResults_getTimeCourse(SBMLResults_t *r, const char *id)
{
  timeCourse_t *tc;
  if ((tc = TimeCourseArray_getTimeCourse(id, r->species))) return tc;
  if ((tc = TimeCourseArray_getTimeCourse(id, r->compartments))) return tc;
  if ((tc = TimeCourseArray_getTimeCourse(id, r->parameters))) return tc;
  return TimeCourseArray_getTimeCourse(id, r->fluxes);
}
const char *TimeCourse_getName(const timeCourse_t *tc) { return tc->name; }
int TimeCourse_getNumValues(const timeCourse_t *tc) { return tc->timepoints; }
double TimeCourse_getValue(const timeCourse_t *tc, int i) { return tc->values[i]; }
double TimeCourse_getSensitivity(const timeCourse_t *tc, int i, int j) { return tc->sensitivity ? tc->sensitivity[i][j] : 0.0; }
/* reference sbmlResults.h:134 declares Model_setValues without defining it; the natural reading next to
   Model_odeSolver: carry the end state of a run back into the SBML model (species keep the quantity they were set in) */
Model_t *Model_setValues(Model_t *m, SBMLResults_t *r)
{
  int k;
  if (!m || !r) return m;
  for (k = 0; r->species && k < r->species->num_val; k++) {
    const timeCourse_t *tc = r->species->tc[k]; Species_t *sp = Model_getSpeciesById(m, tc->name);
    if (!sp || tc->timepoints < 1) continue;
    if (sp->isSetInitialAmount && !sp->isSetInitialConcentration) sp->initialAmount = tc->values[tc->timepoints - 1];
    else { sp->initialConcentration = tc->values[tc->timepoints - 1]; sp->isSetInitialConcentration = 1; }
  }
  for (k = 0; r->compartments && k < r->compartments->num_val; k++) {
    const timeCourse_t *tc = r->compartments->tc[k]; Compartment_t *c = Model_getCompartmentById(m, tc->name);
    if (c && tc->timepoints > 0) { c->size = tc->values[tc->timepoints - 1]; c->isSetSize = 1; }
  }
  for (k = 0; r->parameters && k < r->parameters->num_val; k++) {
    const timeCourse_t *tc = r->parameters->tc[k]; Parameter_t *p = Model_getParameterById(m, tc->name);
    if (p && tc->timepoints > 0) Parameter_setValue(p, tc->values[tc->timepoints - 1]);
  }
  return m;
}
void SBMLResults_dumpSpecies(const SBMLResults_t *r) { printf("## Printing Species time courses\n"); TimeCourseArray_dump(r->species, r->time); }
void SBMLResults_dumpCompartments(const SBMLResults_t *r) { printf("## Printing Variable Compartment time courses\n"); TimeCourseArray_dump(r->compartments, r->time); }
void SBMLResults_dumpParameters(const SBMLResults_t *r) { printf("## Printing Variable Parameter time courses\n"); TimeCourseArray_dump(r->parameters, r->time); }
void SBMLResults_dumpFluxes(const SBMLResults_t *r) { printf("## Printing Reaction Flux time courses\n"); TimeCourseArray_dump(r->fluxes, r->time); }
void SBMLResults_dump(const SBMLResults_t *r)
{
  printf("### Printing All Results \n");
  SBMLResults_dumpCompartments(r); SBMLResults_dumpSpecies(r); SBMLResults_dumpParameters(r); SBMLResults_dumpFluxes(r);
}
void SBMLResults_free(SBMLResults_t *r)
{
  if (!r) return;
  if (r->nsens > 0) {                          /* sensitivity courses hang off the species' time courses */
    timeCourseArray_t *arrs[3]; int a, i, j;
    arrs[0] = r->species; arrs[1] = r->compartments; arrs[2] = r->parameters;
    for (a = 0; a < 3; a++) for (i = 0; arrs[a] && i < arrs[a]->num_val; i++) {
      timeCourse_t *tc = arrs[a]->tc[i];
      if (!tc || !tc->sensitivity) continue;
      for (j = 0; j < r->nsens; j++) free(tc->sensitivity[j]);
      free(tc->sensitivity); tc->sensitivity = NULL;
    }
    for (j = 0; j < r->nsens; j++) free(r->param[j]);
    free(r->param);
  }
  TimeCourse_free(r->time);
  TimeCourseArray_free(r->species); TimeCourseArray_free(r->compartments);
  TimeCourseArray_free(r->parameters); TimeCourseArray_free(r->fluxes);
  free(r);
}
SBMLResultsArray_t *SBMLResultsArray_allocate(int size)
{
  SBMLResultsArray_t *a = (SBMLResultsArray_t *)calloc(1, sizeof *a);
  a->results = (SBMLResults_t **)calloc((size_t)(size > 0 ? size : 1), sizeof(SBMLResults_t *));
  a->size = size;
  return a;
}
void SBMLResultsArray_free(SBMLResultsArray_t *a)
{
  int i;
  if (!a) return;
  for (i = 0; i < a->size; i++) SBMLResults_free(a->results[i]);
  free(a->results); free(a);
}
int SBMLResultsArray_getNumResults(const SBMLResultsArray_t *a) { return a->size; }
SBMLResults_t *SBMLResultsArray_getResults(SBMLResultsArray_t *a, int i) { return (i >= 0 && i < a->size) ? a->results[i] : NULL; }
SBMLResults_t *SBMLResultsMatrix_getResults(SBMLResultsMatrix_t *m, int i, int j) { return (m && i >= 0 && i < m->i && j >= 0 && j < m->j) ? m->results[i][j] : NULL; }

/* containers for all species, the variable compartments and parameters, and all reactions (src/sbmlResults.c:342-430) */
SBMLResults_t *SBMLResults_create(Model_t *m, int timepoints)
{
  unsigned int i; int n;
  SBMLResults_t *r = (SBMLResults_t *)calloc(1, sizeof *r);
  r->time = TimeCourse_create(timepoints, "time");
  r->species = TimeCourseArray_create((int)Model_getNumSpecies(m));
  for (i = 0; i < Model_getNumSpecies(m); i++) r->species->tc[i] = TimeCourse_create(timepoints, Model_getSpecies(m, i)->id);
  for (n = 0, i = 0; i < Model_getNumCompartments(m); i++) if (!Model_getCompartment(m, i)->constant) n++;
  r->compartments = TimeCourseArray_create(n);
  for (n = 0, i = 0; i < Model_getNumCompartments(m); i++)
    if (!Model_getCompartment(m, i)->constant) r->compartments->tc[n++] = TimeCourse_create(timepoints, Model_getCompartment(m, i)->id);
  for (n = 0, i = 0; i < Model_getNumParameters(m); i++) if (!Model_getParameter(m, i)->constant) n++;
  r->parameters = TimeCourseArray_create(n);
  for (n = 0, i = 0; i < Model_getNumParameters(m); i++)
    if (!Model_getParameter(m, i)->constant) r->parameters->tc[n++] = TimeCourse_create(timepoints, Model_getParameter(m, i)->id);
  r->fluxes = TimeCourseArray_create((int)Model_getNumReactions(m));
  for (i = 0; i < Model_getNumReactions(m); i++) r->fluxes->tc[i] = TimeCourse_create(timepoints, Model_getReaction(m, i)->id);
  return r;
}

/* ---- mapping plan: where each time course finds its values, resolved once per model instead of once per time point
   and design point (the reference searches names[] with strcmp inside the time loop, src/odeSolver.c:487-518) ---- */
typedef struct {
  int nsp, ncomp, npar, nflux;
  int *sp, *comp, *par;      /* index into value[] or -1 */
  ASTNode_t **kls;           /* kinetic laws with parameters and constants substituted, names indexed */
} resultMap_t;

static int name_index(const odeModel_t *om, int nvalues, const char *id)
{
  int k, hit = -1;
  for (k = 0; k < nvalues; k++) if (strcmp(id, om->names[k]) == 0) hit = k;     /* the reference's loop keeps the last match */
  return hit;
}
static resultMap_t *ResultMap_create(Model_t *m, const odeModel_t *om, const SBMLResults_t *r)
{
  int j, nvalues = om->neq + om->nass + om->nconst; unsigned int i;
  resultMap_t *p = (resultMap_t *)calloc(1, sizeof *p);
  p->nsp = r->species->num_val; p->ncomp = r->compartments->num_val; p->npar = r->parameters->num_val; p->nflux = r->fluxes->num_val;
  p->sp = (int *)malloc(sizeof(int) * (size_t)(p->nsp + 1)); p->comp = (int *)malloc(sizeof(int) * (size_t)(p->ncomp + 1)); p->par = (int *)malloc(sizeof(int) * (size_t)(p->npar + 1));
  for (j = 0; j < p->nsp; j++) p->sp[j] = name_index(om, nvalues, r->species->tc[j]->name);
  for (j = 0; j < p->ncomp; j++) p->comp[j] = name_index(om, nvalues, r->compartments->tc[j]->name);
  for (j = 0; j < p->npar; j++) p->par[j] = name_index(om, nvalues, r->parameters->tc[j]->name);
  (void)i;
  SBMLResults_fluxASTs(m, om, &p->kls);
  return p;
}
/* the indexed kinetic laws of the model's reactions, in reaction order (NULL where a reaction has none): what the flux
   courses are evaluated from -- on the host by ResultMap_fill, on the device by the emitted flux_row (emit.c).  The caller
   frees the trees and the array.  Returns the number of reactions. */
int SBMLResults_fluxASTs(Model_t *m, const odeModel_t *om, ASTNode_t ***kls)
{
  int nvalues = om->neq + om->nass + om->nconst, n = (int)Model_getNumReactions(m); unsigned int i;
  *kls = (ASTNode_t **)calloc((size_t)(n + 1), sizeof(ASTNode_t *));
  for (i = 0; i < Model_getNumReactions(m); i++) {
    Reaction_t *rx = Model_getReaction(m, i);
    ASTNode_t *kl;
    if (!rx->kineticLaw || !rx->kineticLaw->math) continue;
    kl = copyAST(rx->kineticLaw->math);
    AST_replaceNameByParameters(kl, &rx->kineticLaw->parameters);
    AST_replaceConstants(m, kl);
    (*kls)[i] = indexAST(kl, nvalues, om->names);       /* same values as the name search of evaluateAST, found once */
    ASTNode_free(kl);
  }
  return n;
}
static void ResultMap_free(resultMap_t *p)
{
  int i;
  if (!p) return;
  for (i = 0; i < p->nflux; i++) if (p->kls[i]) ASTNode_free(p->kls[i]);
  free(p->kls); free(p->sp); free(p->comp); free(p->par); free(p);
}
/* one time point: value[] holds the stored values of that time, data is scratch for the flux evaluation */
static void ResultMap_fill(const resultMap_t *p, SBMLResults_t *r, int n, double time, const double *value, cvodeData_t *data, const double *flux)
{
  int j;
  r->time->values[n] = time;
  data->currenttime = (float)time;
  for (j = 0; j < data->nvalues; j++) data->value[j] = value[j];
  for (j = 0; j < p->nsp; j++) if (p->sp[j] >= 0) r->species->tc[j]->values[n] = value[p->sp[j]];
  for (j = 0; j < p->ncomp; j++) if (p->comp[j] >= 0) r->compartments->tc[j]->values[n] = value[p->comp[j]];
  for (j = 0; j < p->npar; j++) if (p->par[j] >= 0) r->parameters->tc[j]->values[n] = value[p->par[j]];
  if (flux) { for (j = 0; j < p->nflux; j++) if (p->kls[j]) r->fluxes->tc[j]->values[n] = flux[j]; }      /* evaluated on the device */
  else for (j = 0; j < p->nflux; j++) if (p->kls[j]) r->fluxes->tc[j]->values[n] = evaluateAST(p->kls[j], data);
}

SBMLResults_t *SBMLResults_fromIntegrator(Model_t *m, integratorInstance_t *ii)
{
  odeModel_t *om; cvodeData_t *data; cvodeResults_t *cv; SBMLResults_t *r; resultMap_t *map; double *row; int n, j;
  if (!m || !ii) return NULL;
  om = ii->om; data = ii->data; cv = ii->results;
  if (!data || !cv) return NULL;
  r = SBMLResults_create(m, cv->nout + 1);
  map = ResultMap_create(m, om, r);
  row = (double *)malloc(sizeof(double) * (size_t)(data->nvalues > 0 ? data->nvalues : 1));
  for (n = 0; n < r->time->timepoints; n++) {
    for (j = 0; j < data->nvalues; j++) row[j] = cv->value[j][n];
    ResultMap_fill(map, r, n, cv->time[n], row, data, NULL);      /* leaves data at the last time point, as the reference does */
  }
  free(row);
  ResultMap_free(map);
  /* SBMLResults_createSens (src/odeSolver.c:545-575): parameter ids, and per ODE variable its sensitivity courses */
  r->nsens = 0;
  if (cv->nsens > 0 && cv->sensitivity && data->os) {
    int i, k;
    r->nsens = cv->nsens;
    r->param = (char **)calloc((size_t)cv->nsens, sizeof(char *));
    for (j = 0; j < cv->nsens; j++) r->param[j] = sdup(om->names[data->os->index_sens[j]]);
    for (i = 0; i < cv->neq; i++) {
      timeCourse_t *tc = SBMLResults_getTimeCourse(r, om->names[i]);
      if (!tc) continue;
      tc->sensitivity = (double **)calloc((size_t)cv->nsens, sizeof(double *));
      for (j = 0; j < cv->nsens; j++) {
        tc->sensitivity[j] = (double *)calloc((size_t)(cv->nout + 1), sizeof(double));
        for (k = 0; k <= cv->nout; k++) tc->sensitivity[j][k] = cv->sensitivity[i][j][k];
      }
    }
  }
  return r;
}

SBMLResults_t *Model_odeSolver(Model_t *m, cvodeSettings_t *set)
{
  odeModel_t *om; integratorInstance_t *ii; SBMLResults_t *r;
  om = ODEModel_create(m);
  if (!om) return NULL;
  ii = IntegratorInstance_create(om, set);
  if (!ii) { om->m = NULL; ODEModel_free(om); return NULL; }
  while (!IntegratorInstance_timeCourseCompleted(ii)) if (!IntegratorInstance_integrateOneStep(ii)) break;
  r = SBMLResults_fromIntegrator(m, ii);
  IntegratorInstance_free(ii);
  om->m = NULL;                      /* the SBML model stays with the caller */
  ODEModel_free(om);
  return r;
}
SBMLResults_t *SBML_odeSolver(SBMLDocument_t *d, cvodeSettings_t *set) { return (d && d->model) ? Model_odeSolver(d->model, set) : NULL; }

/* flat batch results -> one SBMLResults_t per design point.  The model must be in the state the batch ran with
   (scanned kinetic-law parameters still globalised); rows of a failed design point stop at its last stored time */
SBMLResultsArray_t *SBMLResultsArray_fromBatch(Model_t *m, odeModel_t *om, const batchResults_t *b)
{
  SBMLResultsArray_t *a; SBMLResults_t *proto; resultMap_t *map; cvodeData_t *data; int i, n;
  if (!m || !om || !b) return NULL;
  a = SBMLResultsArray_allocate(b->nsets);
  data = CvodeData_create(om);
  proto = SBMLResults_create(m, b->nout);
  map = ResultMap_create(m, om, proto);
  SBMLResults_free(proto);
  for (i = 0; i < b->nsets; i++) {
    const double *v = b->value + (size_t)i * (size_t)b->nout * (size_t)b->nvalues;
    int np = b->nout;
    /* rows a run did not reach (solver failure, halt on event, halt on steady state) are not-a-number in the flat
       results: the reference's results end at the last stored index (integratorInstance.c:1141) */
    while (np > 1 && b->nvalues > 0 && v[(size_t)(np - 1) * (size_t)b->nvalues] != v[(size_t)(np - 1) * (size_t)b->nvalues]) np--;
    a->results[i] = SBMLResults_create(m, np);
    /* fluxes: the device evaluated them (b->flux, Model_odeSolverBatch); else the kinetic laws are walked here */
    for (n = 0; n < np; n++) ResultMap_fill(map, a->results[i], n, b->time[n], v + (size_t)n * (size_t)b->nvalues, data,
                                            (b->flux && b->nflux == map->nflux) ? b->flux + ((size_t)i * (size_t)b->nout + (size_t)n) * (size_t)b->nflux : NULL);
    if (b->nsens > 0 && b->sens) {             /* SBMLResults_createSens per design point (src/odeSolver.c:545-575) */
      SBMLResults_t *r = a->results[i]; int j, e;
      const double *sv = b->sens + (size_t)i * (size_t)b->nout * (size_t)b->nsens * (size_t)b->neq;
      r->nsens = b->nsens;
      r->param = (char **)calloc((size_t)b->nsens, sizeof(char *));
      for (j = 0; j < b->nsens; j++) r->param[j] = sdup(b->sensParam[j]);
      for (e = 0; e < b->neq; e++) {
        timeCourse_t *tc = SBMLResults_getTimeCourse(r, om->names[e]);
        if (!tc) continue;
        tc->sensitivity = (double **)calloc((size_t)b->nsens, sizeof(double *));
        for (j = 0; j < b->nsens; j++) {
          tc->sensitivity[j] = (double *)calloc((size_t)np, sizeof(double));
          for (n = 0; n < np; n++) tc->sensitivity[j][n] = sv[((size_t)n * (size_t)b->nsens + (size_t)j) * (size_t)b->neq + (size_t)e];
        }
      }
    }
  }
  ResultMap_free(map);
  CvodeData_free(data);
  return a;
}
