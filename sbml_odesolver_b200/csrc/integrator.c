/*
 * integrator.c -- integratorInstance_t: the single-instance driver API of the
 * reference (src/integratorInstance.c:105-1790, src/cvodeSolver.c:81-400),
 * kept call-compatible, with the time stepping moved to the GPU.
 *
 * How the solver seam is honoured without a CPU integrator: when the engine
 * is (re)validated -- first step, after IntegratorInstance_setVariableValue or
 * after a fired event invalidated it, exactly where the reference calls
 * CVodeInit/CVodeReInit (src/cvodeSolver.c:91-100) -- the remaining time
 * course is integrated on the device as a batch of one, with every ODE
 * variable and every constant passed as a per-trajectory input.  Each
 * IntegratorInstance_cvodeOneStep then advances solver->t to solver->tout,
 * copies the row computed for that output time into data->value and calls
 * IntegratorInstance_updateData, as the reference's contract demands
 * (src/cvodeSolver.c:72-79, :226-240).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sbmlsolver/batchIntegrator.h"
#include "sbmlsolver/processAST.h"
#include "sbmlsolver/solverError.h"

/* private plan entry points (batch.cu) */
int ODESBatch_setTimeGrid(odesBatch_t *, int nout, const double *tpoints);
int ODESBatch_setInitialTriggers(odesBatch_t *, const int *trigger);
/* on = 0: launches skip the model's event function (IntegratorInstance_integrateOneStepWithoutEventProcessing,
   src/integratorInstance.c:843-856) */
void ODESBatch_setEventProcessing(odesBatch_t *, int on);
int ODESBatch_setInitialSensitivities(odesBatch_t *, const double *s0);

#define ODES_LOOKAHEAD 256        /* output steps computed ahead by one device run of an indefinite integration */
static void drop_course(cvodeSolver_t *s)
{
  free(s->course); free(s->courseFired); free(s->courseSens);
  s->course = NULL; s->courseFired = NULL; s->courseSens = NULL; s->courseLen = 0; s->courseFirst = 0; s->courseStatus = 0;
}

/* the settings a compiled plan depends on (kernel specialisation and buffer sizes); callers change cvodeSettings_t objects
   in place between runs (examples/sensitivity.c switches the sensitivities off and resets), so the plan is keyed on the
   values, not on the pointer */
static unsigned long plan_signature(const cvodeSettings_t *opt)
{
  unsigned long h = 1469598103UL; int i; const char *c;
  const long v[] = { opt->PrintStep, opt->Indefinitely, opt->CvodeMethod, opt->IterMethod, opt->MaxOrder, opt->ResetCvodeOnEvent, opt->SetTStop,
                     opt->Sensitivity, opt->nsens, opt->SensMethod, opt->HaltOnEvent, opt->SteadyState, opt->UseJacobian, opt->sensIDs != NULL,
                     opt->StepResolution > 1 ? opt->StepResolution : 1 };
  for (i = 0; i < (int)(sizeof v / sizeof v[0]); i++) h = (h ^ (unsigned long)v[i]) * 1099511UL;
  for (i = 0; opt->sensIDs && i < opt->nsens; i++) for (c = opt->sensIDs[i]; *c; c++) h = (h ^ (unsigned long)*c) * 1099511UL;
  { double d = opt->ssThreshold + (opt->Indefinitely ? opt->Time : 0.0); unsigned char b[sizeof d]; memcpy(b, &d, sizeof d); for (i = 0; i < (int)sizeof d; i++) h = (h ^ b[i]) * 1099511UL; }
  return h;
}
static int initialize_solver(integratorInstance_t *engine, cvodeData_t *data, cvodeSettings_t *opt, odeModel_t *om)
{
  int i; cvodeSolver_t *solver = engine->solver; cvodeResults_t *results = data->results;
  engine->om = om; engine->opt = opt; engine->data = data; engine->results = data->results;
  if (solver->plan && solver->planSignature != plan_signature(opt)) {
    ODESBatch_free(solver->plan); solver->plan = NULL;
    if (solver->planOpt) { CvodeSettings_free(solver->planOpt); solver->planOpt = NULL; }
  }
  solver->planSignature = plan_signature(opt);
  /* IntegratorInstance_initializeJacobian (integratorInstance.c:291-311) */
  if (om->jacobianFailed > 0 || !opt->UseJacobian) engine->UseJacobian = 0;
  else if (om->jacob != NULL) engine->UseJacobian = 1;
  else engine->UseJacobian = ODEModel_constructJacobian(om);
  if (opt->Indefinitely) { solver->t0 = 0; solver->t = 0; solver->tout = opt->Time; }
  else { solver->t0 = opt->TimePoints[0]; solver->t = opt->TimePoints[0]; solver->tout = opt->PrintStep >= 1 ? opt->TimePoints[1] : opt->TimePoints[0]; }
  solver->nout = opt->PrintStep;
  solver->iout = 1;
  if (opt->StoreResults && results) {
    results->time[0] = data->currenttime;
    for (i = 0; i < data->nvalues; i++) results->value[i][0] = data->value[i];
  }
  /* forward sensitivities (IntegratorInstance_initializeSolver, src/integratorInstance.c:360-420): structures for the
     selected parameters, current sensitivities 0 / 1, row 0 of the stored courses */
  if (engine->os) { ODESense_free(engine->os); engine->os = NULL; }
  if (opt->Sensitivity) {
    engine->os = ODESense_create(om, opt);
    if (!engine->os) return 0;
  }
  if (!CvodeData_initializeSensitivities(data, opt, om, engine->os)) return 0;
  solver->nsens = engine->os ? engine->os->nsens : 0;
  engine->run++;
  engine->isValid = 0;
  engine->clockStarted = 0;
  drop_course(solver);
  return 1;
}

integratorInstance_t *IntegratorInstance_create(odeModel_t *om, cvodeSettings_t *opt)
{
  cvodeData_t *data; integratorInstance_t *engine;
  if (!om || !opt || om->hasCycle) return NULL;
  data = CvodeData_create(om);
  if (!data) return NULL;
  CvodeData_initialize(data, opt, om, 0);
  engine = (integratorInstance_t *)calloc(1, sizeof *engine);
  engine->solver = (cvodeSolver_t *)calloc(1, sizeof(cvodeSolver_t));
  if (!initialize_solver(engine, data, opt, om)) { IntegratorInstance_free(engine); return NULL; }
  return engine;
}
int IntegratorInstance_set(integratorInstance_t *engine, cvodeSettings_t *opt)
{
  if (opt != engine->opt && engine->solver->plan) { ODESBatch_free(engine->solver->plan); engine->solver->plan = NULL; }
  if (opt != engine->opt && engine->solver->planOpt) { CvodeSettings_free(engine->solver->planOpt); engine->solver->planOpt = NULL; }
  CvodeData_initialize(engine->data, opt, engine->om, 0);
  return initialize_solver(engine, engine->data, opt, engine->om);
}
int IntegratorInstance_reset(integratorInstance_t *engine) { return IntegratorInstance_set(engine, engine->opt); }
int IntegratorInstance_resetTime(integratorInstance_t *engine, cvodeSettings_t *opt)
{
  if (!opt) opt = engine->opt;
  CvodeData_initialize(engine->data, opt, engine->om, 1);
  return initialize_solver(engine, engine->data, opt, engine->om);
}
void IntegratorInstance_free(integratorInstance_t *engine)
{
  if (!engine) return;
  if (engine->solver) { drop_course(engine->solver); if (engine->solver->plan) ODESBatch_free(engine->solver->plan); CvodeSettings_free(engine->solver->planOpt); free(engine->solver); }
  CvodeData_free(engine->data);
  ODESense_free(engine->os);
  free(engine);
}
cvodeSettings_t *IntegratorInstance_getSettings(integratorInstance_t *engine) { return engine->opt; }
cvodeData_t *IntegratorInstance_getData(integratorInstance_t *engine) { return engine->data; }
double IntegratorInstance_getTime(const integratorInstance_t *engine) { return engine->solver->t; }
int IntegratorInstance_setNextTimeStep(integratorInstance_t *engine, double nexttime)
{
  if (nexttime <= engine->solver->t) return 0;
  engine->solver->tout = nexttime;
  drop_course(engine->solver); engine->isValid = 0;
  return 1;
}
const cvodeResults_t *IntegratorInstance_getResults(const integratorInstance_t *engine) { return engine->results; }
int IntegratorInstance_timeCourseCompleted(const integratorInstance_t *engine)
{ return engine->solver->iout > engine->solver->nout && !engine->opt->Indefinitely; }

static void refresh_rules(integratorInstance_t *engine)
{
  int i; odeModel_t *om = engine->om; cvodeData_t *data = engine->data;
  for (i = 0; i < om->nass; i++) { nonzeroElem_t *o = om->assignmentOrder[i]; data->value[o->i] = evaluateAST(o->ij, data); }
  data->allRulesUpdated = 1;
}
double IntegratorInstance_getVariableValue(integratorInstance_t *engine, variableIndex_t *vi)
{
  if (!engine->data->allRulesUpdated) refresh_rules(engine);
  return engine->data->value[vi->index];
}
static void set_value_by_index(integratorInstance_t *engine, int idx, double value)
{
  odeModel_t *om = engine->om; cvodeData_t *data = engine->data;
  if (idx >= om->neq && idx < om->neq + om->nass) {
    SolverError_error(WARNING_ERROR_TYPE, SOLVER_ERROR_ATTEMPT_TO_SET_ASSIGNED_VALUE,
                      "Attempted to set a new value for an assigned variable: %s. This is not possible. New value ignored!", om->names[idx]);
    return;
  }
  data->value[idx] = value;
  if (engine->solver->t == 0.0 && data->results != NULL) data->results->value[idx][0] = value;
  if (idx < om->neq || engine->opt->ResetCvodeOnEvent) engine->isValid = 0;
  /* a changed constant alters the right-hand side of the look-ahead course as well */
  engine->isValid = 0;
  refresh_rules(engine);
}
void IntegratorInstance_setVariableValue(integratorInstance_t *engine, variableIndex_t *vi, double value) { if (vi) set_value_by_index(engine, vi->index, value); }
void IntegratorInstance_setVariableValueByID(integratorInstance_t *engine, const char *id, double value)
{
  variableIndex_t *vi = ODEModel_getVariableIndex(engine->om, id);
  if (vi) set_value_by_index(engine, vi->index, value);
  VariableIndex_free(vi);
}
void IntegratorInstance_copyVariableState(integratorInstance_t *target, integratorInstance_t *source)
{
  int i;
  if (target->om == source->om) for (i = 0; i < source->data->nvalues; i++) target->data->value[i] = source->data->value[i];
  else SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_ATTEMPTING_TO_COPY_VARIABLE_STATE_BETWEEN_INSTANCES_OF_DIFFERENT_MODELS,
                         "Attempting to copy variable state between instances of different models");
  target->isValid = 0;
  refresh_rules(target);
}
void IntegratorInstance_dumpData(const integratorInstance_t *engine)
{
  int i; cvodeData_t *data = engine->data;                  /* one row: time, then all values (src/integratorInstance.c:737-746) */
  printf("%g  ", data->currenttime);
  for (i = 0; i < data->nvalues; i++) printf("%g ", data->value[i]);
  printf("\n");
}

/* device look-ahead: integrates outputs iout..nout from the current state as a batch of one */
static int compute_course(integratorInstance_t *engine)
{
  odeModel_t *om = engine->om; cvodeSolver_t *solver = engine->solver; cvodeData_t *data = engine->data;
  int nvalues = data->nvalues, nin = om->neq + om->nconst, i, j, nrem, st = 0;
  double *grid, *values, *rows; unsigned int *fired; int stats[ODES_NUM_STATS];
  /* indefinite integration (opt->Indefinitely: one step of length opt->Time per call, no end; examples/ParameterScanner.c):
     the device computes ODES_LOOKAHEAD output steps ahead from the current state, on the grid the reference walks
     (tout += Time, src/integratorInstance.c:1228-1231); a run longer than that continues with a freshly initialised
     solver at the look-ahead boundary */
  const int indefinite = engine->opt->Indefinitely;
  cvodeSettings_t *planOpt = engine->opt;
  /* an instance keeps the linear multistep method its solver memory was first created with: the reference re-initialises
     that memory for later runs and cannot change the method then ("problem with ReInit: can't use new method",
     src/cvodeSolver.c:541-547) -- examples/simpleIntegratorInstance.c's second run announces BDF and integrates with
     Adams-Moulton (examples/testrun.out:281-310) */
  if (solver->lmmCreated == 0) solver->lmmCreated = 1 + engine->opt->CvodeMethod;
  if (indefinite || engine->opt->CvodeMethod != solver->lmmCreated - 1) {
    if (!solver->planOpt) {
      solver->planOpt = CvodeSettings_clone(engine->opt);
      solver->planOpt->CvodeMethod = solver->lmmCreated - 1;
      if (indefinite) {
        solver->planOpt->Indefinitely = 0;
        CvodeSettings_setTime(solver->planOpt, engine->opt->Time * ODES_LOOKAHEAD, ODES_LOOKAHEAD);
        solver->planOpt->StoreResults = 0;
      }
    }
    /* the clone is made once; what a launch reads from it follows the caller's settings, which may be changed in place
       between runs (CvodeSettings_setErrors and a reset: the ParameterScanner pattern) */
    solver->planOpt->Error = engine->opt->Error; solver->planOpt->RError = engine->opt->RError; solver->planOpt->Mxstep = engine->opt->Mxstep;
    planOpt = solver->planOpt;
  }
  if (!solver->plan) {
    int *idx = (int *)malloc((size_t)(nin > 0 ? nin : 1) * sizeof(int));
    for (i = 0, j = 0; i < nvalues; i++) if (i < om->neq || i >= om->neq + om->nass) idx[j++] = i;
    solver->plan = ODESBatch_create(om, planOpt, nin, idx, 0, NULL, 0);
    free(idx);
    if (!solver->plan) return 0;
  }
  nrem = indefinite ? ODES_LOOKAHEAD : solver->nout - solver->iout + 1;
  grid = (double *)malloc((size_t)(nrem + 1) * sizeof(double));
  grid[0] = solver->t;
  grid[1] = solver->tout;
  for (i = 2; i <= nrem; i++) grid[i] = indefinite ? grid[i - 1] + engine->opt->Time : engine->opt->TimePoints[solver->iout + i - 1];
  values = (double *)malloc((size_t)(nin > 0 ? nin : 1) * sizeof(double));
  for (i = 0, j = 0; i < nvalues; i++) if (i < om->neq || i >= om->neq + om->nass) values[j++] = data->value[i];
  rows = (double *)malloc((size_t)(nrem + 1) * (size_t)nvalues * sizeof(double));
  fired = (unsigned int *)malloc((size_t)((indefinite ? ODES_LOOKAHEAD : solver->nout) + 2) * sizeof(unsigned int));
  if (solver->nsens > 0) {                     /* the course continues from the current sensitivities (src/sensSolver.c:268-277) */
    double *s0 = (double *)malloc((size_t)solver->nsens * (size_t)om->neq * sizeof(double));
    int ok;
    for (j = 0; j < solver->nsens; j++) for (i = 0; i < om->neq; i++) s0[(size_t)j * (size_t)om->neq + (size_t)i] = data->sensitivity[i][j];
    ok = ODESBatch_setInitialSensitivities(solver->plan, s0);
    free(s0);
    if (!ok) { free(grid); free(values); free(rows); free(fired); return 0; }
  }
  ODESBatch_setEventProcessing(solver->plan, engine->processEvents);
  solver->courseEvents = engine->processEvents;
  if (!ODESBatch_setTimeGrid(solver->plan, nrem, grid) || !ODESBatch_setInitialTriggers(solver->plan, data->trigger) ||
      !ODESBatch_setValues(solver->plan, 1, values) || !ODESBatch_integrate(solver->plan, 1) || !ODESBatch_synchronize(solver->plan) ||
      !ODESBatch_getTrajectory(solver->plan, 0, rows) || !ODESBatch_getStatus(solver->plan, 1, &st) ||
      !ODESBatch_getEventLog(solver->plan, 1, fired) || !ODESBatch_getStats(solver->plan, 1, stats)) {
    free(grid); free(values); free(rows); free(fired);
    return 0;
  }
  drop_course(solver);
  if (solver->nsens > 0) {
    solver->courseSens = (double *)malloc((size_t)(planOpt->PrintStep + 1) * (size_t)solver->nsens * (size_t)om->neq * sizeof(double));
    if (!ODESBatch_getSensitivities(solver->plan, 1, solver->courseSens)) { free(grid); free(values); free(rows); free(fired); return 0; }
  }
  solver->course = rows; solver->courseFirst = solver->iout; solver->courseLen = nrem; solver->courseStatus = st;
  solver->courseFired = (int *)malloc((size_t)(nrem + 1) * sizeof(int));
  for (i = 0; i <= nrem; i++) solver->courseFired[i] = (int)fired[i];
  solver->nst += stats[ODES_STAT_NST]; solver->nfe += stats[ODES_STAT_NFE]; solver->nsetups += stats[ODES_STAT_NSETUPS];
  solver->nje += stats[ODES_STAT_NJE]; solver->nni += stats[ODES_STAT_NNI]; solver->ncfn += stats[ODES_STAT_NCFN]; solver->netf += stats[ODES_STAT_NETF];
  free(grid); free(values); free(fired);
  return 1;
}

int IntegratorInstance_updateData(integratorInstance_t *engine)
{
  int i, flag = 1; unsigned int fired = 0;
  cvodeSolver_t *solver = engine->solver; cvodeData_t *data = engine->data; cvodeSettings_t *opt = engine->opt;
  cvodeResults_t *results = engine->results; odeModel_t *om = engine->om;
  data->currenttime = (float)solver->t;
  /* events were processed on the device at this output time (event_f); pick up the outcome */
  if (engine->processEvents && solver->course) {
    int row = solver->iout - solver->courseFirst + 1;
    if (row >= 1 && row <= solver->courseLen) fired = (unsigned int)solver->courseFired[row];
    for (i = 0; i < data->nevents; i++) data->trigger[i] = evaluateAST(om->event[i], data) != 0.0;
    if (fired && opt->HaltOnEvent) {
      for (i = 0; i != data->nevents; i++)
        if (fired & (1u << i)) {
          char *buffer = SBML_formulaToString(om->event[i]);
          SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_EVENT_TRIGGER_FIRED,
                            "Event Trigger %d (%s) fired at time %g. Aborting simulation.", i, buffer, (double)data->currenttime);
          free(buffer);
        }
      flag = 0;
    }
  }
  data->allRulesUpdated = 1;                 /* device rows hold refreshed rules */
  if (opt->StoreResults && results) {
    results->nout = solver->iout;
    results->time[solver->iout] = solver->t;
    for (i = 0; i < data->nvalues; i++) results->value[i][solver->iout] = data->value[i];
    if (results->sensitivity && data->sensitivity) {
      int j;
      for (i = 0; i < results->neq; i++) for (j = 0; j < results->nsens; j++) results->sensitivity[i][j][solver->iout] = data->sensitivity[i][j];
    }
  }
  if (opt->SteadyState == 1 && IntegratorInstance_checkSteadyState(engine)) flag = 0;
  solver->iout++;
  if (opt->Indefinitely) solver->tout += opt->Time;
  else if (solver->iout <= solver->nout) solver->tout = opt->TimePoints[solver->iout];
  return flag;
}

static const char *cvode_message(int flag)
{
  switch (flag) {
  case -2: return "One of the inputs to CVode is illegal.";
  case -4: return "At time %g: The solver took %d internal steps but could not compute variable values.";
  case -5: return "The solver could not satisfy the accuracy requested for some internal step.";
  case -6: return "Error test failures occurred too many times during one internal time step or occurred with |h| = hmin.";
  case -7: return "Convergence test failures occurred too many times during one internal time step or occurred with |h| = hmin.";
  case -9: return "The linear solver's setup routine failed in an unrecoverable manner.";
  case -10: return "The linear solver's solve routine failed in an unrecoverable manner.";
  default: return "???";
  }
}

/* rows the device course did not reach (solver failure, halt) are not-a-number in EVERY value */
static int row_unreached(const double *row, int n)
{
  int i;
  for (i = 0; i < n; i++) if (!isnan(row[i])) return 0;
  return n > 0;
}
int IntegratorInstance_cvodeOneStep(integratorInstance_t *engine)
{
  cvodeSolver_t *solver = engine->solver; cvodeData_t *data = engine->data; int i, row;
  if (!engine->isValid || !solver->course || solver->courseEvents != engine->processEvents ||       /* the device course handled events the other way */
      (engine->opt->Indefinitely && solver->iout - solver->courseFirst + 1 > solver->courseLen)) {       /* look-ahead used up */
    solver->t0 = solver->t;
    if (!compute_course(engine)) return 0;
    engine->isValid = 1;
  }
  if (!engine->clockStarted) { engine->startTime = clock(); engine->clockStarted = 1; }
  row = solver->iout - solver->courseFirst + 1;
  if (row < 1 || row > solver->courseLen || row_unreached(solver->course + (size_t)row * (size_t)data->nvalues, data->nvalues)) {
    int flag = solver->courseStatus ? solver->courseStatus : -2;
    SolverError_error(ERROR_ERROR_TYPE, (errorCode_t)flag, cvode_message(flag), solver->tout, engine->opt->Mxstep);
    SolverError_error(WARNING_ERROR_TYPE, SOLVER_ERROR_INTEGRATION_NOT_SUCCESSFUL, "Integration not successful. Results are not complete.");
    return 0;
  }
  solver->t = solver->tout;
  for (i = 0; i < data->nvalues; i++) data->value[i] = solver->course[(size_t)row * (size_t)data->nvalues + (size_t)i];
  if (solver->courseSens && data->sensitivity) {             /* IntegratorInstance_getForwardSens (src/sensSolver.c:108-140) */
    int j, neq = engine->om->neq;
    for (j = 0; j < solver->nsens; j++) for (i = 0; i < neq; i++)
      data->sensitivity[i][j] = solver->courseSens[((size_t)row * (size_t)solver->nsens + (size_t)j) * (size_t)neq + (size_t)i];
  }
  /* a fired event with re-initialisation was already handled inside the device course */
  return IntegratorInstance_updateData(engine);
}
int IntegratorInstance_simpleOneStep(integratorInstance_t *engine) { return IntegratorInstance_cvodeOneStep(engine); }
int IntegratorInstance_integrateOneStep(integratorInstance_t *engine) { engine->processEvents = 1; return IntegratorInstance_cvodeOneStep(engine); }
int IntegratorInstance_integrateOneStepWithoutEventProcessing(integratorInstance_t *engine) { engine->processEvents = 0; return IntegratorInstance_cvodeOneStep(engine); }
int IntegratorInstance_integrate(integratorInstance_t *engine)
{
  while (engine->solver->iout <= engine->solver->nout) if (!IntegratorInstance_integrateOneStep(engine)) return 0;
  return 1;
}

cvodeResults_t *IntegratorInstance_createResults(const integratorInstance_t *engine)
{
  int i, j; cvodeResults_t *results, *iResults = engine->results;
  if (!engine->opt->StoreResults || iResults == NULL) return NULL;
  results = CvodeResults_create(engine->data, engine->opt->PrintStep);
  if (!results) return NULL;
  results->nout = iResults->nout;
  for (i = 0; i <= results->nout; i++) { results->time[i] = iResults->time[i]; for (j = 0; j < iResults->nvalues; j++) results->value[j][i] = iResults->value[j][i]; }
  if (iResults->sensitivity != NULL) {         /* src/integratorInstance.c:920-931 */
    int k;
    CvodeResults_allocateSens(results, iResults->neq, iResults->nsens, engine->opt->PrintStep);
    for (j = 0; j < iResults->nsens; j++) {
      results->index_sens[j] = iResults->index_sens[j];
      for (i = 0; i < iResults->neq; i++) for (k = 0; k <= results->nout; k++) results->sensitivity[i][j][k] = iResults->sensitivity[i][j][k];
    }
  }
  return results;
}
void IntegratorInstance_printResults(const integratorInstance_t *ii, FILE *fp)
{
  int n, j;
  fprintf(fp, "#t ");
  for (j = 0; j < ii->om->neq; j++) fprintf(fp, "%s ", ii->om->names[j]);
  fprintf(fp, "\n");
  for (n = 0; n < CvodeResults_getNout(ii->results); n++) {
    fprintf(fp, "%g ", CvodeResults_getTime(ii->results, n));
    for (j = 0; j < ii->om->neq; j++) fprintf(fp, "%g ", ii->results->value[j][n]);
    fprintf(fp, "\n");
  }
}
/* reference src/integratorInstance.c:1442-1507: mean of the |rates| plus the deviation of the SIGNED rates from that mean
   (d = rate_i - mean(|rate|), variance over neq - 1 -- for one equation that is 0/0, not-a-number, and no steady state is
   ever reported, exactly like the reference and the emitted device steady_f) below the threshold */
int IntegratorInstance_checkSteadyState(integratorInstance_t *engine)
{
  int i; double dy_mean = 0.0, dy_var = 0.0, dy_std; cvodeData_t *data = engine->data; odeModel_t *om = engine->om;
  if (om->neq == 0) { data->steadystate = 0; return 0; }
  if (!data->allRulesUpdated)
    for (i = 0; i < om->nassbeforeodes; i++) { nonzeroElem_t *o = om->assignmentsBeforeODEs[i]; data->value[o->i] = evaluateAST(o->ij, data); }
  for (i = 0; i < om->neq; i++) dy_mean += fabs(evaluateAST(om->ode[i], data));
  dy_mean = dy_mean / om->neq;
  for (i = 0; i < om->neq; i++) { const double d = evaluateAST(om->ode[i], data) - dy_mean; dy_var += d * d; }
  dy_var = dy_var / (om->neq - 1);
  dy_std = sqrt(dy_var);
  if ((dy_mean + dy_std) < engine->opt->ssThreshold) {
    data->steadystate = 1;
    if (engine->opt->SteadyState)
      SolverError_error(MESSAGE_ERROR_TYPE, SOLVER_MESSAGE_STEADYSTATE_FOUND, "Steady state found. Simulation aborted at %g seconds. Mean of rates: %g, std %g", (double)data->currenttime, dy_mean, dy_std);
    return 1;
  }
  data->steadystate = 0;
  return 0;
}
int IntegratorInstance_handleError(integratorInstance_t *engine)
{
  int errorCode = (int)SolverError_getLastCode(ERROR_ERROR_TYPE);
  (void)engine;
  return errorCode;
}
int IntegratorInstance_printCVODEStatistics(const integratorInstance_t *engine, FILE *f)
{
  const cvodeSolver_t *s = engine->solver;
  fprintf(f, "## CVode Statistics:\n");
  fprintf(f, "## nst = %-6ld nfe  = %-6ld nsetups = %-6ld nje = %ld\n", s->nst, s->nfe, s->nsetups, s->nje);
  fprintf(f, "## nni = %-6ld ncfn = %-6ld netf = %ld\n", s->nni, s->ncfn, s->netf);
  return 1;
}
/* src/sensSolver.c:680-720: the state's counters and, with sensitivities, a note on what the device counted */
int IntegratorInstance_printCVODESStatistics(const integratorInstance_t *engine, FILE *f)
{
  IntegratorInstance_printCVODEStatistics(engine, f);
  if (engine->os) fprintf(f, "##\n## CVode Sensitivity: %d sensitivity vectors iterated with the state (simultaneous or staggered corrector)\n", engine->os->nsens);
  return 1;
}
void IntegratorInstance_printStatistics(const integratorInstance_t *engine, FILE *f)
{
  odeModel_t *om = engine->om;
  fprintf(f, "## Integration Parameters:\n");
  fprintf(f, "## mxstep   = %d rel.err. = %g abs.err. = %g \n", engine->opt->Mxstep, engine->opt->RError, engine->opt->Error);
  fprintf(f, "## CVode Done.\n## Model Statistics:\n## ODEs = %d assignments = %d constants = %d\n", om->neq, om->nass, om->nconst);
  IntegratorInstance_printCVODEStatistics(engine, f);
}
double IntegratorInstance_getIntegrationTime(const integratorInstance_t *engine)
{ return engine->clockStarted ? ((double)(clock() - engine->startTime)) / CLOCKS_PER_SEC : 0.0; }

/* ---- smaller entries of the reference's integratorInstance.h ---- */
/* src/integratorInstance.c:508-527 */
int IntegratorInstance_setInitialTime(integratorInstance_t *engine, double initialtime)
{
  if ((engine->isValid == 0 && engine->solver->t == engine->solver->t0) && engine->solver->tout > initialtime) {
    engine->solver->t0 = initialtime; engine->solver->t = initialtime; engine->data->currenttime = (float)initialtime;
    return 1;
  }
  SolverError_error(WARNING_ERROR_TYPE, SOLVER_ERROR_ATTEMPTING_TO_SET_IMPOSSIBLE_INITIAL_TIME,
                    "Requested intial time (%f) is not possible! Reset integrator first, and make sure that the first output time (%f) "
                    "is smaller than the requested initial time! New setting ignored!", initialtime, engine->solver->tout);
  return 0;
}
const char *IntegratorInstance_getVariableName(integratorInstance_t *engine, variableIndex_t *vi) { return VariableIndex_getName(vi, engine->om); }
double *IntegratorInstance_getValues(integratorInstance_t *engine) { return engine->data->value; }
void IntegratorInstance_dumpNames(integratorInstance_t *engine) { ODEModel_dumpNames(engine->om); }
/* src/integratorInstance.c:1658-1690 (the CVODE statistics come from the device counters) */
void IntegratorInstance_dumpSolver(integratorInstance_t *engine)
{
  cvodeSolver_t *solver = engine->solver;
  printf("\nINTEGRATOR STATE:\n\nCurrent Time Settings:\n");
  printf("start time:          %g\ncurrent time:        %g\nnext time:           %g\n", solver->t0, solver->t, solver->tout);
  printf("current step number: %d\ntotal step number:   %d\n\n", solver->iout, solver->nout);
  if (engine->om->neq) {
    printf("CVODE Error Settings:\nabsolute error tolerance: %g\nrelative error tolerance: %g\nmax. internal step nr.:   %d\n\n",
           engine->opt->Error, engine->opt->RError, engine->opt->Mxstep);
    IntegratorInstance_printCVODEStatistics(engine, stdout);
  }
}
/* edge-triggered event check on the current values (src/integratorInstance.c:1371-1428) */
int IntegratorInstance_checkTrigger(integratorInstance_t *engine)
{
  int i, j, fired = 0; cvodeSettings_t *opt = engine->opt; cvodeData_t *data = engine->data; odeModel_t *om = engine->om;
  for (i = 0; i < om->nassbeforeevents; i++) { nonzeroElem_t *o = om->assignmentsBeforeEvents[i]; data->value[o->i] = evaluateAST(o->ij, data); }
  for (i = 0; i < data->nevents; i++) {
    if (data->trigger[i] == 0) {
      if (evaluateAST(om->event[i], data)) {
        if (opt->HaltOnEvent) {
          char *eq = SBML_formulaToString(om->event[i]);
          SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_EVENT_TRIGGER_FIRED, "Event Trigger %d (%s) fired at time %g. Aborting simulation.", i, eq, data->currenttime);
          free(eq);
        }
        fired++;
        data->trigger[i] = 1;
        for (j = 0; j < om->neventAss[i]; j++) set_value_by_index(engine, om->eventIndex[i][j], evaluateAST(om->eventAssignment[i][j], data));
      }
    } else if (!evaluateAST(om->event[i], data)) data->trigger[i] = 0;
  }
  return fired;
}
/* writes the last stored values back into the SBML model (src/integratorInstance.c:966-1003) */
int IntegratorInstance_updateModel(integratorInstance_t *engine)
{
  int i; odeModel_t *om = engine->om; cvodeResults_t *results = engine->results; Model_t *m = om->m;
  Species_t *s; Compartment_t *c; Parameter_t *p;
  if (!m || !results) return 0;
  for (i = 0; i < engine->data->nvalues; i++) {
    const double v = results->value[i][results->nout];
    if ((s = Model_getSpeciesById(m, om->names[i])) != NULL) {
      c = Model_getCompartmentById(m, s->compartment);
      if (!s->hasOnlySubstanceUnits && c && c->spatialDimensions != 0) { s->initialConcentration = v; s->isSetInitialConcentration = 1; s->isSetInitialAmount = 0; }
      else { s->initialAmount = v; s->isSetInitialAmount = 1; s->isSetInitialConcentration = 0; }
    }
    else if ((c = Model_getCompartmentById(m, om->names[i])) != NULL) { c->size = v; c->isSetSize = 1; }
    else if ((p = Model_getParameterById(m, om->names[i])) != NULL) { p->value = v; p->isSetValue = 1; }
    else return 0;
  }
  return 1;
}

/* ----------------------------------------------------------- forward sensitivities (src/integratorInstance.c:687-790) */
odeSense_t *IntegratorInstance_getSensitivityModel(integratorInstance_t *engine) { return engine->os; }
int IntegratorInstance_getNsens(integratorInstance_t *engine) { return engine->os ? engine->os->nsens : 0; }
double IntegratorInstance_getSensitivityByNum(integratorInstance_t *engine, int y, int p)
{
  if (!engine->os || !engine->data->sensitivity || y < 0 || y >= engine->om->neq || p < 0 || p >= engine->os->nsens) return 0.0;
  return engine->data->sensitivity[y][p];
}
double IntegratorInstance_getSensitivity(integratorInstance_t *engine, variableIndex_t *y, variableIndex_t *s)
{
  int i;
  if (!engine->os || !engine->data->sensitivity) { fprintf(stderr, "WARNING: sensitivity analysis has not been initialized\n"); return 0.0; }
  if (y->index >= engine->om->neq) {
    fprintf(stderr, "WARNING: ID is not a variable, no sensitivities can be calculated for %s \n", engine->om->names[y->index]);
    return 0.0;
  }
  for (i = 0; i < engine->os->nsens && engine->os->index_sens[i] != s->index; i++) ;
  if (i == engine->os->nsens) {
    fprintf(stderr, "WARNING: No sensitivities have been calculated for requested ID %s\n", engine->om->names[s->index]);
    return 0.0;
  }
  return engine->data->sensitivity[y->index][i];
}
char *IntegratorInstance_getSensVariableName(integratorInstance_t *engine, int i)
{ return (engine->os && i >= 0 && i < engine->os->nsens) ? engine->om->names[engine->os->index_sens[i]] : NULL; }
void IntegratorInstance_dumpYSensitivities(integratorInstance_t *engine, variableIndex_t *y)
{
  int j; cvodeData_t *data = engine->data;
  if (!data->sensitivity) return;
  if (y->index >= engine->om->neq) {
    printf("Warning: ID is not a variable, no sensitivities ");
    printf("can be calculated for %s \n", engine->om->names[y->index]);
    return;
  }
  printf("%g  ", data->currenttime);
  printf("%g  ", data->value[y->index]);
  for (j = 0; j < data->nsens; j++) printf("%g ", data->sensitivity[y->index][j]);
  printf("\n");
}
void IntegratorInstance_dumpPSensitivities(integratorInstance_t *engine, variableIndex_t *p)
{
  int i, j; cvodeData_t *data = engine->data;
  if (!engine->os || !data->sensitivity) return;
  for (j = 0; j < engine->os->nsens && engine->os->index_sens[j] != p->index; j++) ;
  if (j == engine->os->nsens) { printf("Warning: no sensitivity requested for ID %s\n", engine->om->names[p->index]); return; }
  printf("%g  ", data->currenttime);
  for (i = 0; i < engine->om->neq; i++) printf("%g ", data->sensitivity[i][j]);
  printf("\n");
}
