/*
 * settings.c -- cvodeSettings_t.  Defaults, setters and the output time grid
 * follow the reference exactly (src/integratorSettings.c:62-79 defaults:
 * atol 1e-18, rtol 1e-10, mxstep 10000, BDF, Newton, Jacobian on, results
 * stored; :128-176 createWith; :416-456 time grid i*EndTime/PrintStep).
 */
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sbmlsolver/integratorSettings.h"
#include "sbmlsolver/solverError.h"

cvodeSettings_t *CvodeSettings_create(void) { return CvodeSettings_createWithTime(1., 10); }
cvodeSettings_t *CvodeSettings_createWithTime(double Time, int PrintStep)
{ return CvodeSettings_createWith(Time, PrintStep, 1e-18, 1e-10, 10000, 0, 0, 1, 0, 0, 0, 1, 0, 0); }

cvodeSettings_t *CvodeSettings_createWith(double Time, int PrintStep, double Error, double RError, int Mxstep,
    int Method, int IterMethod, int UseJacobian, int Indefinitely, int HaltOnEvent, int SteadyState,
    int StoreResults, int Sensitivity, int SensMethod)
{
  cvodeSettings_t *set = (cvodeSettings_t *)calloc(1, sizeof *set);
  if (!set) return NULL;
  CvodeSettings_setErrors(set, Error, RError, Mxstep);
  set->CvodeMethod = Method; set->IterMethod = IterMethod;
  set->MaxOrder = (Method == 0) ? 5 : 12;
  set->compileFunctions = 0;
  set->ResetCvodeOnEvent = 1;
  set->StepResolution = 1;
  CvodeSettings_setSwitches(set, UseJacobian, Indefinitely, HaltOnEvent, SteadyState, StoreResults, Sensitivity, SensMethod);
  set->TimePoints = NULL;
  if (!Indefinitely) CvodeSettings_setTime(set, Time, PrintStep);
  else { set->Time = Time; set->PrintStep = PrintStep; }
  set->sensIDs = NULL; set->nsens = 0; set->DoAdjoint = 0; set->observation_data_type = 0;
  set->SetTStop = 0; set->DetectNegState = 0;
  set->ssThreshold = 1e-11;
  return set;
}
cvodeSettings_t *CvodeSettings_clone(cvodeSettings_t *set)
{
  cvodeSettings_t *c = (cvodeSettings_t *)malloc(sizeof *c);
  int i;
  *c = *set; c->TimePoints = NULL; c->sensIDs = NULL; c->nsens = 0;
  if (set->TimePoints) {
    c->TimePoints = (double *)malloc((size_t)(set->PrintStep + 1) * sizeof(double));
    for (i = 0; i <= set->PrintStep; i++) c->TimePoints[i] = set->TimePoints[i];
  }
  if (set->sensIDs) CvodeSettings_setSensParams(c, set->sensIDs, set->nsens);
  return c;
}
void CvodeSettings_unsetSensParams(cvodeSettings_t *set)
{
  int i;
  if (set->sensIDs) for (i = 0; i < set->nsens; i++) free(set->sensIDs[i]);
  free(set->sensIDs); set->sensIDs = NULL; set->nsens = 0;
}
int CvodeSettings_setSensParams(cvodeSettings_t *set, char **sensIDs, int nsens)
{
  int i;
  CvodeSettings_unsetSensParams(set);
  if (sensIDs) {
    set->sensIDs = (char **)calloc((size_t)(nsens > 0 ? nsens : 1), sizeof(char *));
    for (i = 0; i < nsens; i++) { set->sensIDs[i] = (char *)malloc(strlen(sensIDs[i]) + 1); strcpy(set->sensIDs[i], sensIDs[i]); }
    set->nsens = nsens;
  }
  return 1;
}
void CvodeSettings_free(cvodeSettings_t *set) { if (!set) return; CvodeSettings_unsetSensParams(set); free(set->TimePoints); free(set); }

int CvodeSettings_setTime(cvodeSettings_t *set, double EndTime, int PrintStep)
{
  int i, j; double *ts;
  assert(PrintStep > 0);
  ts = (double *)calloc((size_t)PrintStep, sizeof(double));
  if (EndTime == 0.0) PrintStep = 0;       /* only the initial row (reference :427-429) */
  for (i = 1; i <= PrintStep; i++) ts[i - 1] = i * EndTime / PrintStep;
  j = CvodeSettings_setTimeSeries(set, ts, PrintStep);
  free(ts);
  return j;
}
int CvodeSettings_setTimeSeries(cvodeSettings_t *set, double *timeseries, int PrintStep)
{
  int i;
  free(set->TimePoints);
  set->TimePoints = (double *)calloc((size_t)PrintStep + 1, sizeof(double));
  if (!set->TimePoints) return 0;
  set->Time = PrintStep > 0 ? timeseries[PrintStep - 1] : 0.0;
  set->PrintStep = PrintStep;
  set->TimePoints[0] = 0.0;
  for (i = 1; i <= PrintStep; i++) set->TimePoints[i] = timeseries[i - 1];
  return 1;
}
int CvodeSettings_setTimeStep(cvodeSettings_t *set, int i, double time)
{
  if (0 < i && i <= set->PrintStep) { set->TimePoints[i] = time; return 1; }
  return 0;
}
void CvodeSettings_setSwitches(cvodeSettings_t *set, int UseJacobian, int Indefinitely, int HaltOnEvent,
    int SteadyState, int StoreResults, int Sensitivity, int SensMethod)
{
  set->UseJacobian = UseJacobian; set->Indefinitely = Indefinitely; set->HaltOnEvent = HaltOnEvent;
  set->SteadyState = SteadyState; set->StoreResults = StoreResults;
  set->Sensitivity = Sensitivity; set->SensMethod = SensMethod;
}
void CvodeSettings_setErrors(cvodeSettings_t *s, double Error, double RError, int Mxstep) { s->Error = Error; s->RError = RError; s->Mxstep = Mxstep; }
void CvodeSettings_setError(cvodeSettings_t *s, double v) { s->Error = v; }
void CvodeSettings_setRError(cvodeSettings_t *s, double v) { s->RError = v; }
void CvodeSettings_setMxstep(cvodeSettings_t *s, int v) { s->Mxstep = v; }
void CvodeSettings_setDetectNegState(cvodeSettings_t *s, int v) { s->DetectNegState = v; }
void CvodeSettings_setStepResolution(cvodeSettings_t *s, int v) { s->StepResolution = (v == 0) ? 1 : v; }
int CvodeSettings_getStepResolution(const cvodeSettings_t *s) { return s->StepResolution; }
void CvodeSettings_setTStop(cvodeSettings_t *s, int v) { s->SetTStop = v; }
void CvodeSettings_setCompileFunctions(cvodeSettings_t *s, int v) { s->compileFunctions = v; }
void CvodeSettings_setMethod(cvodeSettings_t *s, int m, int maxord)
{
  if (0 <= m && m < 2) {
    s->CvodeMethod = m;
    if (m == 0) s->MaxOrder = (maxord > 5) ? 5 : maxord;      /* BDF <= 5, Adams <= 12 */
    else s->MaxOrder = (maxord > 12) ? 12 : maxord;
  }
}
void CvodeSettings_setIterMethod(cvodeSettings_t *s, int i) { if (0 <= i && i < 2) s->IterMethod = i; }
void CvodeSettings_setMaxOrder(cvodeSettings_t *s, int o) { CvodeSettings_setMethod(s, s->CvodeMethod, o); }
void CvodeSettings_setJacobian(cvodeSettings_t *s, int v) { s->UseJacobian = v; }
void CvodeSettings_setIndefinitely(cvodeSettings_t *s, int v) { s->Indefinitely = v; }
void CvodeSettings_setHaltOnEvent(cvodeSettings_t *s, int v) { s->HaltOnEvent = v; }
void CvodeSettings_setResetCvodeOnEvent(cvodeSettings_t *s, int v) { s->ResetCvodeOnEvent = v; }
void CvodeSettings_setHaltOnSteadyState(cvodeSettings_t *s, int v) { s->SteadyState = v; }
void CvodeSettings_setSteadyStateThreshold(cvodeSettings_t *s, double v) { s->ssThreshold = v; }
void CvodeSettings_setStoreResults(cvodeSettings_t *s, int v) { s->StoreResults = v; }
void CvodeSettings_setSensitivity(cvodeSettings_t *s, int v) { s->Sensitivity = v; }
void CvodeSettings_setSensMethod(cvodeSettings_t *s, int v) { if (0 <= v && v < 3) s->SensMethod = v; }
double CvodeSettings_getEndTime(cvodeSettings_t *s) { return s->Indefinitely ? s->Time : s->TimePoints[s->PrintStep]; }
int CvodeSettings_getPrintsteps(cvodeSettings_t *s) { return s->PrintStep; }
double CvodeSettings_getTimeStep(cvodeSettings_t *s) { return s->Indefinitely ? s->Time : s->TimePoints[1]; }
double CvodeSettings_getTime(cvodeSettings_t *s, int i) { return s->Indefinitely ? i * s->Time : s->TimePoints[i]; }
double CvodeSettings_getError(cvodeSettings_t *s) { return s->Error; }
double CvodeSettings_getRError(cvodeSettings_t *s) { return s->RError; }
int CvodeSettings_getMxstep(cvodeSettings_t *s) { return s->Mxstep; }
const char *CvodeSettings_getMethod(const cvodeSettings_t *s) { static const char *m[] = { "BDF", "ADAMS-MOULTON" }; return m[s->CvodeMethod]; }
const char *CvodeSettings_getIterMethod(const cvodeSettings_t *s) { static const char *m[] = { "NEWTON", "FUNCTIONAL" }; return m[s->IterMethod]; }
int CvodeSettings_getMaxOrder(cvodeSettings_t *s) { return s->MaxOrder; }
int CvodeSettings_getCompileFunctions(cvodeSettings_t *s) { return s->compileFunctions; }
int CvodeSettings_getResetCvodeOnEvent(cvodeSettings_t *s) { return s->ResetCvodeOnEvent; }
int CvodeSettings_getJacobian(cvodeSettings_t *s) { return s->UseJacobian; }
int CvodeSettings_getIndefinitely(cvodeSettings_t *s) { return s->Indefinitely; }
int CvodeSettings_getHaltOnEvent(cvodeSettings_t *s) { return s->HaltOnEvent; }
int CvodeSettings_getHaltOnSteadyState(cvodeSettings_t *s) { return s->SteadyState; }
double CvodeSettings_getSteadyStateThreshold(cvodeSettings_t *s) { return s->ssThreshold; }
int CvodeSettings_getStoreResults(cvodeSettings_t *s) { return s->StoreResults; }
int CvodeSettings_getSensitivity(cvodeSettings_t *s) { return s->Sensitivity; }
const char *CvodeSettings_getSensMethod(const cvodeSettings_t *s) { static const char *m[] = { "simultaneous", "staggered", "staggered1" }; return m[s->SensMethod]; }
/* the reference's printout, line for line (src/integratorSettings.c:83-127; examples/testrun.out:246-266) */
void CvodeSettings_dump(cvodeSettings_t *set)
{
  printf("\nSOSlib INTEGRATION SETTINGS\n1) CVODE SPECIFIC SETTINGS:\n");
  printf("absolute error tolerance for each output time:   %g\n", set->Error);
  printf("relative error tolerance for each output time:   %g\n", set->RError);
  printf("max. nr. of steps to reach next output time:     %d\n", set->Mxstep);
  printf("Nonlinear solver method:                         %d: %s\n          Maximum Order:                         %d\n",
         set->CvodeMethod, CvodeSettings_getMethod(set), set->MaxOrder);
  printf("Iteration method:                                %d: %s\n", set->IterMethod, CvodeSettings_getIterMethod(set));
  printf("Sensitivity:                                     %s\n", set->Sensitivity ? "1: yes " : "0: no");
  printf("     method:                                     %d: %s\n", set->SensMethod, CvodeSettings_getSensMethod(set));
  printf("2) SOSlib SPECIFIC SETTINGS:\nJacobian matrix: %s\n", set->UseJacobian ? "1: generate Jacobian" : "0: CVODE's internal approximation");
  printf("Indefinitely:    %s\n", set->Indefinitely ? "1: infinite integration" : "0: finite integration");
  printf("Event Handling:  %s\n", set->HaltOnEvent ? "1: stop integration" : "0: keep integrating");
  printf("Steady States:   %s\n", set->SteadyState ? "1: stop integrating" : "0: keep integrating");
  printf("Steady state threshold: %g\n", set->ssThreshold);
  printf("Store Results:   %s\n", set->StoreResults ? "1: store results (only for finite integration)" : "0: don't store results");
  printf("3) TIME SETTINGS:\n");
  if (set->Indefinitely) printf("Infinite integration with time step %g", set->Time);
  else printf("endtime: %g\nsteps:   %d", set->TimePoints[set->PrintStep], set->PrintStep);
  printf("\n\n");
}

/* src/integratorSettings.c:1153-1183; tmult keeps the reference's sign, (t0 - tend) / nout (unittest/test_integratorSettings.c:7-15) */
timeSettings_t *TimeSettings_create(double t0, double tend, int nout)
{
  timeSettings_t *ts;
  assert(nout > 0);
  ts = (timeSettings_t *)calloc(1, sizeof *ts);
  ts->t0 = t0; ts->tend = tend; ts->nout = nout; ts->tmult = (t0 - tend) / nout;
  return ts;
}
void TimeSettings_free(timeSettings_t *ts) { free(ts); }
cvodeSettings_t *CvodeSettings_createFromTimeSettings(timeSettings_t *ts) { return CvodeSettings_createWithTime(ts->tend, ts->nout); }
