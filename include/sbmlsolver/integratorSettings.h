/*
 * integratorSettings.h -- cvodeSettings_t, field-for-field the reference's
 * src/sbmlsolver/integratorSettings.h:56-148 for the forward path (forward
 * sensitivities with the simultaneous or the staggered corrector are honoured by the batched kernel; the adjoint switches are
 * inert fields).
 * Functions: reference integratorSettings.h:172-245.
 */
#ifndef ODES_B200_INTEGRATORSETTINGS_H
#define ODES_B200_INTEGRATORSETTINGS_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cvodeSettings cvodeSettings_t;
struct cvodeSettings {
  double Time; int PrintStep; int StepResolution; double *TimePoints; int Indefinitely;
  double Error; double RError; int Mxstep; int DetectNegState;
  int CvodeMethod; int IterMethod; int MaxOrder; int ResetCvodeOnEvent; int SetTStop;
  int Sensitivity; char **sensIDs; int nsens; int SensMethod;
  int HaltOnEvent; int SteadyState; double ssThreshold;
  int UseJacobian; int StoreResults; int compileFunctions;
  int observation_data_type; int DoAdjoint;
  /* adjoint / discrete-data / FIM tail of the reference's struct (integratorSettings.h:118-148): inert here -- kept so that
     sizeof(struct cvodeSettings) and every field offset are the reference's for callers that allocate or copy the struct */
  double AdjTime; int AdjPrintStep; double *AdjTimePoints;
  int nSaveSteps; int ncheck;
  double AdjError; double AdjRError; int AdjStoreResults;
  int OffSet; int InterStep;
  int doFIM;
};

/* reference integratorSettings.h:150-169 ("not used currently"): start and end time, step, number of steps */
typedef struct timeSettings { double t0; double tmult; double tend; int nout; } timeSettings_t;
timeSettings_t *TimeSettings_create(double t0, double tend, int nout);
void TimeSettings_free(timeSettings_t *);
cvodeSettings_t *CvodeSettings_createFromTimeSettings(timeSettings_t *);

cvodeSettings_t *CvodeSettings_create(void);
cvodeSettings_t *CvodeSettings_createWithTime(double Time, int PrintStep);
cvodeSettings_t *CvodeSettings_createWith(double EndTime, int PrintStep, double Error, double RError,
    int Mxstep, int Method, int IterMethod, int UseJacobian, int Indefinitely, int HaltOnEvent,
    int SteadyState, int StoreResults, int Sensitivity, int SensMethod);
int CvodeSettings_setTime(cvodeSettings_t *, double EndTime, int PrintStep);
int CvodeSettings_setTimeStep(cvodeSettings_t *, int, double);
int CvodeSettings_setTimeSeries(cvodeSettings_t *, double *timeseries, int PrintStep);
void CvodeSettings_setSwitches(cvodeSettings_t *, int UseJacobian, int Indefinitely, int HaltOnEvent,
    int SteadyState, int StoreResults, int Sensitivity, int SensMethod);
void CvodeSettings_setErrors(cvodeSettings_t *, double Error, double RError, int Mxstep);
void CvodeSettings_setError(cvodeSettings_t *, double);
void CvodeSettings_setRError(cvodeSettings_t *, double);
void CvodeSettings_setMxstep(cvodeSettings_t *, int);
void CvodeSettings_setDetectNegState(cvodeSettings_t *, int);
/* cvodeSettings_t.StepResolution (reference integratorSettings.h:61-63: a documented field without a setter): 1 events at the
   output times (the reference's behaviour); k > 1: k sub-steps per print step; k < 0: events located inside the internal steps
    */
void CvodeSettings_setStepResolution(cvodeSettings_t *, int);
int CvodeSettings_getStepResolution(const cvodeSettings_t *);
void CvodeSettings_setTStop(cvodeSettings_t *, int);
void CvodeSettings_setCompileFunctions(cvodeSettings_t *, int);
void CvodeSettings_setMethod(cvodeSettings_t *, int, int);
void CvodeSettings_setIterMethod(cvodeSettings_t *, int);
void CvodeSettings_setMaxOrder(cvodeSettings_t *, int);
void CvodeSettings_setJacobian(cvodeSettings_t *, int);
void CvodeSettings_setIndefinitely(cvodeSettings_t *, int);
void CvodeSettings_setHaltOnEvent(cvodeSettings_t *, int);
void CvodeSettings_setResetCvodeOnEvent(cvodeSettings_t *, int);
void CvodeSettings_setHaltOnSteadyState(cvodeSettings_t *, int);
void CvodeSettings_setSteadyStateThreshold(cvodeSettings_t *, double);
void CvodeSettings_setStoreResults(cvodeSettings_t *, int);
void CvodeSettings_setSensitivity(cvodeSettings_t *, int);
void CvodeSettings_setSensMethod(cvodeSettings_t *, int);
/* reference src/integratorSettings.c:877-908: ids of the parameters / variables whose sensitivities are wanted;
   NULL selects the default (all constants of the model) */
int CvodeSettings_setSensParams(cvodeSettings_t *, char **sensIDs, int nsens);
void CvodeSettings_unsetSensParams(cvodeSettings_t *);
void CvodeSettings_dump(cvodeSettings_t *);
void CvodeSettings_free(cvodeSettings_t *);
cvodeSettings_t *CvodeSettings_clone(cvodeSettings_t *);
double CvodeSettings_getEndTime(cvodeSettings_t *);
int CvodeSettings_getPrintsteps(cvodeSettings_t *);
double CvodeSettings_getTimeStep(cvodeSettings_t *);
double CvodeSettings_getTime(cvodeSettings_t *, int);
double CvodeSettings_getError(cvodeSettings_t *);
double CvodeSettings_getRError(cvodeSettings_t *);
int CvodeSettings_getMxstep(cvodeSettings_t *);
const char *CvodeSettings_getMethod(const cvodeSettings_t *);
const char *CvodeSettings_getIterMethod(const cvodeSettings_t *);
int CvodeSettings_getMaxOrder(cvodeSettings_t *);
int CvodeSettings_getCompileFunctions(cvodeSettings_t *);
int CvodeSettings_getResetCvodeOnEvent(cvodeSettings_t *);
int CvodeSettings_getJacobian(cvodeSettings_t *);
int CvodeSettings_getIndefinitely(cvodeSettings_t *);
int CvodeSettings_getHaltOnEvent(cvodeSettings_t *);
int CvodeSettings_getHaltOnSteadyState(cvodeSettings_t *);
double CvodeSettings_getSteadyStateThreshold(cvodeSettings_t *);
int CvodeSettings_getStoreResults(cvodeSettings_t *);
int CvodeSettings_getSensitivity(cvodeSettings_t *);
const char *CvodeSettings_getSensMethod(const cvodeSettings_t *);

#ifdef __cplusplus
}
#endif
#endif
