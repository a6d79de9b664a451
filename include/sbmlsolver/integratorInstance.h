/*
 * integratorInstance.h -- integratorInstance_t and its drivers.
 * Reference: src/sbmlsolver/integratorInstance.h:52-201 and
 * src/sbmlsolver/cvodeSolver.h:54-62.  The solver seam
 * (IntegratorInstance_cvodeOneStep) keeps its contract -- advance solver->t
 * to solver->tout, write data->value[0..neq), call
 * IntegratorInstance_updateData -- but the time stepping itself runs on the
 * GPU through the batched kernel (batchIntegrator.h) with a batch of one.
 */
#ifndef ODES_B200_INTEGRATORINSTANCE_H
#define ODES_B200_INTEGRATORINSTANCE_H
#include <stdio.h>
#include <time.h>
#include "sbmlsolver/cvodeData.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct cvodeSolver cvodeSolver_t;
typedef struct integratorInstance integratorInstance_t;
struct odesBatch;

struct cvodeSolver {
  double t0, t, tout; int nout, iout;
  double reltol, atol1;
  /* device-side integration of the remaining time course (batch of one) */
  struct odesBatch *plan;
  unsigned long planSignature; /* of the settings the plan was specialised for */
  int lmmCreated;              /* 0: no solver yet, else 1 + CvodeMethod of the first run (kept for the life of the instance) */
  cvodeSettings_t *planOpt;   /* indefinite integration: the finite look-ahead grid the plan was created with */
  double *course;      /* [nsteps][nvalues] rows computed ahead by the device */
  int *courseFired;    /* [nsteps] event bitmask per row                       */
  int courseFirst, courseLen, courseStatus, courseFailRow;
  int courseEvents;    /* the course was computed with (1) or without (0) event processing */
  double *courseSens;  /* [nsteps][nsens][neq] forward sensitivities of the course, or NULL */
  int nsens;
  long nst, nfe, nsetups, nje, nni, ncfn, netf;
};

struct integratorInstance {
  int isValid;
  odeModel_t *om; cvodeSettings_t *opt; cvodeData_t *data; cvodeSolver_t *solver; cvodeResults_t *results;
  int run; int UseJacobian; int processEvents;
  int clockStarted; clock_t startTime;
  odeSense_t *os;      /* forward-sensitivity structures (opt->Sensitivity), owned by the instance */
};

integratorInstance_t *IntegratorInstance_create(odeModel_t *, cvodeSettings_t *);
int IntegratorInstance_set(integratorInstance_t *, cvodeSettings_t *);
int IntegratorInstance_reset(integratorInstance_t *);
int IntegratorInstance_resetTime(integratorInstance_t *, cvodeSettings_t *);
void IntegratorInstance_free(integratorInstance_t *);
cvodeSettings_t *IntegratorInstance_getSettings(integratorInstance_t *);
void IntegratorInstance_copyVariableState(integratorInstance_t *target, integratorInstance_t *source);
double IntegratorInstance_getTime(const integratorInstance_t *);
int IntegratorInstance_setNextTimeStep(integratorInstance_t *, double);
double IntegratorInstance_getVariableValue(integratorInstance_t *, variableIndex_t *);
void IntegratorInstance_setVariableValue(integratorInstance_t *, variableIndex_t *, double);
void IntegratorInstance_setVariableValueByID(integratorInstance_t *, const char *, double);
void IntegratorInstance_dumpData(const integratorInstance_t *);
cvodeData_t *IntegratorInstance_getData(integratorInstance_t *);
int IntegratorInstance_integrate(integratorInstance_t *);
int IntegratorInstance_integrateOneStep(integratorInstance_t *);
int IntegratorInstance_integrateOneStepWithoutEventProcessing(integratorInstance_t *);
int IntegratorInstance_simpleOneStep(integratorInstance_t *);
int IntegratorInstance_cvodeOneStep(integratorInstance_t *);
int IntegratorInstance_updateData(integratorInstance_t *);
int IntegratorInstance_timeCourseCompleted(const integratorInstance_t *);
const cvodeResults_t *IntegratorInstance_getResults(const integratorInstance_t *);
cvodeResults_t *IntegratorInstance_createResults(const integratorInstance_t *);
void IntegratorInstance_printResults(const integratorInstance_t *, FILE *);
int IntegratorInstance_checkSteadyState(integratorInstance_t *);
int IntegratorInstance_handleError(integratorInstance_t *);
void IntegratorInstance_printStatistics(const integratorInstance_t *, FILE *);
int IntegratorInstance_printCVODEStatistics(const integratorInstance_t *, FILE *);   /* 1, like the reference (cvodeSolver.h:55) */
double IntegratorInstance_getIntegrationTime(const integratorInstance_t *);
int IntegratorInstance_setInitialTime(integratorInstance_t *, double);
const char *IntegratorInstance_getVariableName(integratorInstance_t *, variableIndex_t *);
double *IntegratorInstance_getValues(integratorInstance_t *);
void IntegratorInstance_dumpNames(integratorInstance_t *);
void IntegratorInstance_dumpSolver(integratorInstance_t *);
int IntegratorInstance_checkTrigger(integratorInstance_t *);
int IntegratorInstance_updateModel(integratorInstance_t *);
/* forward sensitivities (reference integratorInstance.h:180-187, src/integratorInstance.c:687-790) */
odeSense_t *IntegratorInstance_getSensitivityModel(integratorInstance_t *);
double IntegratorInstance_getSensitivity(integratorInstance_t *, variableIndex_t *y, variableIndex_t *p);
double IntegratorInstance_getSensitivityByNum(integratorInstance_t *, int y, int p);
int IntegratorInstance_getNsens(integratorInstance_t *);
char *IntegratorInstance_getSensVariableName(integratorInstance_t *, int);
void IntegratorInstance_dumpYSensitivities(integratorInstance_t *, variableIndex_t *);
void IntegratorInstance_dumpPSensitivities(integratorInstance_t *, variableIndex_t *);

#ifdef __cplusplus
}
#endif
#endif
