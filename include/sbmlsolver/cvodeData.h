/*
 * cvodeData.h -- per-instance integration data and stored results.
 * Reference: src/sbmlsolver/cvodeData.h:92-215.
 */
#ifndef ODES_B200_CVODEDATA_H
#define ODES_B200_CVODEDATA_H
#include "sbmlsolver/odeModel.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct cvodeData cvodeData_t;
typedef struct cvodeResults cvodeResults_t;

struct cvodeResults {
  int nout;            /* last stored time index */
  double *time;        /* [nout+1] */
  int nvalues;
  double **value;      /* [nvalues][nout+1], one time row per variable (cvodeData.c:596-613) */
  int neq, nsens; double ***sensitivity;   /* unused on this path, kept NULL */
};

struct cvodeData {
  odeModel_t *model;
  int nvalues; double *value;
  float currenttime;              /* float, as in the reference (cvodeData.h:116) */
  int allRulesUpdated;
  int nevents; int *trigger;
  int steadystate;
  int neq;
  cvodeSettings_t *opt;
  cvodeResults_t *results;
  int run;
};

cvodeData_t *CvodeData_create(odeModel_t *);
void CvodeData_initializeValues(cvodeData_t *);
int CvodeData_initialize(cvodeData_t *, cvodeSettings_t *, odeModel_t *, int keepValues);
void CvodeData_free(cvodeData_t *);
cvodeResults_t *CvodeResults_create(cvodeData_t *, int nout);
void CvodeResults_free(cvodeResults_t *);
double CvodeResults_getTime(const cvodeResults_t *, int);
double CvodeResults_getValue(cvodeResults_t *, variableIndex_t *, int);
int CvodeResults_getNout(const cvodeResults_t *);

#ifdef __cplusplus
}
#endif
#endif
