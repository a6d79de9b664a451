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
  /* forward sensitivities (reference cvodeData.h:178-184): sensitivity[i][j][k] = d value_i(t_k) / d p_j */
  int neq, nsens; int *index_sens; double ***sensitivity;
  double **directional;  /* [neq][nout+1] sum_j sensitivity[i][j][k] * dp[j] (CvodeResults_computeDirectional) */
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
  /* current sensitivities d value_i / d p_j, [neq][nsens] (reference cvodeData.h:95-122) */
  odeSense_t *os; int nsens; double **sensitivity;
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
/* reference cvodeData.h:212-226 */
double CvodeResults_getSensitivityByNum(cvodeResults_t *, int value, int parameter, int timestep);
double CvodeResults_getSensitivity(cvodeResults_t *, variableIndex_t *y, variableIndex_t *p, int timestep);
void CvodeResults_computeDirectional(cvodeResults_t *, const double *dp);
int CvodeResults_allocateSens(cvodeResults_t *, int neq, int nsens, int nout);
int CvodeData_initializeSensitivities(cvodeData_t *, cvodeSettings_t *, odeModel_t *, odeSense_t *);

#ifdef __cplusplus
}
#endif
#endif
