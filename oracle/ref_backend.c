/*
 * ref_backend.c -- ORACLE (test infrastructure, not product code).
 * Adapter that exposes the reference's vendored, UNMODIFIED SUNDIALS 2.1.1
 * CVODE 2.3.0 (compiled from /root/reference/Win32/WinCVODE/sundials where the
 * sources lie, see oracle/Makefile) through the neutral backend interface of
 * oracle.h.  Set-up mirrors the reference's own solver creation,
 * src/cvodeSolver.c:412-703: BDF + Newton, CVodeMalloc/CVodeReInit at t0 with
 * scalar rtol and a vector atol of equal entries (CV_SV), f_data, max steps,
 * CVDense(neq), optional user Jacobian.
 */
#include <stdlib.h>
#include "cvode.h"
#include "cvdense.h"
#include "nvector_serial.h"
#include "oracle.h"

/* linear multistep method of solvers created from now on: 0 BDF (the default), 1 Adams-Moulton (cvodeSettings_t.CvodeMethod);
   set by the test helper before a run */
int refcv_adams = 0;
/* ... and their iteration: 0 Newton (the default), 1 functional (cvodeSettings_t.IterMethod) */
int refcv_functional = 0;

typedef struct {
  int n; oracle_rhs_fn f; oracle_jac_fn jac; void *user;
  void *mem; N_Vector y, abstol; int created; long tot[7];
} ref_t;

static void f_cb(realtype t, N_Vector y, N_Vector ydot, void *f_data)
{ ref_t *r = (ref_t *)f_data; r->f(t, NV_DATA_S(y), NV_DATA_S(ydot), r->user); }
static void jac_cb(long int N, DenseMat J, realtype t, N_Vector y, N_Vector fy, void *jac_data,
                   N_Vector t1, N_Vector t2, N_Vector t3)
{ ref_t *r = (ref_t *)jac_data; (void)N; (void)t1; (void)t2; (void)t3; r->jac(t, NV_DATA_S(y), NV_DATA_S(fy), J->data[0], r->user); }

void *refcv_create(int n, oracle_rhs_fn f, oracle_jac_fn jac, void *user)
{
  ref_t *r = (ref_t *)calloc(1, sizeof *r);
  r->n = n; r->f = f; r->jac = jac; r->user = user;
  r->y = N_VNew_Serial(n); r->abstol = N_VNew_Serial(n);
  return r;
}
static void fold(ref_t *r)
{
  long v; int q;
  CVodeGetNumSteps(r->mem, &v); r->tot[0] += v;
  CVodeGetNumRhsEvals(r->mem, &v); r->tot[1] += v;
  CVodeGetNumLinSolvSetups(r->mem, &v); r->tot[2] += v;
  if (!refcv_functional) { CVDenseGetNumJacEvals(r->mem, &v); r->tot[3] += v; }      /* functional iteration never sets the linear solver up: its counter is not initialised */
  CVodeGetNumNonlinSolvIters(r->mem, &v); r->tot[4] += v;
  CVodeGetNumNonlinSolvConvFails(r->mem, &v); r->tot[5] += v;
  CVodeGetNumErrTestFails(r->mem, &v); r->tot[6] += v;
  (void)q;
}
int refcv_init(void *h, double t0, const double *y0, double reltol, double abstol, long mxstep, int reinit)
{
  ref_t *r = (ref_t *)h; int i, flag;
  for (i = 0; i < r->n; i++) { NV_Ith_S(r->y, i) = y0[i]; NV_Ith_S(r->abstol, i) = abstol; }
  if (!r->created) {
    r->mem = CVodeCreate(refcv_adams ? CV_ADAMS : CV_BDF, refcv_functional ? CV_FUNCTIONAL : CV_NEWTON);
    if (!r->mem) return -1;
    CVodeSetErrFile(r->mem, NULL);
    flag = CVodeMalloc(r->mem, f_cb, t0, r->y, CV_SV, reltol, r->abstol);
    r->created = 1;
    for (i = 0; i < 7; i++) r->tot[i] = 0;
  } else {
    if (reinit) fold(r); else for (i = 0; i < 7; i++) r->tot[i] = 0;
    flag = CVodeReInit(r->mem, f_cb, t0, r->y, CV_SV, reltol, r->abstol);
  }
  if (flag != CV_SUCCESS) return flag;
  CVodeSetFdata(r->mem, r);
  CVodeSetMaxNumSteps(r->mem, mxstep);
  flag = CVDense(r->mem, r->n);
  if (flag != CVDENSE_SUCCESS) return -8;
  CVDenseSetJacFn(r->mem, r->jac ? jac_cb : NULL, r);
  return 0;
}
void refcv_set_stop_time(void *h, double tstop) { ref_t *r = (ref_t *)h; CVodeSetStopTime(r->mem, tstop); }
int refcv_solve(void *h, double tout, double *yout, double *tret, int use_tstop)
{
  ref_t *r = (ref_t *)h; int i, flag; realtype t = 0;
  flag = CVode(r->mem, tout, r->y, &t, use_tstop ? CV_NORMAL_TSTOP : CV_NORMAL);
  for (i = 0; i < r->n; i++) yout[i] = NV_Ith_S(r->y, i);
  *tret = t;
  return flag;
}
/* one internal step (CV_ONE_STEP), the dense output and the step geometry: what the event-location procedure of
   driver.c (run_rootfind) drives the vendored CVODE with */
int refcv_step(void *h, double tout, double *yout, double *tret)
{
  ref_t *r = (ref_t *)h; int i, flag; realtype t = 0;
  flag = CVode(r->mem, tout, r->y, &t, CV_ONE_STEP);
  for (i = 0; i < r->n; i++) yout[i] = NV_Ith_S(r->y, i);
  *tret = t;
  return flag;
}
int refcv_dky(void *h, double t, double *yout)
{
  ref_t *r = (ref_t *)h; int i, flag;
  flag = CVodeGetDky(r->mem, t, 0, r->y);
  for (i = 0; i < r->n; i++) yout[i] = NV_Ith_S(r->y, i);
  return flag;
}
void refcv_step_info(void *h, double *tn, double *hu)
{
  ref_t *r = (ref_t *)h; realtype v = 0;
  CVodeGetCurrentTime(r->mem, &v); *tn = v;
  CVodeGetLastStep(r->mem, &v); *hu = v;
}
void refcv_stats(void *h, long *s)
{
  ref_t *r = (ref_t *)h; long save[7]; int i;
  for (i = 0; i < 7; i++) save[i] = r->tot[i];
  fold(r);
  for (i = 0; i < 7; i++) { s[i] = r->tot[i]; r->tot[i] = save[i]; }
}
void refcv_free(void *h)
{
  ref_t *r = (ref_t *)h;
  if (!r) return;
  if (r->created) CVodeFree(r->mem);
  N_VDestroy_Serial(r->y); N_VDestroy_Serial(r->abstol); free(r);
}
