/*
 * sens_ref.c -- ORACLE (test infrastructure, not product code).
 *
 * Forward sensitivity analysis with the reference's vendored, UNMODIFIED SUNDIALS 2.1.1 CVODES 2.3.0 (compiled from
 * /root/reference/Win32/WinCVODE/sundials/cvodes where the sources lie, see oracle/Makefile) set up the way the
 * reference sets it up:
 *   src/sensSolver.c:197-346   simultaneous corrector (SensMethod 0), analytic sensitivity right-hand side, one
 *                              sensitivity at a time (CVodeSensInit1 / CVSensRhs1Fn), CVodeSetSensParams(p, NULL, NULL),
 *                              estimated tolerances (CV_EE), CVodeSetSensErrCon(FALSE)
 *   src/cvodeSolver.c:412-703  BDF + Newton, scalar rtol, vector atol of equal entries, CVDense + analytic Jacobian
 *   src/sensSolver.c:108-140   CVodeGetSens at the output time after every CVode call
 *   src/cvodeData.c:453-461    initial sensitivities: 1 for a variable with respect to its own initial value, else 0
 *   src/odeSolver.c:313-330    the design-point loop: reset, setVariableValue x nvary, integrate
 * The model functions are the emitted host-C flavour (ODEModel_generateCSourceSens: ode_f, jacobi_f, sense_f = the
 * reference's compiled mode, src/odeModel.c:2673-2918, :2994-3100) loaded from a shared object.
 * There is no restated port behind this file: for the sensitivity row of SURVEY section 8(f) the checker is the
 * reference's own CVODES arithmetic.
 * Events (optional, om != NULL): examined at output times only, edge-triggered, in model order, exactly as in
 * oracle/driver.c (src/integratorInstance.c:1029-1077); a fired event that changes a variable (or any value with
 * ResetCvodeOnEvent) invalidates the solver, which is then re-initialised at the output time from the current values
 * AND the current sensitivities (src/sensSolver.c:268-277: "yS are 0.0 or 1.0 in a new run or old values in case of
 * events") with CVodeReInit + CVodeSensReInit.
 * Difference quotients (refcvs_sens_dq / refcvs_jac_dq, set by the test helper): src/sensSolver.c:312-321 -- without a
 * parametric matrix the reference registers no sensitivity right-hand side (CVODES' internal CVSensRhsDQ, centred,
 * rhomax 0) and lets its f read the sensitivity parameters from the array CVODES perturbs (data->use_p,
 * src/cvodeSolver.c:981-984); without a Jacobian CVDENSE's CVDenseDQJac is used (src/cvodeSolver.c:620-623).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "cvodes.h"
#include "cvdense.h"
#include "nvector_serial.h"
#include "oracle.h"            /* odeModel_t and the independent AST interpreter, for the output-time event handling */

typedef void (*c_ode_fn)(double t, const double *y, double *ydot, double *value);
typedef void (*c_jac_fn)(double t, const double *y, double *J, double *value);
typedef void (*c_sens_fn)(int iS, double t, const double *y, const double *yS, double *ySdot, double *value);

/* SensMethod of the next batch calls (0 simultaneous, 1 staggered, 2 staggered1); set by the test helper before a run */
int refcvs_sens_method = 0;
/* 1: no sensitivity right-hand side is registered (CVSensRhsDQ) and f reads value[refcvs_sens_index[is]] from p[is] */
int refcvs_sens_dq = 0;
int refcvs_sens_index[64];
/* 1: no Jacobian function is registered (CVDenseDQJac) */
int refcvs_jac_dq = 0;

typedef struct { int neq; double *value; c_ode_fn f; c_jac_fn j; c_sens_fn s; int ns; const double *p; double p_orig[64]; } ctx_t;

/* use_p (src/cvodeSolver.c:981-984, :1019-1022, :1060-1063, :1084-1087): the parameters come from the array CVODES
   perturbs and are put back to their original values after the evaluation, in f and in the Jacobian alike */
static void f_cb(realtype t, N_Vector y, N_Vector ydot, void *f_data)
{
  ctx_t *c = (ctx_t *)f_data; int is;
  if (c->p) for (is = 0; is < c->ns; is++) c->value[refcvs_sens_index[is]] = c->p[is];
  c->f(t, NV_DATA_S(y), NV_DATA_S(ydot), c->value);
  if (c->p) for (is = 0; is < c->ns; is++) c->value[refcvs_sens_index[is]] = c->p_orig[is];
}
static void jac_cb(long int N, DenseMat J, realtype t, N_Vector y, N_Vector fy, void *jac_data, N_Vector t1, N_Vector t2, N_Vector t3)
{
  ctx_t *c = (ctx_t *)jac_data; int is; (void)N; (void)fy; (void)t1; (void)t2; (void)t3;
  if (c->p) for (is = 0; is < c->ns; is++) c->value[refcvs_sens_index[is]] = c->p[is];
  c->j(t, NV_DATA_S(y), J->data[0], c->value);
  if (c->p) for (is = 0; is < c->ns; is++) c->value[refcvs_sens_index[is]] = c->p_orig[is];
}
static void fs_cb(int Ns, realtype t, N_Vector y, N_Vector ydot, int iS, N_Vector yS, N_Vector ySdot, void *fS_data, N_Vector t1, N_Vector t2)
{ ctx_t *c = (ctx_t *)fS_data; (void)Ns; (void)ydot; (void)t1; (void)t2; c->s(iS, t, NV_DATA_S(y), NV_DATA_S(yS), NV_DATA_S(ySdot), c->value); }

typedef struct {
  int first, last, neq, nvalues, nsens, nvary, nout; long mxstep;
  const int *vary_idx, *sens_variable; const double *values, *base, *tpoints;
  double rtol, atol;
  double *out, *sens; int *status, *stats;
  c_ode_fn f; c_jac_fn j; c_sens_fn s;
  int failed;
  odeModel_t *om; int resetOnEvent, haltOnEvent; unsigned int *fired; const int *trigger0;
  int ism;                       /* CV_SIMULTANEOUS or CV_STAGGERED (cvodeSettings_t.SensMethod 0 / 1, src/sensSolver.c:286-290) */
} job_t;

static void set_up_solver(void *mem, ctx_t *ctx, job_t *jb, double *p)
{
  CVodeSetFdata(mem, ctx);
  CVodeSetMaxNumSteps(mem, jb->mxstep);
  int is;
  CVDense(mem, jb->neq);
  if (!refcvs_jac_dq) CVDenseSetJacFn(mem, jac_cb, ctx);
  if (!refcvs_sens_dq) { CVodeSetSensRhs1Fn(mem, fs_cb); CVodeSetSensFdata(mem, ctx); }
  else for (is = 0; is < jb->nsens; is++) p[is] = ctx->p_orig[is] = ctx->value[refcvs_sens_index[is]];
  CVodeSetSensParams(mem, p, NULL, NULL);
  CVodeSetSensTolerances(mem, CV_EE, 0.0, NULL);
  CVodeSetSensErrCon(mem, FALSE);
}
static void fold_stats(void *mem, long *tot)
{
  long v = 0;
  CVodeGetNumSteps(mem, &v); tot[0] += v;
  CVodeGetNumRhsEvals(mem, &v); tot[1] += v;
  CVodeGetNumLinSolvSetups(mem, &v); tot[2] += v;
  CVDenseGetNumJacEvals(mem, &v); tot[3] += v;
  CVodeGetNumNonlinSolvIters(mem, &v); tot[4] += v;
  CVodeGetNumNonlinSolvConvFails(mem, &v); tot[5] += v;
  CVodeGetNumErrTestFails(mem, &v); tot[6] += v;
  CVodeGetNumSensRhsEvals(mem, &v); tot[7] += v;
}

static void *run_range(void *arg)
{
  job_t *jb = (job_t *)arg;
  int neq = jb->neq, ns = jb->nsens, nout = jb->nout, s, i, k, is, e;
  ctx_t ctx; void *mem = NULL; int created = 0; long tot[8];
  odeModel_t *om = jb->om;
  int *trigger = (int *)calloc((size_t)((om && om->nevents > 0) ? om->nevents : 1), sizeof(int));
  N_Vector y = N_VNew_Serial(neq), abstol = N_VNew_Serial(neq);
  N_Vector *yS = N_VNewVectorArray_Serial(ns, neq);
  double *p = (double *)calloc((size_t)ns, sizeof(double));
  ctx.neq = neq; ctx.value = (double *)malloc((size_t)jb->nvalues * sizeof(double)); ctx.f = jb->f; ctx.j = jb->j; ctx.s = jb->s; ctx.ns = ns; ctx.p = refcvs_sens_dq ? p : NULL;

  for (s = jb->first; s < jb->last; s++) {
    double t = jb->tpoints[0]; int flag = 0, iout, halted = 0;
    double *rows = jb->out + (size_t)s * (size_t)(nout + 1) * (size_t)neq;
    double *srows = jb->sens + (size_t)s * (size_t)(nout + 1) * (size_t)ns * (size_t)neq;
    memcpy(ctx.value, jb->base, (size_t)jb->nvalues * sizeof(double));
    for (k = 0; k < jb->nvary; k++) ctx.value[jb->vary_idx[k]] = jb->values[(size_t)s * (size_t)jb->nvary + (size_t)k];
    for (i = 0; i < neq; i++) { NV_Ith_S(y, i) = ctx.value[i]; NV_Ith_S(abstol, i) = jb->atol; rows[i] = ctx.value[i]; }
    for (is = 0; is < ns; is++) for (i = 0; i < neq; i++) {
      NV_Ith_S(yS[is], i) = (jb->sens_variable[is] == i) ? 1.0 : 0.0;
      srows[(size_t)is * (size_t)neq + (size_t)i] = NV_Ith_S(yS[is], i);
    }
    if (!created) {
      mem = CVodeCreate(CV_BDF, CV_NEWTON);
      CVodeSetErrFile(mem, NULL);
      flag = CVodeMalloc(mem, f_cb, t, y, CV_SV, jb->rtol, abstol);
      if (flag == CV_SUCCESS) flag = CVodeSensMalloc(mem, ns, jb->ism, yS);
      created = 1;
    } else {
      flag = CVodeReInit(mem, f_cb, t, y, CV_SV, jb->rtol, abstol);
      if (flag == CV_SUCCESS) flag = CVodeSensReInit(mem, jb->ism, yS);
    }
    if (flag == CV_SUCCESS) set_up_solver(mem, &ctx, jb, p);
    jb->status[s] = 0;
    iout = 1;
    for (i = 0; i < 8; i++) tot[i] = 0;
    if (om) for (e = 0; e < om->nevents; e++) trigger[e] = jb->trigger0[e];
    if (jb->fired) jb->fired[(size_t)s * (size_t)(nout + 1)] = 0;
    if (flag == CV_SUCCESS) {
      for (; iout <= nout; iout++) {
        realtype tret = 0; unsigned int firedmask = 0; int isValid = 1;
        flag = CVode(mem, jb->tpoints[iout], y, &tret, CV_NORMAL);
        if (flag < 0) break;
        flag = CVodeGetSens(mem, tret, yS);
        if (flag < 0) break;
        if (om && om->nevents > 0) {                 /* oracle/driver.c: rules before events, edge-triggered firing */
          double *value = ctx.value; int jj;
          for (i = 0; i < neq; i++) value[i] = NV_Ith_S(y, i);
          for (i = 0; i < om->nassbeforeevents; i++) { nonzeroElem_t *o = om->assignmentsBeforeEvents[i]; value[o->i] = oracle_eval(o->ij, value, tret); }
          for (e = 0; e < om->nevents; e++) {
            if (trigger[e] == 0) {
              if (oracle_eval(om->event[e], value, tret)) {
                firedmask |= (1u << e);
                trigger[e] = 1;
                for (jj = 0; jj < om->neventAss[e]; jj++) {
                  double r = oracle_eval(om->eventAssignment[e][jj], value, tret);
                  int idx = om->eventIndex[e][jj], a;
                  if (idx >= neq && idx < neq + om->nass) continue;
                  value[idx] = r;
                  if (idx < neq || jb->resetOnEvent) isValid = 0;
                  for (a = 0; a < om->nass; a++) { nonzeroElem_t *o = om->assignmentOrder[a]; value[o->i] = oracle_eval(o->ij, value, tret); }
                }
              }
            } else if (!oracle_eval(om->event[e], value, tret)) trigger[e] = 0;
          }
          for (i = 0; i < neq; i++) NV_Ith_S(y, i) = value[i];
        }
        for (i = 0; i < neq; i++) rows[(size_t)iout * (size_t)neq + (size_t)i] = NV_Ith_S(y, i);
        for (is = 0; is < ns; is++) for (i = 0; i < neq; i++)
          srows[((size_t)iout * (size_t)ns + (size_t)is) * (size_t)neq + (size_t)i] = NV_Ith_S(yS[is], i);
        if (jb->fired) jb->fired[(size_t)s * (size_t)(nout + 1) + (size_t)iout] = firedmask;
        if (firedmask && jb->haltOnEvent) { iout++; halted = 1; break; }
        if (!isValid && iout < nout) {               /* lazy re-creation at the next step in the reference; same arithmetic */
          fold_stats(mem, tot);
          flag = CVodeReInit(mem, f_cb, tret, y, CV_SV, jb->rtol, abstol);
          if (flag == CV_SUCCESS) flag = CVodeSensReInit(mem, jb->ism, yS);
          if (flag != CV_SUCCESS) { iout++; break; }
          set_up_solver(mem, &ctx, jb, p);
        }
      }
    }
    if (flag < 0 || halted) {
      if (flag < 0) { jb->status[s] = flag; jb->failed++; }
      for (; iout <= nout; iout++) {
        if (jb->fired) jb->fired[(size_t)s * (size_t)(nout + 1) + (size_t)iout] = 0;
        for (i = 0; i < neq; i++) rows[(size_t)iout * (size_t)neq + (size_t)i] = NAN;
        for (i = 0; i < ns * neq; i++) srows[(size_t)iout * (size_t)ns * (size_t)neq + (size_t)i] = NAN;
      }
    }
    if (jb->stats && mem) {
      int *st = jb->stats + (size_t)s * 8;
      fold_stats(mem, tot);
      for (i = 0; i < 8; i++) st[i] = (int)tot[i];
    }
  }
  if (mem) CVodeFree(mem);
  N_VDestroy_Serial(y); N_VDestroy_Serial(abstol); N_VDestroyVectorArray_Serial(yS, ns);
  free(ctx.value); free(p); free(trigger);
  return NULL;
}

int refcvs_integrate_batch_events(const char *compiled_so, int neq, int nvalues, int nsens, const int *sens_variable,
                                  const double *base, int nsets, int nvary, const int *vary_idx, const double *values,
                                  const double *tpoints, int nout, double rtol, double atol, long mxstep,
                                  double *out, double *sens, int *status, int *stats, int nthreads,
                                  odeModel_t *om, int resetOnEvent, int haltOnEvent, const int *trigger0, unsigned int *fired);
/* out [nsets][nout+1][neq], sens [nsets][nout+1][nsens][neq], status [nsets], stats [nsets][8] (nst, nfe, nsetups, nje,
   nni, ncfn, netf, nfSe) or NULL.  sens_variable[is]: equation index of an initial-condition sensitivity, or -1.
   Returns the number of failed sets, -1 on a set-up error. */
int refcvs_integrate_batch(const char *compiled_so, int neq, int nvalues, int nsens, const int *sens_variable,
                           const double *base, int nsets, int nvary, const int *vary_idx, const double *values,
                           const double *tpoints, int nout, double rtol, double atol, long mxstep,
                           double *out, double *sens, int *status, int *stats, int nthreads)
{
  return refcvs_integrate_batch_events(compiled_so, neq, nvalues, nsens, sens_variable, base, nsets, nvary, vary_idx, values, tpoints, nout, rtol, atol,
                                       mxstep, out, sens, status, stats, nthreads, NULL, 0, 0, NULL, NULL);
}
/* the same with the model's events handled at the output times (om from libodes_b200.so, initial trigger flags from the
   plan); fired: NULL or [nsets][nout+1] event bit masks */
int refcvs_integrate_batch_events(const char *compiled_so, int neq, int nvalues, int nsens, const int *sens_variable,
                                  const double *base, int nsets, int nvary, const int *vary_idx, const double *values,
                                  const double *tpoints, int nout, double rtol, double atol, long mxstep,
                                  double *out, double *sens, int *status, int *stats, int nthreads,
                                  odeModel_t *om, int resetOnEvent, int haltOnEvent, const int *trigger0, unsigned int *fired)
{
  void *so = dlopen(compiled_so, RTLD_NOW | RTLD_LOCAL);
  job_t *jobs; pthread_t *th; int k, failed = 0;
  c_ode_fn f; c_jac_fn j; c_sens_fn s;
  if (!so) return -1;
  f = (c_ode_fn)dlsym(so, "ode_f"); j = (c_jac_fn)dlsym(so, "jacobi_f"); s = (c_sens_fn)dlsym(so, "sense_f");
  if (!f || (!j && !refcvs_jac_dq) || (!s && !refcvs_sens_dq)) { dlclose(so); return -1; }
  if (nthreads < 1) nthreads = 1;
  if (nthreads > nsets) nthreads = nsets > 0 ? nsets : 1;
  jobs = (job_t *)calloc((size_t)nthreads, sizeof *jobs);
  th = (pthread_t *)calloc((size_t)nthreads, sizeof *th);
  for (k = 0; k < nthreads; k++) {
    job_t *jb = &jobs[k];
    jb->first = (int)((long long)nsets * k / nthreads); jb->last = (int)((long long)nsets * (k + 1) / nthreads);
    jb->neq = neq; jb->nvalues = nvalues; jb->nsens = nsens; jb->nvary = nvary; jb->nout = nout; jb->mxstep = mxstep;
    jb->vary_idx = vary_idx; jb->sens_variable = sens_variable; jb->values = values; jb->base = base; jb->tpoints = tpoints;
    jb->rtol = rtol; jb->atol = atol; jb->out = out; jb->sens = sens; jb->status = status; jb->stats = stats;
    jb->f = f; jb->j = j; jb->s = s;
    jb->om = om; jb->resetOnEvent = resetOnEvent; jb->haltOnEvent = haltOnEvent; jb->trigger0 = trigger0; jb->fired = fired;
    jb->ism = refcvs_sens_method == 1 ? CV_STAGGERED : refcvs_sens_method == 2 ? CV_STAGGERED1 : CV_SIMULTANEOUS;
    if (nthreads == 1) run_range(jb); else pthread_create(&th[k], NULL, run_range, jb);
  }
  for (k = 0; k < nthreads; k++) { if (nthreads > 1) pthread_join(th[k], NULL); failed += jobs[k].failed; }
  free(jobs); free(th);
  dlclose(so);
  return failed;
}
