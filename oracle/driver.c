/*
 * driver.c -- ORACLE (test infrastructure, not product code).
 *
 * CPU restatement of the reference's batch path around the CVODE calls:
 *   src/odeSolver.c:313-330               design-point loop: setVariableValue x nvary, integrate, reset
 *   src/cvodeData.c:156-190, :327-402     initial values: model values, then initial + assignment rules
 *                                         in initAssignmentOrder, trigger seeding
 *   src/integratorInstance.c:333-433      solver time set-up, row 0 of the results
 *   src/integratorInstance.c:1543-1594    setVariableValueByIndex (row-0 fix-up, invalidation, rule refresh)
 *   src/cvodeSolver.c:81-240              cvodeOneStep: lazy (re)creation at solver->t, optional tstop,
 *                                         CVode to tout, error -> stop, copy y
 *   src/cvodeSolver.c:953-1089            interpreted f and JacODE
 *   src/integratorInstance.c:1029-1077    edge-triggered events, assignments, re-arming
 *   src/integratorInstance.c:1088-1237    updateData: events, rule refresh, row store, next tout
 * The step arithmetic comes from a backend: the restated port (cvode_port.c)
 * or the reference's vendored CVODE 2.3.0 itself (oracle/_ref/libcvode_ref.so).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

static char last_error[512];
const char *oracle_last_error(void) { return last_error; }

typedef void (*compiled_ode_fn)(double t, const double *y, double *ydot, double *value);
typedef void (*compiled_jac_fn)(double t, const double *y, double *J, double *value);

typedef struct {
  odeModel_t *om; double *value; int nvalues;
  compiled_ode_fn cf; compiled_jac_fn cj;
} ctx_t;

/* interpreted callbacks, src/cvodeSolver.c:953-1035 and :1047-1089 */
static void rhs_interp(double t, const double *y, double *ydot, void *user)
{
  ctx_t *c = (ctx_t *)user; odeModel_t *om = c->om; int i;
  for (i = 0; i < om->neq; i++) c->value[i] = y[i];
  for (i = 0; i < om->nassbeforeodes; i++) {
    nonzeroElem_t *o = om->assignmentsBeforeODEs[i];
    c->value[o->i] = oracle_eval(o->ij, c->value, t);
  }
  for (i = 0; i < om->neq; i++) ydot[i] = oracle_eval(om->ode[i], c->value, t);
}
static void jac_interp(double t, const double *y, const double *fy, double *J, void *user)
{
  ctx_t *c = (ctx_t *)user; odeModel_t *om = c->om; int i;
  (void)fy;
  for (i = 0; i < om->neq; i++) c->value[i] = y[i];
  for (i = 0; i < om->sparsesize; i++) {
    nonzeroElem_t *nz = om->jacobSparse[i];
    J[(size_t)nz->j * (size_t)om->neq + (size_t)nz->i] = oracle_eval(nz->ij, c->value, t);
  }
}
static void rhs_compiled(double t, const double *y, double *ydot, void *user) { ctx_t *c = (ctx_t *)user; c->cf(t, y, ydot, c->value); }
static void jac_compiled(double t, const double *y, const double *fy, double *J, void *user) { ctx_t *c = (ctx_t *)user; (void)fy; c->cj(t, y, J, c->value); }

static void refresh_rules(odeModel_t *om, double *value, double t)
{
  int i;
  for (i = 0; i < om->nass; i++) { nonzeroElem_t *o = om->assignmentOrder[i]; value[o->i] = oracle_eval(o->ij, value, t); }
}

static double trigger_margin(const ASTNode_t *trig, const double *value, double t)
{
  if (trig && trig->type >= AST_RELATIONAL_EQ && trig->type <= AST_RELATIONAL_NEQ && trig->nchildren == 2)
    return fabs(oracle_eval(trig->children[0], value, t) - oracle_eval(trig->children[1], value, t));
  return INFINITY;
}

typedef struct {
  odeModel_t *om; const cvodeSettings_t *opt; const oracle_backend_t *be;
  int first, last, nvary; const int *vary_idx; const double *values;
  double *out; int *status; int *stats; unsigned int *fired; double *margin;
  const double *base; const int *trigger0;
  int use_jac; compiled_ode_fn cf; compiled_jac_fn cj;
  int failed;
} job_t;

/* ---- events located inside the steps (cvodeSettings_t.StepResolution < 0) ----
   The procedure of the kernel's root mode (bdf_warp_kernel.cuh, integrate_warp), driven over the vendored CVODE step by
   step (CV_ONE_STEP + CVodeGetDky): after every internal step the triggers are evaluated at the step's end; a trigger that
   became true since the last known trigger state is traced back by bisection on the dense output to 100 uround
   (|tn| + |hu|); output times before the located time are handled first (interpolation, the reference's output-time event
   handling, row store), then the event is applied at the located time and -- if it assigned -- the solver re-initialised
   there.  The step limit counts per output interval, like the CVode calls of the reference. */
static unsigned int trigger_mask(odeModel_t *om, double *value, double t)
{
  unsigned int m = 0; int i, e;
  for (i = 0; i < om->nassbeforeevents; i++) { nonzeroElem_t *o = om->assignmentsBeforeEvents[i]; value[o->i] = oracle_eval(o->ij, value, t); }
  for (e = 0; e < om->nevents; e++) if (oracle_eval(om->event[e], value, t)) m |= 1u << e;
  return m;
}
static unsigned int apply_events(odeModel_t *om, const cvodeSettings_t *opt, double *value, int *trigger, double t, int *isValid)
{
  unsigned int firedmask = 0; int i, j, e, neq = om->neq;
  for (i = 0; i < om->nassbeforeevents; i++) { nonzeroElem_t *o = om->assignmentsBeforeEvents[i]; value[o->i] = oracle_eval(o->ij, value, t); }
  for (e = 0; e < om->nevents; e++) {
    if (trigger[e] == 0) {
      if (oracle_eval(om->event[e], value, t)) {
        firedmask |= (1u << e);
        trigger[e] = 1;
        for (j = 0; j < om->neventAss[e]; j++) {
          double r = oracle_eval(om->eventAssignment[e][j], value, t);
          int idx = om->eventIndex[e][j];
          if (idx >= neq && idx < neq + om->nass) continue;
          value[idx] = r;
          if (idx < neq || opt->ResetCvodeOnEvent) *isValid = 0;
          refresh_rules(om, value, t);
        }
      }
    } else if (!oracle_eval(om->event[e], value, t)) trigger[e] = 0;
  }
  return firedmask;
}
/* returns the CVODE flag (< 0 on failure); *piout = first row not stored */
static int run_rootfind(odeModel_t *om, const cvodeSettings_t *opt, const oracle_backend_t *be, void *h, double *value, double *y,
                        int *trigger, double *rows, unsigned int *fired, int *pcreated, int *piout)
{
  const int nvalues = om->neq + om->nass + om->nconst, neq = om->neq, nout = opt->PrintStep;
  double t = opt->TimePoints[0], troot = 0.0, rlo = 0.0, tn = t, hu = 0.0, tdummy;
  double *vtmp = (double *)malloc((size_t)nvalues * sizeof(double)), *yend = (double *)malloc((size_t)(neq > 0 ? neq : 1) * sizeof(double));
  int isValid = 0, pending = 0, fresh = 1, need_check = 0, iout = 1, flag = 0, i, e; long nstloc = 0;
  unsigned int acc = 0;
  while (iout <= nout) {
    const double tout = opt->TimePoints[iout];
    int at_output;
    if (!isValid) {
      for (i = 0; i < neq; i++) y[i] = value[i];
      flag = be->init(h, t, y, opt->RError, opt->Error, (long)opt->Mxstep, *pcreated);
      *pcreated = 1; isValid = 1;
      if (flag < 0) break;
      pending = 0; fresh = 1; need_check = 0; nstloc = 0; tn = t; hu = 0.0;
    }
    at_output = !fresh && tn >= tout;
    if (need_check) {
      unsigned int known = 0, cand;
      need_check = 0;
      for (e = 0; e < om->nevents; e++) if (trigger[e]) known |= 1u << e;
      memcpy(vtmp, value, (size_t)nvalues * sizeof(double));
      for (i = 0; i < neq; i++) vtmp[i] = yend[i];              /* zn[0]: the state at the end of the step */
      cand = trigger_mask(om, vtmp, tn) & ~known;
      if (cand) {
        double lo = rlo, hi = tn; int it;
        const double ttol = 100.0 * 2.2204460492503131e-16 * (fabs(tn) + fabs(hu));
        for (it = 0; it < 200 && (hi - lo) > ttol; it++) {
          const double mid = lo + 0.5 * (hi - lo);
          be->dky(h, mid, y);
          for (i = 0; i < neq; i++) vtmp[i] = y[i];
          if (trigger_mask(om, vtmp, mid) & cand) hi = mid; else lo = mid;
        }
        troot = hi; pending = 1;
      }
      continue;
    }
    if (pending && troot <= tout) {
      be->dky(h, troot, y);
      for (i = 0; i < neq; i++) value[i] = y[i];
      acc |= apply_events(om, opt, value, trigger, troot, &isValid);
      pending = 0;
      if (!isValid) t = troot;
      else { rlo = troot; need_check = 1; }
      continue;
    }
    if (at_output) {
      t = tout;
      be->dky(h, t, y);
      for (i = 0; i < neq; i++) value[i] = y[i];
      acc |= apply_events(om, opt, value, trigger, t, &isValid);
      refresh_rules(om, value, t);
      if (rows) memcpy(rows + (size_t)iout * (size_t)nvalues, value, (size_t)nvalues * sizeof(double));
      if (fired) fired[iout] = acc;
      acc = 0;
      iout++;
      nstloc = 0;
      continue;
    }
    if (nstloc >= (long)opt->Mxstep) { flag = ORACLE_CV_TOO_MUCH_WORK; break; }
    flag = be->step(h, tout, y, &tdummy);
    if (flag < 0) break;
    flag = 0;
    nstloc++; fresh = 0;
    for (i = 0; i < neq; i++) yend[i] = y[i];
    be->step_info(h, &tn, &hu);
    rlo = tn - hu; need_check = 1;
  }
  free(vtmp); free(yend);
  *piout = iout;
  return flag;
}

static void *run_range(void *arg)
{
  job_t *jb = (job_t *)arg;
  odeModel_t *om = jb->om; const cvodeSettings_t *opt = jb->opt; const oracle_backend_t *be = jb->be;
  int nvalues = om->neq + om->nass + om->nconst, neq = om->neq, nout = opt->PrintStep, s, i, j, e;
  double *value = (double *)malloc((size_t)nvalues * sizeof(double));
  double *y = (double *)malloc((size_t)(neq > 0 ? neq : 1) * sizeof(double));
  int *trigger = (int *)malloc((size_t)(om->nevents > 0 ? om->nevents : 1) * sizeof(int));
  int use_tstop = opt->SetTStop || om->npiecewise;
  ctx_t ctx; void *h = NULL;
  ctx.om = om; ctx.value = value; ctx.nvalues = nvalues; ctx.cf = jb->cf; ctx.cj = jb->cj;
  if (neq > 0)
    h = be->create(neq, jb->cf ? rhs_compiled : rhs_interp, jb->use_jac ? (jb->cj ? jac_compiled : jac_interp) : NULL, &ctx);

  for (s = jb->first; s < jb->last; s++) {
    double t = opt->TimePoints[0];
    int isValid = 0, created = 0, iout, flag = 0, halted = 0;
    double *rows = jb->out ? jb->out + (size_t)s * (size_t)(nout + 1) * (size_t)nvalues : NULL;
    memcpy(value, jb->base, (size_t)nvalues * sizeof(double));
    for (e = 0; e < om->nevents; e++) trigger[e] = jb->trigger0[e];
    if (rows) memcpy(rows, value, (size_t)nvalues * sizeof(double));            /* row 0 */
    if (jb->fired) jb->fired[(size_t)s * (size_t)(nout + 1)] = 0;
    if (jb->margin) jb->margin[(size_t)s * (size_t)(nout + 1)] = INFINITY;
    /* setVariableValue for every varied entry */
    for (j = 0; j < jb->nvary; j++) {
      int idx = jb->vary_idx[j]; double v = jb->values[(size_t)s * (size_t)jb->nvary + (size_t)j];
      if (idx >= neq && idx < neq + om->nass) continue;       /* assigned values cannot be set */
      value[idx] = v;
      if (t == 0.0 && rows) rows[idx] = v;
      refresh_rules(om, value, t);
    }
    if (jb->status) jb->status[s] = 0;

    if (opt->StepResolution < 0 && om->nevents > 0 && neq > 0 && be->step && be->dky && be->step_info) {
      flag = run_rootfind(om, opt, be, h, value, y, trigger, rows, jb->fired ? jb->fired + (size_t)s * (size_t)(nout + 1) : NULL, &created, &iout);
      if (jb->margin) for (i = 1; i <= nout; i++) jb->margin[(size_t)s * (size_t)(nout + 1) + (size_t)i] = INFINITY;
    } else
    for (iout = 1; iout <= nout; iout++) {
      double tout = opt->TimePoints[iout];
      unsigned int firedmask = 0; double mg = INFINITY;
      if (neq == 0) t = tout;
      else {
        if (!isValid) {
          for (i = 0; i < neq; i++) y[i] = value[i];
          flag = be->init(h, t, y, opt->RError, opt->Error, (long)opt->Mxstep, created);
          created = 1; isValid = 1;
          if (flag < 0) break;
        }
        if (use_tstop) be->set_stop_time(h, tout);
        flag = be->solve(h, tout, y, &t, use_tstop);
        if (flag < 0) break;
        for (i = 0; i < neq; i++) value[i] = y[i];
      }
      /* events: rules needed by triggers, then edge-triggered firing in model order */
      for (i = 0; i < om->nassbeforeevents; i++) { nonzeroElem_t *o = om->assignmentsBeforeEvents[i]; value[o->i] = oracle_eval(o->ij, value, t); }
      for (e = 0; e < om->nevents; e++) {
        double m = trigger_margin(om->event[e], value, t);
        if (m < mg) mg = m;
        if (trigger[e] == 0) {
          if (oracle_eval(om->event[e], value, t)) {
            firedmask |= (1u << e);
            trigger[e] = 1;
            for (j = 0; j < om->neventAss[e]; j++) {
              double r = oracle_eval(om->eventAssignment[e][j], value, t);
              int idx = om->eventIndex[e][j];
              if (idx >= neq && idx < neq + om->nass) continue;
              value[idx] = r;
              if (idx < neq || opt->ResetCvodeOnEvent) isValid = 0;
              refresh_rules(om, value, t);
            }
          }
        } else if (!oracle_eval(om->event[e], value, t)) trigger[e] = 0;
      }
      if (firedmask && opt->HaltOnEvent) halted = 1;
      /* store: all rules refreshed, all values written */
      refresh_rules(om, value, t);
      if (rows) memcpy(rows + (size_t)iout * (size_t)nvalues, value, (size_t)nvalues * sizeof(double));
      if (jb->fired) jb->fired[(size_t)s * (size_t)(nout + 1) + (size_t)iout] = firedmask;
      if (jb->margin) jb->margin[(size_t)s * (size_t)(nout + 1) + (size_t)iout] = mg;
      /* steady-state detection after the row is stored (src/integratorInstance.c:1176-1182, :1442-1507): mean and
         standard deviation of the right-hand sides on the refreshed values */
      if (opt->SteadyState == 1 && neq > 0) {
        double mean = 0.0, var = 0.0;
        for (i = 0; i < neq; i++) mean += fabs(oracle_eval(om->ode[i], value, t));
        mean = mean / neq;
        for (i = 0; i < neq; i++) { double dd = oracle_eval(om->ode[i], value, t) - mean; var += dd * dd; }
        var = var / (neq - 1);
        if ((mean + sqrt(var)) < opt->ssThreshold) halted = 1;
      }
      if (halted) { iout++; break; }
    }
    if (flag < 0 || halted) {
      if (flag < 0) { if (jb->status) jb->status[s] = flag; jb->failed++; }
      for (; iout <= nout; iout++) {           /* rows never reached */
        if (rows) for (i = 0; i < nvalues; i++) rows[(size_t)iout * (size_t)nvalues + (size_t)i] = NAN;
        if (jb->fired) jb->fired[(size_t)s * (size_t)(nout + 1) + (size_t)iout] = 0;
        if (jb->margin) jb->margin[(size_t)s * (size_t)(nout + 1) + (size_t)iout] = INFINITY;
      }
    }
    if (jb->stats) {
      long st[7] = {0, 0, 0, 0, 0, 0, 0};
      if (h && created) be->stats(h, st);
      for (i = 0; i < 7; i++) jb->stats[(size_t)s * 7 + (size_t)i] = (int)st[i];
    }
  }
  if (h) be->destroy(h);
  free(value); free(y); free(trigger);
  return NULL;
}

static int load_ref_backend(oracle_backend_t *be)
{
  static void *lib = NULL;
  if (!lib) {
    Dl_info info; char path[4096]; char *slash;
    if (!dladdr((void *)&load_ref_backend, &info) || !info.dli_fname) return 0;
    snprintf(path, sizeof path - 32, "%s", info.dli_fname);
    slash = strrchr(path, '/');
    if (slash) *slash = 0; else strcpy(path, ".");
    strcat(path, "/_ref/libcvode_ref.so");
    lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!lib) { snprintf(last_error, sizeof last_error, "cannot load %s: %s", path, dlerror()); return 0; }
  }
  be->create = (void *(*)(int, oracle_rhs_fn, oracle_jac_fn, void *))dlsym(lib, "refcv_create");
  be->init = (int (*)(void *, double, const double *, double, double, long, int))dlsym(lib, "refcv_init");
  be->set_stop_time = (void (*)(void *, double))dlsym(lib, "refcv_set_stop_time");
  be->solve = (int (*)(void *, double, double *, double *, int))dlsym(lib, "refcv_solve");
  be->stats = (void (*)(void *, long *))dlsym(lib, "refcv_stats");
  be->destroy = (void (*)(void *))dlsym(lib, "refcv_free");
  be->step = (int (*)(void *, double, double *, double *))dlsym(lib, "refcv_step");
  be->dky = (int (*)(void *, double, double *))dlsym(lib, "refcv_dky");
  be->step_info = (void (*)(void *, double *, double *))dlsym(lib, "refcv_step_info");
  return be->create && be->init && be->solve && be->stats && be->destroy && be->set_stop_time;
}

int oracle_integrate_batch(odeModel_t *om, const cvodeSettings_t *opt, int nsets, int nvary,
                           const int *vary_idx, const double *values, double *out, int *status,
                           int *stats, unsigned int *fired, double *margin,
                           int backend, int rhs_mode, const char *compiled_so, int nthreads)
{
  oracle_backend_t be; job_t *jobs; pthread_t *th;
  int nvalues, i, k, failed = 0, *trigger0; double *base; void *so = NULL;
  compiled_ode_fn cf = NULL; compiled_jac_fn cj = NULL;

  if (!om || !opt || om->hasCycle || opt->Indefinitely || !opt->TimePoints) { snprintf(last_error, sizeof last_error, "bad model or settings"); return -1; }
  if (backend == ORACLE_BACKEND_REF) { if (!load_ref_backend(&be)) return -1; }
  else { be.create = portcv_create; be.init = portcv_init; be.set_stop_time = portcv_set_stop_time; be.solve = portcv_solve; be.stats = portcv_stats; be.destroy = portcv_free; be.step = NULL; be.dky = NULL; be.step_info = NULL; }
  /* Jacobian: constructed on first use like IntegratorInstance_initializeJacobian (integratorInstance.c:291-311) */
  if (opt->UseJacobian && !om->jacob && !om->jacobianFailed) ODEModel_constructJacobian(om);
  if (rhs_mode == ORACLE_RHS_COMPILED) {
    if (!compiled_so || !(so = dlopen(compiled_so, RTLD_NOW | RTLD_LOCAL))) { snprintf(last_error, sizeof last_error, "cannot load compiled model %s", compiled_so ? compiled_so : "(null)"); return -1; }
    cf = (compiled_ode_fn)dlsym(so, "ode_f"); cj = (compiled_jac_fn)dlsym(so, "jacobi_f");
    if (!cf || !cj) { snprintf(last_error, sizeof last_error, "ode_f/jacobi_f missing in %s", compiled_so); return -1; }
  }
  nvalues = om->neq + om->nass + om->nconst;
  /* initial values: model values, then the complete ordered rule set (cvodeData.c:156-190) */
  base = (double *)malloc((size_t)nvalues * sizeof(double));
  for (i = 0; i < nvalues; i++) base[i] = om->values[i];
  for (i = 0; i < om->nass + om->ninitAss; i++) {
    nonzeroElem_t *o = om->initAssignmentOrder[i];
    int idx = o->i == -1 ? o->j : o->i;
    base[idx] = oracle_eval(o->ij, base, 0.0);
  }
  if (opt->TimePoints[0] != 0.0) refresh_rules(om, base, opt->TimePoints[0]);
  trigger0 = (int *)malloc((size_t)(om->nevents > 0 ? om->nevents : 1) * sizeof(int));
  for (i = 0; i < om->nevents; i++) trigger0[i] = (int)oracle_eval(om->event[i], base, opt->TimePoints[0]);

  if (nthreads < 1) nthreads = 1;
  if (nthreads > nsets) nthreads = nsets > 0 ? nsets : 1;
  jobs = (job_t *)calloc((size_t)nthreads, sizeof *jobs);
  th = (pthread_t *)calloc((size_t)nthreads, sizeof *th);
  for (k = 0; k < nthreads; k++) {
    job_t *jb = &jobs[k];
    jb->om = om; jb->opt = opt; jb->be = &be;
    jb->first = (int)((long long)nsets * k / nthreads); jb->last = (int)((long long)nsets * (k + 1) / nthreads);
    jb->nvary = nvary; jb->vary_idx = vary_idx; jb->values = values;
    jb->out = out; jb->status = status; jb->stats = stats; jb->fired = fired; jb->margin = margin;
    jb->base = base; jb->trigger0 = trigger0;
    jb->use_jac = opt->UseJacobian && om->jacobian; jb->cf = cf; jb->cj = cj;
    if (nthreads == 1) run_range(jb); else pthread_create(&th[k], NULL, run_range, jb);
  }
  for (k = 0; k < nthreads; k++) { if (nthreads > 1) pthread_join(th[k], NULL); failed += jobs[k].failed; }
  free(jobs); free(th); free(base); free(trigger0);
  if (so) dlclose(so);
  return failed;
}
