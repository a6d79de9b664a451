/*
 * oracle.h -- ORACLE (test infrastructure, not product code).
 * CPU restatement of the reference's integration path, used only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 */
#ifndef ODES_ORACLE_H
#define ODES_ORACLE_H
#include "sbmlsolver/odeModel.h"
#include "sbmlsolver/integratorSettings.h"
#ifdef __cplusplus
extern "C" {
#endif

/* CVODE return flags, numbering of the vendored cvode.h:712-725 */
#define ORACLE_CV_SUCCESS 0
#define ORACLE_CV_TSTOP_RETURN 1
#define ORACLE_CV_ILL_INPUT (-2)
#define ORACLE_CV_TOO_MUCH_WORK (-4)
#define ORACLE_CV_TOO_MUCH_ACC (-5)
#define ORACLE_CV_ERR_FAILURE (-6)
#define ORACLE_CV_CONV_FAILURE (-7)
#define ORACLE_CV_LSETUP_FAIL (-9)
#define ORACLE_CV_LSOLVE_FAIL (-10)
#define ORACLE_CV_BAD_T (-15)

typedef void (*oracle_rhs_fn)(double t, const double *y, double *ydot, void *user);
/* J is column-major n x n, pre-zeroed */
typedef void (*oracle_jac_fn)(double t, const double *y, const double *fy, double *J, void *user);

/* solver backend interface; implemented by cvode_port.c (portcv_*) and by the vendored
   CVODE 2.3.0 adapter ref_backend.c (refcv_*, built into oracle/_ref/libcvode_ref.so) */
typedef struct {
  void *(*create)(int n, oracle_rhs_fn f, oracle_jac_fn jac, void *user);
  int (*init)(void *h, double t0, const double *y0, double reltol, double abstol, long mxstep, int reinit);
  void (*set_stop_time)(void *h, double tstop);
  int (*solve)(void *h, double tout, double *yout, double *tret, int use_tstop);
  void (*stats)(void *h, long *stats7);
  void (*destroy)(void *h);
  /* optional (the vendored CVODE only): one internal step, dense output at t, current time and last step size */
  int (*step)(void *h, double tout, double *yout, double *tret);
  int (*dky)(void *h, double t, double *yout);
  void (*step_info)(void *h, double *tn, double *hu);
} oracle_backend_t;

void *portcv_create(int n, oracle_rhs_fn f, oracle_jac_fn jac, void *user);
int portcv_init(void *h, double t0, const double *y0, double reltol, double abstol, long mxstep, int reinit);
void portcv_set_stop_time(void *h, double tstop);
int portcv_solve(void *h, double tout, double *yout, double *tret, int use_tstop);
void portcv_stats(void *h, long *stats7);
void portcv_free(void *h);

/* independent AST interpreter (restates src/evaluateAST.c) */
double oracle_eval(const ASTNode_t *n, const double *value, double time);

#define ORACLE_BACKEND_PORT 0
#define ORACLE_BACKEND_REF 1
#define ORACLE_RHS_INTERPRETED 0
#define ORACLE_RHS_COMPILED 1

/* Batch driver: per set `reset; setVariableValue x nvary; integrate` exactly as
   src/odeSolver.c:313-330 with the event/result handling of src/integratorInstance.c:1029-1237.
   out:    [nsets][nout+1][nvalues]   (NULL allowed)
   status: [nsets] CVODE flag of the failing call or 0
   stats:  [nsets][7] nst,nfe,nsetups,nje,nni,ncfn,netf (NULL allowed)
   fired:  [nsets][nout+1] event bitmasks (NULL allowed)
   margin: [nsets][nout+1] min over events of |lhs-rhs| of the trigger relation (NULL allowed)
   compiled_so: path of a shared object with ode_f/jacobi_f (ODEModel_generateCSource), or NULL.
   Returns number of failed sets, -1 on setup error. */
int oracle_integrate_batch(odeModel_t *om, const cvodeSettings_t *opt, int nsets, int nvary,
                           const int *vary_idx, const double *values, double *out, int *status,
                           int *stats, unsigned int *fired, double *margin,
                           int backend, int rhs_mode, const char *compiled_so, int nthreads);
const char *oracle_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
