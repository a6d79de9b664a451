/*
 * cvode_port.c -- ORACLE (test infrastructure, not product code).
 *
 * Plain-C restatement of the time stepping the reference delegates to
 * SUNDIALS CVODE(S): variable-order variable-step BDF in Nordsieck form with
 * modified Newton iteration on a dense LU.  It follows the CVODE 2.3.0 sources
 * vendored in the reference (SUNDIALS 2.1.1) function by function:
 *
 *   Win32/WinCVODE/sundials/cvode/source/cvode.c
 *     :875-1210   CVode            -> portcv_solve
 *     :1228-1279  CVodeGetDky      -> get_dky
 *     :1514-1636  CVHin, CVUpperBoundH0, CVYddNorm
 *     :1657-1710  CVStep           -> cv_step
 *     :1723-1902  CVAdjustParams/Order, CVIncreaseBDF, CVDecreaseBDF, CVRescale
 *     :1912-1920  CVPredict   :1930-1942 CVSet   :2086-2146 CVSetBDF, CVSetTqBDF
 *     :2234-2371  CVnlsNewton, CVNewtonIteration
 *     :2400-2448  CVHandleNFlag, CVRestore
 *     :2470-2526  CVDoErrorTest
 *     :2540-2705  CVCompleteStep, CVPrepareNextStep, CVSetEta, CVComputeEtaqm1/qp1, CVChooseEta
 *     :3496-3549  CVEwtSet (SV: ewt_i = 1/(rtol*|y_i| + atol_i))
 *   Win32/WinCVODE/sundials/cvode/source/cvdense.c
 *     :368-412    CVDenseSetup     :423-441 CVDenseSolve    :479-537 CVDenseDQJac
 *   Win32/WinCVODE/sundials/shared/source/smalldense.c :58 gefa, :138 gesl
 *   Win32/WinCVODE/sundials/nvec_ser/nvector_serial.c :591 N_VWrmsNorm
 *
 * Floating-point operations are kept in the order of those sources (no FMA
 * contraction when compiled for baseline x86-64), so that this port reproduces
 * the vendored library bit for bit; tests/test_oracle.py checks that against
 * oracle/_ref/libcvode_ref.so (the vendored sources compiled as they are).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define L_MAX 6
#define NUM_TESTS 5
#define QMAX 5

/* control constants, cvode.c:40-143 and cvdense.h */
#define FUZZ_FACTOR 100.0
#define HLB_FACTOR 100.0
#define HUB_FACTOR 0.1
#define H_BIAS 0.5
#define MAX_ITERS 4
#define CORTES 0.1
#define THRESH 1.5
#define ETAMX1 10000.0
#define ETAMX2 10.0
#define ETAMX3 10.0
#define ETAMXF 0.2
#define ETAMIN 0.1
#define ETACF 0.25
#define ADDON 0.000001
#define BIAS1 6.0
#define BIAS2 6.0
#define BIAS3 10.0
#define ONEPSM 1.000001
#define SMALL_NST 10
#define MXNCF 10
#define MXNEF 7
#define MXNEF1 3
#define SMALL_NEF 2
#define LONG_WAIT 10
#define NLS_MAXCOR 3
#define CRDOWN 0.3
#define DGMAX 0.3
#define RDIV 2.0
#define MSBP 20
#define CVD_MSBJ 50
#define CVD_DGMAX 0.2
#define MIN_INC_MULT 1000.0
#define MXHNIL_DEFAULT 10
#define UROUND 2.2204460492503131e-16      /* DBL_EPSILON */

/* CVStep / CVnls internal codes */
#define SUCCESS_STEP 0
#define REP_ERR_FAIL (-1)
#define REP_CONV_FAIL (-2)
#define SETUP_FAILED (-3)
#define SOLVE_FAILED (-4)
#define PREDICT_AGAIN (-5)
#define DO_ERROR_TEST 1
#define SOLVED 0
#define CONV_FAIL (-1)
#define SETUP_FAIL_UNREC (-2)
#define SOLVE_FAIL_UNREC (-3)
#define FIRST_CALL 0
#define PREV_CONV_FAIL (-1)
#define PREV_ERR_FAIL (-2)
#define TRY_AGAIN 99
#define CV_NO_FAILURES 0
#define CV_FAIL_BAD_J 1
#define CV_FAIL_OTHER 2

typedef struct {
  int n;
  oracle_rhs_fn f; oracle_jac_fn jac; void *user;
  double reltol, abstol;
  long mxstep;
  double *zn[L_MAX], *ewt, *y, *acor, *tempv, *ftemp;
  double *M, *savedJ; long *pivots;            /* column-major n x n */
  int q, qprime, next_q, qwait, L;
  double h, hprime, next_h, eta, hscale, tn;
  double tau[L_MAX + 1], tq[NUM_TESTS + 1], l[L_MAX];
  double rl1, gamma, gammap, gamrat, crate, acnrm;
  int mnewt;
  double etamax, etaqm1, etaq, etaqp1;
  long nst, nfe, ncfn, netf, nni, nsetups, nhnil, nstlp, nje, nstlj, nfeD;
  long tot[7];                                  /* totals over re-inits */
  int qu; double hu, h0u, saved_tq5, tolsf;
  int jcur;
  double hmin, hmax_inv;
  int tstopset; double tstop;
} cvmem_t;

#define COL(A, j) ((A) + (size_t)(j) * (size_t)cv->n)

static double rpower_i(double base, int exponent)
{
  int i, expt = abs(exponent); double prod = 1.0;
  for (i = 1; i <= expt; i++) prod *= base;
  if (exponent < 0) prod = 1.0 / prod;
  return prod;
}
static double rpower_r(double base, double exponent) { return base <= 0.0 ? 0.0 : pow(base, exponent); }
static double rsqrt_(double x) { return x <= 0.0 ? 0.0 : sqrt(x); }

static double wrms_norm(const cvmem_t *cv, const double *x, const double *w)
{
  int i; double sum = 0.0, prodi;
  for (i = 0; i < cv->n; i++) { prodi = x[i] * w[i]; sum += prodi * prodi; }
  return rsqrt_(sum / cv->n);
}
/* z = a*x + b*y */
static void linear_sum(const cvmem_t *cv, double a, const double *x, double b, const double *y, double *z)
{
  int i, n = cv->n;
  /* the serial nvector special-cases a,b = +-1 but each case computes the same IEEE result
     as the general expression except that 1*x is x exactly, which also holds here */
  if (a == 1.0 && b == 1.0) for (i = 0; i < n; i++) z[i] = x[i] + y[i];
  else if (a == 1.0 && b == -1.0) for (i = 0; i < n; i++) z[i] = x[i] - y[i];
  else if (a == -1.0 && b == 1.0) for (i = 0; i < n; i++) z[i] = y[i] - x[i];
  else if (a == 1.0) for (i = 0; i < n; i++) z[i] = b * y[i] + x[i];
  else if (b == 1.0) for (i = 0; i < n; i++) z[i] = a * x[i] + y[i];
  else if (a == -1.0) for (i = 0; i < n; i++) z[i] = b * y[i] - x[i];
  else if (b == -1.0) for (i = 0; i < n; i++) z[i] = a * x[i] - y[i];
  else if (a == b) for (i = 0; i < n; i++) z[i] = a * (x[i] + y[i]);
  else if (a == -b) for (i = 0; i < n; i++) z[i] = a * (x[i] - y[i]);
  else for (i = 0; i < n; i++) z[i] = a * x[i] + b * y[i];
}
static void vscale(const cvmem_t *cv, double c, const double *x, double *z)
{
  int i, n = cv->n;
  if (c == 1.0) { if (z != x) for (i = 0; i < n; i++) z[i] = x[i]; }
  else if (c == -1.0) for (i = 0; i < n; i++) z[i] = -x[i];
  else for (i = 0; i < n; i++) z[i] = c * x[i];
}

static int ewt_set(const cvmem_t *cv, const double *ycur, double *weight)
{
  int i; double mn = 0;
  for (i = 0; i < cv->n; i++) {
    cv->tempv[i] = cv->reltol * fabs(ycur[i]) + cv->abstol;
    if (i == 0 || cv->tempv[i] < mn) mn = cv->tempv[i];
  }
  if (mn <= 0.0) return -1;
  for (i = 0; i < cv->n; i++) weight[i] = 1.0 / cv->tempv[i];
  return 0;
}

/* -------------------------------------------------------- dense LU (gefa/gesl) */
static long gefa(const cvmem_t *cv, double *a, long *p)
{
  long n = cv->n, i, j, k, l;
  for (k = 0; k < n - 1; k++, p++) {
    double *col_k = COL(a, k), *diag_k = col_k + k, temp, mult, a_kj;
    int swap;
    l = k;
    for (i = k + 1; i < n; i++) if (fabs(col_k[i]) > fabs(col_k[l])) l = i;
    *p = l;
    if (col_k[l] == 0.0) return k + 1;
    if ((swap = (l != k))) { temp = col_k[l]; col_k[l] = *diag_k; *diag_k = temp; }
    mult = -1.0 / (*diag_k);
    for (i = k + 1; i < n; i++) col_k[i] *= mult;
    for (j = k + 1; j < n; j++) {
      double *col_j = COL(a, j);
      a_kj = col_j[l];
      if (swap) { col_j[l] = col_j[k]; col_j[k] = a_kj; }
      if (a_kj != 0.0) for (i = k + 1; i < n; i++) col_j[i] += a_kj * col_k[i];
    }
  }
  *p = n - 1;
  if (COL(a, n - 1)[n - 1] == 0.0) return n;
  return 0;
}
static void gesl(const cvmem_t *cv, const double *a, const long *p, double *b)
{
  long n = cv->n, k, l, i; double mult;
  for (k = 0; k < n - 1; k++) {
    const double *col_k = COL(a, k);
    l = p[k]; mult = b[l];
    if (l != k) { b[l] = b[k]; b[k] = mult; }
    for (i = k + 1; i < n; i++) b[i] += mult * col_k[i];
  }
  for (k = n - 1; k >= 0; k--) {
    const double *col_k = COL(a, k);
    b[k] /= col_k[k];
    mult = -b[k];
    for (i = 0; i < k; i++) b[i] += mult * col_k[i];
  }
}

/* difference-quotient Jacobian, cvdense.c:479-537 (used when no analytic Jacobian is given) */
static void dq_jac(cvmem_t *cv, double *J, const double *fy, double *yv, double *ftmp)
{
  int j, i, n = cv->n;
  double srur = rsqrt_(UROUND), fnorm = wrms_norm(cv, fy, cv->ewt);
  double minInc = (fnorm != 0.0) ? (MIN_INC_MULT * fabs(cv->h) * UROUND * n * fnorm) : 1.0;
  for (j = 0; j < n; j++) {
    double yjsaved = yv[j], inc = fmax(srur * fabs(yjsaved), minInc / cv->ewt[j]), inc_inv;
    yv[j] += inc;
    cv->f(cv->tn, yv, ftmp, cv->user);
    yv[j] = yjsaved;
    inc_inv = 1.0 / inc;
    for (i = 0; i < n; i++) COL(J, j)[i] = inc_inv * (ftmp[i] - fy[i]);   /* N_VLinearSum a == -b case */
  }
  cv->nfeD += n;
}

static int dense_setup(cvmem_t *cv, int convfail, double *ypred, const double *fpred, int *jcurPtr,
                       double *vtemp1, double *vtemp2)
{
  int n = cv->n, i; size_t nn = (size_t)n * (size_t)n, k;
  double dgamma = fabs((cv->gamma / cv->gammap) - 1.0);
  int jbad = (cv->nst == 0) || (cv->nst > cv->nstlj + CVD_MSBJ) ||
             ((convfail == CV_FAIL_BAD_J) && (dgamma < CVD_DGMAX)) || (convfail == CV_FAIL_OTHER);
  long ier;
  if (!jbad) { *jcurPtr = 0; memcpy(cv->M, cv->savedJ, nn * sizeof(double)); }
  else {
    cv->nje++; cv->nstlj = cv->nst; *jcurPtr = 1;
    for (k = 0; k < nn; k++) cv->M[k] = 0.0;
    if (cv->jac) cv->jac(cv->tn, ypred, fpred, cv->M, cv->user);
    else dq_jac(cv, cv->M, fpred, ypred, vtemp1);
    (void)vtemp2;
    memcpy(cv->savedJ, cv->M, nn * sizeof(double));
  }
  for (k = 0; k < nn; k++) cv->M[k] *= -cv->gamma;            /* M = I - gamma*J */
  for (i = 0; i < n; i++) COL(cv->M, i)[i] += 1.0;
  ier = gefa(cv, cv->M, cv->pivots);
  return ier > 0 ? 1 : 0;
}
static int dense_solve(cvmem_t *cv, double *b)
{
  gesl(cv, cv->M, cv->pivots, b);
  if (cv->gamrat != 1.0) vscale(cv, 2.0 / (1.0 + cv->gamrat), b, b);
  return 0;
}

/* -------------------------------------------------------------- initial step */
static double upper_bound_h0(cvmem_t *cv, double tdist)
{
  int i; double hub_inv = 0.0, hub, *temp1 = cv->tempv, *temp2 = cv->acor;
  for (i = 0; i < cv->n; i++) temp2[i] = fabs(cv->zn[0][i]);
  /* efun(zn[0], temp1) uses tempv as scratch AND as output: weights first, then inverted back */
  for (i = 0; i < cv->n; i++) temp1[i] = cv->reltol * fabs(cv->zn[0][i]) + cv->abstol;
  for (i = 0; i < cv->n; i++) temp1[i] = 1.0 / temp1[i];      /* ewt values            */
  for (i = 0; i < cv->n; i++) temp1[i] = 1.0 / temp1[i];      /* N_VInv(temp1, temp1)  */
  for (i = 0; i < cv->n; i++) temp1[i] = HUB_FACTOR * temp2[i] + temp1[i];
  for (i = 0; i < cv->n; i++) temp2[i] = fabs(cv->zn[1][i]);
  for (i = 0; i < cv->n; i++) temp1[i] = temp2[i] / temp1[i];
  for (i = 0; i < cv->n; i++) if (fabs(temp1[i]) > hub_inv) hub_inv = fabs(temp1[i]);
  hub = HUB_FACTOR * tdist;
  if (hub * hub_inv > 1.0) hub = 1.0 / hub_inv;
  return hub;
}
static double ydd_norm(cvmem_t *cv, double hg)
{
  linear_sum(cv, hg, cv->zn[1], 1.0, cv->zn[0], cv->y);
  cv->f(cv->tn + hg, cv->y, cv->tempv, cv->user); cv->nfe++;
  linear_sum(cv, 1.0, cv->tempv, -1.0, cv->zn[1], cv->tempv);
  vscale(cv, 1.0 / hg, cv->tempv, cv->tempv);
  return wrms_norm(cv, cv->tempv, cv->ewt);
}
static int cv_hin(cvmem_t *cv, double tout)
{
  int sign, count; double tdiff, tdist, tround, hlb, hub, hg, hgs, hnew = 0, hrat, h0, yddnrm;
  if ((tdiff = tout - cv->tn) == 0.0) return 0;
  sign = (tdiff > 0.0) ? 1 : -1;
  tdist = fabs(tdiff);
  tround = UROUND * fmax(fabs(cv->tn), fabs(tout));
  if (tdist < 2.0 * tround) return 0;
  hlb = HLB_FACTOR * tround;
  hub = upper_bound_h0(cv, tdist);
  hg = rsqrt_(hlb * hub);
  if (hub < hlb) { if (sign == -1) hg = -hg; cv->h = hg; return 1; }
  count = 0;
  for (;;) {
    hgs = hg * sign;
    yddnrm = ydd_norm(cv, hgs);
    hnew = (yddnrm * hub * hub > 2.0) ? rsqrt_(2.0 / yddnrm) : rsqrt_(hg * hub);
    count++;
    if (count >= MAX_ITERS) break;
    hrat = hnew / hg;
    if ((hrat > 0.5) && (hrat < 2.0)) break;
    if ((count >= 2) && (hrat > 2.0)) { hnew = hg; break; }
    hg = hnew;
  }
  h0 = H_BIAS * hnew;
  if (h0 < hlb) h0 = hlb;
  if (h0 > hub) h0 = hub;
  if (sign == -1) h0 = -h0;
  cv->h = h0;
  return 1;
}

/* ------------------------------------------------------------ step machinery */
static void cv_rescale(cvmem_t *cv)
{
  int j; double factor = cv->eta;
  for (j = 1; j <= cv->q; j++) { vscale(cv, factor, cv->zn[j], cv->zn[j]); factor *= cv->eta; }
  cv->h = cv->hscale * cv->eta;
  cv->hscale = cv->h;
}
static void cv_increase_bdf(cvmem_t *cv)
{
  double alpha0, alpha1, prod, xi, xiold, hsum, A1; int i, j;
  for (i = 0; i <= QMAX; i++) cv->l[i] = 0.0;
  cv->l[2] = alpha1 = prod = xiold = 1.0;
  alpha0 = -1.0;
  hsum = cv->hscale;
  if (cv->q > 1)
    for (j = 1; j < cv->q; j++) {
      hsum += cv->tau[j + 1];
      xi = hsum / cv->hscale;
      prod *= xi;
      alpha0 -= 1.0 / (j + 1);
      alpha1 += 1.0 / xi;
      for (i = j + 2; i >= 2; i--) cv->l[i] = cv->l[i] * xiold + cv->l[i - 1];
      xiold = xi;
    }
  A1 = (-alpha0 - alpha1) / prod;
  vscale(cv, A1, cv->zn[QMAX], cv->zn[cv->L]);
  for (j = 2; j <= cv->q; j++) linear_sum(cv, cv->l[j], cv->zn[cv->L], 1.0, cv->zn[j], cv->zn[j]);
}
static void cv_decrease_bdf(cvmem_t *cv)
{
  double hsum, xi; int i, j;
  for (i = 0; i <= QMAX; i++) cv->l[i] = 0.0;
  cv->l[2] = 1.0;
  hsum = 0.0;
  for (j = 1; j <= cv->q - 2; j++) {
    hsum += cv->tau[j];
    xi = hsum / cv->hscale;
    for (i = j + 2; i >= 2; i--) cv->l[i] = cv->l[i] * xi + cv->l[i - 1];
  }
  for (j = 2; j < cv->q; j++) linear_sum(cv, -cv->l[j], cv->zn[cv->q], 1.0, cv->zn[j], cv->zn[j]);
}
static void cv_adjust_order(cvmem_t *cv, int deltaq)
{
  if ((cv->q == 2) && (deltaq != 1)) return;
  if (deltaq == 1) cv_increase_bdf(cv); else if (deltaq == -1) cv_decrease_bdf(cv);
}
static void cv_adjust_params(cvmem_t *cv)
{
  if (cv->qprime != cv->q) { cv_adjust_order(cv, cv->qprime - cv->q); cv->q = cv->qprime; cv->L = cv->q + 1; cv->qwait = cv->L; }
  cv_rescale(cv);
}
static void cv_predict(cvmem_t *cv)
{
  int j, k;
  cv->tn += cv->h;
  for (k = 1; k <= cv->q; k++)
    for (j = cv->q; j >= k; j--) linear_sum(cv, 1.0, cv->zn[j - 1], 1.0, cv->zn[j], cv->zn[j - 1]);
}
static void cv_restore(cvmem_t *cv, double saved_t)
{
  int j, k;
  cv->tn = saved_t;
  for (k = 1; k <= cv->q; k++)
    for (j = cv->q; j >= k; j--) linear_sum(cv, 1.0, cv->zn[j - 1], -1.0, cv->zn[j], cv->zn[j - 1]);
}
static void cv_set_bdf(cvmem_t *cv)
{
  double alpha0, alpha0_hat, xi_inv, xistar_inv, hsum, A1, A2, A3, A4, A5, A6, C, CPrime, CPrimePrime;
  int i, j, q = cv->q;
  double *l = cv->l, *tq = cv->tq, h = cv->h;
  l[0] = l[1] = xi_inv = xistar_inv = 1.0;
  for (i = 2; i <= q; i++) l[i] = 0.0;
  alpha0 = alpha0_hat = -1.0;
  hsum = h;
  if (q > 1) {
    for (j = 2; j < q; j++) {
      hsum += cv->tau[j - 1];
      xi_inv = h / hsum;
      alpha0 -= 1.0 / j;
      for (i = j; i >= 1; i--) l[i] += l[i - 1] * xi_inv;
    }
    alpha0 -= 1.0 / q;
    xistar_inv = -l[1] - alpha0;
    hsum += cv->tau[q - 1];
    xi_inv = h / hsum;
    alpha0_hat = -l[1] - xi_inv;
    for (i = q; i >= 1; i--) l[i] += l[i - 1] * xistar_inv;
  }
  /* CVSetTqBDF */
  A1 = 1.0 - alpha0_hat + alpha0;
  A2 = 1.0 + q * A1;
  tq[2] = fabs(alpha0 * (A2 / A1));
  tq[5] = fabs((A2) / (l[q] * xi_inv / xistar_inv));
  if (cv->qwait == 1) {
    C = xistar_inv / l[q];
    A3 = alpha0 + 1.0 / q;
    A4 = alpha0_hat + xi_inv;
    CPrime = A3 / (1.0 - A4 + A3);
    tq[1] = fabs(CPrime / C);
    hsum += cv->tau[q];
    xi_inv = h / hsum;
    A5 = alpha0 - (1.0 / (q + 1));
    A6 = alpha0_hat - xi_inv;
    CPrimePrime = A2 / (1.0 - A6 + A5);
    tq[3] = fabs(CPrimePrime * xi_inv * (q + 2) * A5);
  }
  tq[4] = CORTES * tq[2];
}
static void cv_set(cvmem_t *cv)
{
  cv_set_bdf(cv);
  cv->rl1 = 1.0 / cv->l[1];
  cv->gamma = cv->h * cv->rl1;
  if (cv->nst == 0) cv->gammap = cv->gamma;
  cv->gamrat = (cv->nst > 0) ? cv->gamma / cv->gammap : 1.0;
}

static int newton_iteration(cvmem_t *cv)
{
  int m = 0, ret; double del, delp = 0.0, dcon, *b;
  cv->mnewt = 0;
  for (;;) {
    linear_sum(cv, cv->rl1, cv->zn[1], 1.0, cv->acor, cv->tempv);
    linear_sum(cv, cv->gamma, cv->ftemp, -1.0, cv->tempv, cv->tempv);
    b = cv->tempv;
    ret = dense_solve(cv, b);
    cv->nni++;
    if (ret < 0) return SOLVE_FAIL_UNREC;
    if (ret > 0) { if (!cv->jcur) return TRY_AGAIN; return CONV_FAIL; }
    del = wrms_norm(cv, b, cv->ewt);
    linear_sum(cv, 1.0, cv->acor, 1.0, b, cv->acor);
    linear_sum(cv, 1.0, cv->zn[0], 1.0, cv->acor, cv->y);
    if (m > 0) cv->crate = fmax(CRDOWN * cv->crate, del / delp);
    dcon = del * fmin(1.0, cv->crate) / cv->tq[4];
    if (dcon <= 1.0) {
      cv->acnrm = (m == 0) ? del : wrms_norm(cv, cv->acor, cv->ewt);
      cv->jcur = 0;
      return SOLVED;
    }
    cv->mnewt = ++m;
    if ((m == NLS_MAXCOR) || ((m >= 2) && (del > RDIV * delp))) { if (!cv->jcur) return TRY_AGAIN; return CONV_FAIL; }
    delp = del;
    cv->f(cv->tn, cv->y, cv->ftemp, cv->user); cv->nfe++;
  }
}
static int nls_newton(cvmem_t *cv, int nflag)
{
  int convfail, ier, callSetup, i;
  convfail = ((nflag == FIRST_CALL) || (nflag == PREV_ERR_FAIL)) ? CV_NO_FAILURES : CV_FAIL_OTHER;
  callSetup = (nflag == PREV_CONV_FAIL) || (nflag == PREV_ERR_FAIL) || (cv->nst == 0) ||
              (cv->nst >= cv->nstlp + MSBP) || (fabs(cv->gamrat - 1.0) > DGMAX);
  for (;;) {
    cv->f(cv->tn, cv->zn[0], cv->ftemp, cv->user); cv->nfe++;
    if (callSetup) {
      ier = dense_setup(cv, convfail, cv->zn[0], cv->ftemp, &cv->jcur, cv->acor, cv->y);
      cv->nsetups++;
      callSetup = 0;
      cv->gamrat = cv->crate = 1.0;
      cv->gammap = cv->gamma;
      cv->nstlp = cv->nst;
      if (ier < 0) return SETUP_FAIL_UNREC;
      if (ier > 0) return CONV_FAIL;
    }
    for (i = 0; i < cv->n; i++) cv->acor[i] = 0.0;
    vscale(cv, 1.0, cv->zn[0], cv->y);
    ier = newton_iteration(cv);
    if (ier != TRY_AGAIN) return ier;
    callSetup = 1;
    convfail = CV_FAIL_BAD_J;
  }
}
static int handle_nflag(cvmem_t *cv, int *nflagPtr, double saved_t, int *ncfPtr)
{
  int nflag = *nflagPtr;
  if (nflag == SOLVED) return DO_ERROR_TEST;
  cv->ncfn++;
  cv_restore(cv, saved_t);
  if (nflag == SETUP_FAIL_UNREC) return SETUP_FAILED;
  if (nflag == SOLVE_FAIL_UNREC) return SOLVE_FAILED;
  (*ncfPtr)++;
  cv->etamax = 1.0;
  if ((fabs(cv->h) <= cv->hmin * ONEPSM) || (*ncfPtr == MXNCF)) return REP_CONV_FAIL;
  cv->eta = fmax(ETACF, cv->hmin / fabs(cv->h));
  *nflagPtr = PREV_CONV_FAIL;
  cv_rescale(cv);
  return PREDICT_AGAIN;
}
static int do_error_test(cvmem_t *cv, int *nflagPtr, int *kflagPtr, double saved_t, int *nefPtr, double *dsmPtr)
{
  double dsm = cv->acnrm / cv->tq[2];
  *dsmPtr = dsm;
  if (dsm <= 1.0) return 1;
  (*nefPtr)++; cv->netf++;
  *nflagPtr = PREV_ERR_FAIL;
  cv_restore(cv, saved_t);
  if ((fabs(cv->h) <= cv->hmin * ONEPSM) || (*nefPtr == MXNEF)) { *kflagPtr = REP_ERR_FAIL; return 0; }
  cv->etamax = 1.0;
  if (*nefPtr <= MXNEF1) {
    cv->eta = 1.0 / (rpower_r(BIAS2 * dsm, 1.0 / cv->L) + ADDON);
    cv->eta = fmax(ETAMIN, fmax(cv->eta, cv->hmin / fabs(cv->h)));
    if (*nefPtr >= SMALL_NEF) cv->eta = fmin(cv->eta, ETAMXF);
    cv_rescale(cv);
    return 0;
  }
  if (cv->q > 1) {
    cv->eta = fmax(ETAMIN, cv->hmin / fabs(cv->h));
    cv_adjust_order(cv, -1);
    cv->L = cv->q; cv->q--; cv->qwait = cv->L;
    cv_rescale(cv);
    return 0;
  }
  cv->eta = fmax(ETAMIN, cv->hmin / fabs(cv->h));
  cv->h *= cv->eta;
  cv->hscale = cv->h;
  cv->qwait = LONG_WAIT;
  cv->f(cv->tn, cv->zn[0], cv->tempv, cv->user); cv->nfe++;
  vscale(cv, cv->h, cv->tempv, cv->zn[1]);
  return 0;
}
static void complete_step(cvmem_t *cv)
{
  int i, j;
  cv->nst++;
  cv->hu = cv->h; cv->qu = cv->q;
  for (i = cv->q; i >= 2; i--) cv->tau[i] = cv->tau[i - 1];
  if ((cv->q == 1) && (cv->nst > 1)) cv->tau[2] = cv->tau[1];
  cv->tau[1] = cv->h;
  for (j = 0; j <= cv->q; j++) linear_sum(cv, cv->l[j], cv->acor, 1.0, cv->zn[j], cv->zn[j]);
  cv->qwait--;
  if ((cv->qwait == 1) && (cv->q != QMAX)) { vscale(cv, 1.0, cv->acor, cv->zn[QMAX]); cv->saved_tq5 = cv->tq[5]; }
}
static void set_eta(cvmem_t *cv)
{
  if (cv->eta < THRESH) { cv->eta = 1.0; cv->hprime = cv->h; }
  else {
    cv->eta = fmin(cv->eta, cv->etamax);
    cv->eta /= fmax(1.0, fabs(cv->h) * cv->hmax_inv * cv->eta);
    cv->hprime = cv->h * cv->eta;
  }
}
static void prepare_next_step(cvmem_t *cv, double dsm)
{
  double etam;
  if (cv->etamax == 1.0) { cv->qwait = cv->qwait > 2 ? cv->qwait : 2; cv->qprime = cv->q; cv->hprime = cv->h; cv->eta = 1.0; return; }
  cv->etaq = 1.0 / (rpower_r(BIAS2 * dsm, 1.0 / cv->L) + ADDON);
  if (cv->qwait != 0) { cv->eta = cv->etaq; cv->qprime = cv->q; set_eta(cv); return; }
  cv->qwait = 2;
  /* etaqm1 */
  cv->etaqm1 = 0.0;
  if (cv->q > 1) {
    double ddn = wrms_norm(cv, cv->zn[cv->q], cv->ewt) / cv->tq[1];
    cv->etaqm1 = 1.0 / (rpower_r(BIAS1 * ddn, 1.0 / cv->q) + ADDON);
  }
  /* etaqp1 */
  cv->etaqp1 = 0.0;
  if (cv->q != QMAX) {
    double cquot = (cv->tq[5] / cv->saved_tq5) * rpower_i(cv->h / cv->tau[2], cv->L), dup;
    linear_sum(cv, -cquot, cv->zn[QMAX], 1.0, cv->acor, cv->tempv);
    dup = wrms_norm(cv, cv->tempv, cv->ewt) / cv->tq[3];
    cv->etaqp1 = 1.0 / (rpower_r(BIAS3 * dup, 1.0 / (cv->L + 1)) + ADDON);
  }
  /* choose */
  etam = fmax(cv->etaqm1, fmax(cv->etaq, cv->etaqp1));
  if (etam < THRESH) { cv->eta = 1.0; cv->qprime = cv->q; }
  else if (etam == cv->etaq) { cv->eta = cv->etaq; cv->qprime = cv->q; }
  else if (etam == cv->etaqm1) { cv->eta = cv->etaqm1; cv->qprime = cv->q - 1; }
  else { cv->eta = cv->etaqp1; cv->qprime = cv->q + 1; vscale(cv, 1.0, cv->acor, cv->zn[QMAX]); }
  set_eta(cv);
}
static int cv_step(cvmem_t *cv)
{
  double saved_t = cv->tn, dsm = 0.0; int ncf = 0, nef = 0, nflag = FIRST_CALL, kflag, passed;
  if ((cv->nst > 0) && (cv->hprime != cv->h)) cv_adjust_params(cv);
  for (;;) {
    cv_predict(cv);
    cv_set(cv);
    nflag = nls_newton(cv, nflag);
    kflag = handle_nflag(cv, &nflag, saved_t, &ncf);
    if (kflag == PREDICT_AGAIN) continue;
    if (kflag != DO_ERROR_TEST) return kflag;
    passed = do_error_test(cv, &nflag, &kflag, saved_t, &nef, &dsm);
    if ((!passed) && (kflag == REP_ERR_FAIL)) return kflag;
    if (passed) break;
  }
  complete_step(cv);
  prepare_next_step(cv, dsm);
  cv->etamax = (cv->nst <= SMALL_NST) ? ETAMX2 : ETAMX3;
  vscale(cv, 1.0 / cv->tq[2], cv->acor, cv->acor);
  return SUCCESS_STEP;
}

static int get_dky0(cvmem_t *cv, double t, double *dky)
{
  double s, tfuzz, tp, tn1; int j;
  tfuzz = FUZZ_FACTOR * UROUND * (fabs(cv->tn) + fabs(cv->hu));
  if (cv->hu < 0.0) tfuzz = -tfuzz;
  tp = cv->tn - cv->hu - tfuzz;
  tn1 = cv->tn + tfuzz;
  if ((t - tp) * (t - tn1) > 0.0) return ORACLE_CV_BAD_T;
  s = (t - cv->tn) / cv->h;
  for (j = cv->q; j >= 0; j--) {
    if (j == cv->q) vscale(cv, 1.0, cv->zn[cv->q], dky);
    else linear_sum(cv, 1.0, cv->zn[j], s, dky, dky);
  }
  return ORACLE_CV_SUCCESS;
}

/* ------------------------------------------------------------------- public */
void *portcv_create(int n, oracle_rhs_fn f, oracle_jac_fn jac, void *user)
{
  int j; cvmem_t *cv = (cvmem_t *)calloc(1, sizeof *cv);
  size_t nb = (size_t)(n > 0 ? n : 1) * sizeof(double);
  cv->n = n; cv->f = f; cv->jac = jac; cv->user = user;
  for (j = 0; j < L_MAX; j++) cv->zn[j] = (double *)calloc(1, nb);
  cv->ewt = (double *)calloc(1, nb); cv->y = (double *)calloc(1, nb); cv->acor = (double *)calloc(1, nb);
  cv->tempv = (double *)calloc(1, nb); cv->ftemp = (double *)calloc(1, nb);
  cv->M = (double *)calloc((size_t)(n > 0 ? n : 1) * (size_t)(n > 0 ? n : 1), sizeof(double));
  cv->savedJ = (double *)calloc((size_t)(n > 0 ? n : 1) * (size_t)(n > 0 ? n : 1), sizeof(double));
  cv->pivots = (long *)calloc((size_t)(n > 0 ? n : 1), sizeof(long));
  return cv;
}
void portcv_free(void *h)
{
  int j; cvmem_t *cv = (cvmem_t *)h;
  if (!cv) return;
  for (j = 0; j < L_MAX; j++) free(cv->zn[j]);
  free(cv->ewt); free(cv->y); free(cv->acor); free(cv->tempv); free(cv->ftemp);
  free(cv->M); free(cv->savedJ); free(cv->pivots); free(cv);
}
static void fold_totals(cvmem_t *cv)
{
  cv->tot[0] += cv->nst; cv->tot[1] += cv->nfe; cv->tot[2] += cv->nsetups; cv->tot[3] += cv->nje;
  cv->tot[4] += cv->nni; cv->tot[5] += cv->ncfn; cv->tot[6] += cv->netf;
}
/* CVodeMalloc / CVodeReInit + tolerances + CVDense: cvode.c:318-479, :496-629 */
int portcv_init(void *h, double t0, const double *y0, double reltol, double abstol, long mxstep, int reinit)
{
  cvmem_t *cv = (cvmem_t *)h; int i;
  if (reinit) fold_totals(cv); else memset(cv->tot, 0, sizeof cv->tot);
  if (reltol < 0.0 || abstol < 0.0) return ORACLE_CV_ILL_INPUT;
  cv->reltol = reltol; cv->abstol = abstol; cv->mxstep = mxstep;
  cv->tn = t0;
  cv->q = 1; cv->L = 2; cv->qwait = cv->L; cv->etamax = ETAMX1;
  cv->qu = 0; cv->hu = 0.0; cv->tolsf = 1.0;
  for (i = 0; i < cv->n; i++) cv->zn[0][i] = y0[i];
  cv->nst = cv->nfe = cv->ncfn = cv->netf = cv->nni = cv->nsetups = cv->nhnil = cv->nstlp = 0;
  cv->nje = cv->nfeD = cv->nstlj = 0;                       /* CVDenseInit */
  cv->hmin = 0.0; cv->hmax_inv = 0.0; cv->tstopset = 0;
  return ORACLE_CV_SUCCESS;
}
void portcv_set_stop_time(void *h, double tstop) { cvmem_t *cv = (cvmem_t *)h; cv->tstop = tstop; cv->tstopset = 1; }

/* CVode(), itask = CV_NORMAL or CV_NORMAL_TSTOP */
int portcv_solve(void *h, double tout, double *yout, double *tret, int use_tstop)
{
  cvmem_t *cv = (cvmem_t *)h;
  long nstloc; int kflag, istate = ORACLE_CV_SUCCESS, ier, istop = use_tstop && cv->tstopset;
  double troundoff, rh;

  if (cv->nst == 0) {
    if (ewt_set(cv, cv->zn[0], cv->ewt) != 0) return ORACLE_CV_ILL_INPUT;
    cv->f(cv->tn, cv->zn[0], cv->zn[1], cv->user); cv->nfe++;
    cv->h = 0.0;
    if (!cv_hin(cv, tout)) return ORACLE_CV_ILL_INPUT;        /* tout too close to t0 */
    rh = fabs(cv->h) * cv->hmax_inv;
    if (rh > 1.0) cv->h /= rh;
    if (fabs(cv->h) < cv->hmin) cv->h *= cv->hmin / fabs(cv->h);
    if (istop) {
      if ((cv->tstop - cv->tn) * cv->h < 0.0) return ORACLE_CV_ILL_INPUT;
      if ((cv->tn + cv->h - cv->tstop) * cv->h > 0.0) cv->h = cv->tstop - cv->tn;
    }
    cv->hscale = cv->h; cv->h0u = cv->h; cv->hprime = cv->h;
    vscale(cv, cv->h, cv->zn[1], cv->zn[1]);
  }
  if (cv->nst > 0) {
    if (istop && ((cv->tstop - cv->tn) * cv->h < 0.0)) return ORACLE_CV_ILL_INPUT;
    if ((cv->tn - tout) * cv->h >= 0.0) {
      *tret = tout;
      ier = get_dky0(cv, tout, yout);
      if (ier != ORACLE_CV_SUCCESS) return ORACLE_CV_ILL_INPUT;
      return ORACLE_CV_SUCCESS;
    }
    if (istop) {
      troundoff = FUZZ_FACTOR * UROUND * (fabs(cv->tn) + fabs(cv->h));
      if (fabs(cv->tn - cv->tstop) <= troundoff) {
        ier = get_dky0(cv, cv->tstop, yout);
        if (ier != ORACLE_CV_SUCCESS) return ORACLE_CV_ILL_INPUT;
        *tret = cv->tstop;
        return ORACLE_CV_TSTOP_RETURN;
      }
      if ((cv->tn + cv->hprime - cv->tstop) * cv->h > 0.0) { cv->hprime = cv->tstop - cv->tn; cv->eta = cv->hprime / cv->h; }
    }
  }
  nstloc = 0;
  for (;;) {
    cv->next_h = cv->h; cv->next_q = cv->q;
    if (cv->nst > 0 && ewt_set(cv, cv->zn[0], cv->ewt) != 0) { istate = ORACLE_CV_ILL_INPUT; *tret = cv->tn; vscale(cv, 1.0, cv->zn[0], yout); break; }
    if (nstloc >= cv->mxstep) { istate = ORACLE_CV_TOO_MUCH_WORK; *tret = cv->tn; vscale(cv, 1.0, cv->zn[0], yout); break; }
    if ((cv->tolsf = UROUND * wrms_norm(cv, cv->zn[0], cv->ewt)) > 1.0) {
      istate = ORACLE_CV_TOO_MUCH_ACC; *tret = cv->tn; vscale(cv, 1.0, cv->zn[0], yout); cv->tolsf *= 2.0; break;
    } else cv->tolsf = 1.0;
    if (cv->tn + cv->h == cv->tn) cv->nhnil++;
    kflag = cv_step(cv);
    if (kflag != SUCCESS_STEP) {
      istate = (kflag == REP_ERR_FAIL) ? ORACLE_CV_ERR_FAILURE : (kflag == REP_CONV_FAIL) ? ORACLE_CV_CONV_FAILURE
             : (kflag == SETUP_FAILED) ? ORACLE_CV_LSETUP_FAIL : ORACLE_CV_LSOLVE_FAIL;
      *tret = cv->tn; vscale(cv, 1.0, cv->zn[0], yout);
      break;
    }
    nstloc++;
    if (istop) {
      troundoff = FUZZ_FACTOR * UROUND * (fabs(cv->tn) + fabs(cv->h));
      if (fabs(cv->tn - cv->tstop) <= troundoff) {
        (void)get_dky0(cv, cv->tstop, yout);
        *tret = cv->tstop; istate = ORACLE_CV_TSTOP_RETURN;
        break;
      }
      if ((cv->tn + cv->hprime - cv->tstop) * cv->h > 0.0) { cv->hprime = cv->tstop - cv->tn; cv->eta = cv->hprime / cv->h; }
    }
    if ((cv->tn - tout) * cv->h >= 0.0) {
      istate = ORACLE_CV_SUCCESS; *tret = tout;
      (void)get_dky0(cv, tout, yout);
      cv->next_q = cv->qprime; cv->next_h = cv->hprime;
      break;
    }
  }
  return istate;
}
void portcv_stats(void *h, long *s)
{
  cvmem_t *cv = (cvmem_t *)h;
  s[0] = cv->tot[0] + cv->nst; s[1] = cv->tot[1] + cv->nfe; s[2] = cv->tot[2] + cv->nsetups; s[3] = cv->tot[3] + cv->nje;
  s[4] = cv->tot[4] + cv->nni; s[5] = cv->tot[5] + cv->ncfn; s[6] = cv->tot[6] + cv->netf;
}
