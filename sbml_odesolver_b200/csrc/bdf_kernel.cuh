/*
 * bdf_kernel.cuh -- batched variable-order (1..5) variable-step BDF integrator,
 * one trajectory per thread, written for sm_100a (B200).
 *
 * This text is concatenated after the model section produced by
 * ODEModel_generateCUDASource (emit.c) and compiled at run time with NVRTC
 * (-arch=sm_100a).  The model section supplies NEQ/NVAL/... and the device
 * functions model_init, ode_f, jacobi_f, event_f, store_row, store_row0,
 * so the right-hand side, the sparse analytic Jacobian and the event logic are
 * inlined and every loop over the state dimension is unrolled at compile time.
 *
 * What it replaces in the reference: the per-design-point call chain
 *   IntegratorInstance_cvodeOneStep -> CVode -> CVStep -> CVnlsNewton ->
 *   CVDenseSetup/CVDenseSolve -> gefa/gesl, plus IntegratorInstance_updateData
 *   (src/cvodeSolver.c:81-240, src/integratorInstance.c:1029-1237, vendored
 *   CVODE 2.3.0 cvode.c:875-2713, cvdense.c:368-441, smalldense.c:58-160).
 * The controller (error test, step/order selection, Newton convergence test,
 * Jacobian/LU reuse rules, initial step) uses CVODE's constants and decision
 * order so that each trajectory takes the same steps as the reference's
 * CVODE integration of the same parameter set.
 *
 * Data layout: everything per trajectory that lives in global memory is
 * structure-of-arrays with the batch index fastest --
 *   vary[j*stride + set], out[(row*NSTORE + k)*stride + set], stats[k*stride + set]
 * -- so a warp reads/writes 32 consecutive doubles (256 B, fully coalesced).
 * On chip: Nordsieck array, weights and Newton vectors are registers (all
 * indices are compile-time constants); the Newton matrix M = I - gamma*J, the
 * saved sparse Jacobian and the LU right-hand side live in shared memory,
 * element-major / thread-minor (word k of thread t at smem[k*blockDim.x + t]),
 * which makes the run-time pivot indexing of the LU bank-conflict free.
 *
 * Build knobs (defined by the host before this text):
 *   ODES_BLOCK        threads per block
 *   ODES_MINBLOCKS    resident blocks per SM for __launch_bounds__
 *   ODES_MAT_LOCAL    1: keep M/J/b in local memory instead of shared memory
 */
#ifndef ODES_BLOCK
#define ODES_BLOCK 64
#endif
#ifndef ODES_MINBLOCKS
#define ODES_MINBLOCKS 1
#endif
#ifndef ODES_MAT_LOCAL
#define ODES_MAT_LOCAL 0
#endif

#define QMAX 5
#define LMAX 6
#define UROUND 2.2204460492503131e-16

/* CVODE control constants (cvode.c:40-143, cvdense.h:43-44) */
#define K_FUZZ_FACTOR 100.0
#define K_HLB_FACTOR 100.0
#define K_HUB_FACTOR 0.1
#define K_H_BIAS 0.5
#define K_MAX_ITERS 4
#define K_CORTES 0.1
#define K_THRESH 1.5
#define K_ETAMX1 10000.0
#define K_ETAMX2 10.0
#define K_ETAMX3 10.0
#define K_ETAMXF 0.2
#define K_ETAMIN 0.1
#define K_ETACF 0.25
#define K_ADDON 0.000001
#define K_BIAS1 6.0
#define K_BIAS2 6.0
#define K_BIAS3 10.0
#define K_ONEPSM 1.000001
#define K_SMALL_NST 10
#define K_MXNCF 10
#define K_MXNEF 7
#define K_MXNEF1 3
#define K_SMALL_NEF 2
#define K_LONG_WAIT 10
#define K_NLS_MAXCOR 3
#define K_CRDOWN 0.3
#define K_DGMAX 0.3
#define K_RDIV 2.0
#define K_MSBP 20
#define K_MSBJ 50
#define K_DGMAX_J 0.2
#define K_MIN_INC_MULT 1000.0

/* CVODE flags, numbering of the vendored cvode.h:712-725 */
#define CV_SUCCESS 0
#define CV_TSTOP_RETURN 1
#define CV_ILL_INPUT (-2)
#define CV_TOO_MUCH_WORK (-4)
#define CV_TOO_MUCH_ACC (-5)
#define CV_ERR_FAILURE (-6)
#define CV_CONV_FAILURE (-7)
#define CV_LSETUP_FAIL (-9)
#define CV_LSOLVE_FAIL (-10)

#define FIRST_CALL 0
#define PREV_CONV_FAIL (-1)
#define PREV_ERR_FAIL (-2)
#define CV_NO_FAILURES 0
#define CV_FAIL_BAD_J 1
#define CV_FAIL_OTHER 2

struct OdesArgs {
  const double *vary;          /* [NVARY][stride]                       */
  double *out;                 /* [nout+1][NSTORE][stride]              */
  int *status;                 /* [stride]                              */
  int *stats;                  /* [7][stride]  nst,nfe,nsetups,nje,nni,ncfn,netf */
  unsigned int *fired;         /* [nout+1][stride] event bitmasks       */
  const double *tpoints;       /* [nout+1] output grid                  */
  unsigned long long stride;
  int nsets, nout, mxstep, pad;
  double reltol, abstol;
};

#define FORI _Pragma("unroll") for (int i = 0; i < NEQ; i++)
#define NMAT (NEQ * NEQ)
#if HAS_JAC
#define NSJ NNZ
#else
#define NSJ NMAT
#endif
/* shared-memory words per thread: M, saved Jacobian, LU right-hand side */
#define ZOFF (NMAT + NSJ + NEQ)                   /* zn[2..5] */
#define PVOFF (ZOFF + (LMAX - 2) * NEQ)           /* per-trajectory constants */
#define EWOFF (PVOFF + NPV)                       /* error weights */
#define TAUOFF (EWOFF + NEQ)                      /* tau[0..6], tq[0..5], l[0..5] */
#define ODES_SMEM_WORDS (TAUOFF + (LMAX + 1) + 6 + LMAX)

DEV double rpower_r(double base, double e) { return base <= 0.0 ? 0.0 : pow(base, e); }
DEV double rsqrt_(double x) { return x <= 0.0 ? 0.0 : sqrt(x); }

struct Solver {
  /* Nordsieck history and work vectors */
  double z0[A1(NEQ)], z1[A1(NEQ)];            /* z0, z1 in registers; zn[2..5] in shared memory */
  double acor[A1(NEQ)], tempv[A1(NEQ)], ftemp[A1(NEQ)], y[A1(NEQ)];   /* ewt, tau, tq, l: shared memory */
  PvRef pv;
  double h, hprime, eta, hscale, tn, rl1, gamma, gammap, gamrat, crate, acnrm, etamax, hu, saved_tq5;
  double reltol, abstol, tstop;
  int q, qprime, qwait, L, jcur;
  int nst, nstlp, nstlj;
  int c_nst, c_nfe, c_nsetups, c_nje, c_nni, c_ncfn, c_netf;      /* totals over re-inits */
#if ODES_MAT_LOCAL
  double mat[ODES_SMEM_WORDS];
#define SM(k) mat[(k)]
#else
  double *sm; unsigned int sstride;
#define SM(k) sm[(size_t)(k) * sstride]
#endif
#define MAT(i, j) SM((j) * NEQ + (i))          /* column-major like DenseMat */
#define SJ(k) SM(NMAT + (k))
#define BV(i) SM(NMAT + NSJ + (i))
#define EW(i) SM(EWOFF + (i))
#define TAU(j) SM(TAUOFF + (j))
#define TQ(j) SM(TAUOFF + (LMAX + 1) + (j))
#define LL(j) SM(TAUOFF + (LMAX + 1) + 6 + (j))
/* Nordsieck row j, component i: j is a compile-time constant at every use (unrolled loops) */
#define Z(j, i) (*((j) == 0 ? &z0[i] : ((j) == 1 ? &z1[i] : &SM(ZOFF + ((j) - 2) * NEQ + (i)))))
  int piv[A1(NEQ)];

  DEV void f(double t, const double (&yy)[A1(NEQ)], double (&out)[A1(NEQ)]) { ode_f(t, yy, pv, out); c_nfe++; }

  DEV double wrms(const double (&x)[A1(NEQ)]) const
  {
    double sum = 0.0;
    FORI { const double p = x[i] * EW(i); sum += p * p; }
    return rsqrt_(sum / NEQ);
  }
  DEV bool ewt_set(const double (&ycur)[A1(NEQ)])
  {
    double mn = 1.0; bool first = true;
    FORI { tempv[i] = reltol * fabs(ycur[i]) + abstol; if (first || tempv[i] < mn) mn = tempv[i]; first = false; }
    if (mn <= 0.0) return false;
    FORI EW(i) = 1.0 / tempv[i];
    return true;
  }

  /* ---- history access with a run-time order: unrolled selects keep zn in registers ---- */
  DEV void clear_history()
  {
    _Pragma("unroll") for (int j = 0; j <= LMAX; j++) TAU(j) = 0.0;
    _Pragma("unroll") for (int j = 0; j < 6; j++) TQ(j) = 0.0;
    _Pragma("unroll") for (int j = 0; j < LMAX; j++) LL(j) = 0.0;
    _Pragma("unroll") for (int j = 0; j < LMAX; j++) { FORI Z(j, i) = 0.0; }
  }
  DEV void zn_get(int j, double (&out)[A1(NEQ)])
  {
    _Pragma("unroll") for (int jj = 0; jj < LMAX; jj++) if (jj == j) { FORI out[i] = Z(jj, i); }
  }

  DEV void rescale()
  {
    double factor = eta;
    _Pragma("unroll") for (int j = 1; j <= QMAX; j++) if (j <= q) { FORI Z(j, i) = factor * Z(j, i); factor *= eta; }
    h = hscale * eta;
    hscale = h;
  }
  DEV void predict()
  {
    tn += h;
    _Pragma("unroll") for (int k = 1; k <= QMAX; k++)
      _Pragma("unroll") for (int j = QMAX; j >= k; j--)
        if (k <= q && j <= q) { FORI Z(j - 1, i) = Z(j - 1, i) + Z(j, i); }
  }
  DEV void restore(double saved_t)
  {
    tn = saved_t;
    _Pragma("unroll") for (int k = 1; k <= QMAX; k++)
      _Pragma("unroll") for (int j = QMAX; j >= k; j--)
        if (k <= q && j <= q) { FORI Z(j - 1, i) = Z(j - 1, i) - Z(j, i); }
  }
  DEV void increase_bdf()
  {
    double alpha0, alpha1, prod, xi, xiold, hsum, A1c;
    _Pragma("unroll") for (int i = 0; i <= QMAX; i++) LL(i) = 0.0;
    LL(2) = alpha1 = prod = xiold = 1.0;
    alpha0 = -1.0;
    hsum = hscale;
    _Pragma("unroll") for (int j = 1; j < QMAX; j++) if (j < q) {
      hsum += TAU(j + 1);
      xi = hsum / hscale;
      prod *= xi;
      alpha0 -= 1.0 / (j + 1);
      alpha1 += 1.0 / xi;
      _Pragma("unroll") for (int i = QMAX; i >= 2; i--) if (i <= j + 2) LL(i) = LL(i) * xiold + LL(i - 1);
      xiold = xi;
    }
    A1c = (-alpha0 - alpha1) / prod;
    /* zn[L] = A1 * zn[qmax]; zn[j] += LL(j) * zn[L], j = 2..q   (L = q+1 <= 5 here) */
    double znL[A1(NEQ)];
    FORI znL[i] = A1c * Z(QMAX, i);
    _Pragma("unroll") for (int j = 2; j <= QMAX; j++) {
      if (j == L) { FORI Z(j, i) = znL[i]; }
      else if (j <= q) { FORI Z(j, i) = LL(j) * znL[i] + Z(j, i); }
    }
  }
  DEV void decrease_bdf()
  {
    double hsum, xi;
    _Pragma("unroll") for (int i = 0; i <= QMAX; i++) LL(i) = 0.0;
    LL(2) = 1.0;
    hsum = 0.0;
    _Pragma("unroll") for (int j = 1; j <= QMAX - 2; j++) if (j <= q - 2) {
      hsum += TAU(j);
      xi = hsum / hscale;
      _Pragma("unroll") for (int i = QMAX; i >= 2; i--) if (i <= j + 2) LL(i) = LL(i) * xi + LL(i - 1);
    }
    double znq[A1(NEQ)];
    zn_get(q, znq);
    _Pragma("unroll") for (int j = 2; j < QMAX; j++) if (j < q) { FORI Z(j, i) = (-LL(j)) * znq[i] + Z(j, i); }
  }
  DEV void adjust_order(int deltaq)
  {
    if ((q == 2) && (deltaq != 1)) return;
    if (deltaq == 1) increase_bdf(); else if (deltaq == -1) decrease_bdf();
  }
  DEV void adjust_params()
  {
    if (qprime != q) { adjust_order(qprime - q); q = qprime; L = q + 1; qwait = L; }
    rescale();
  }
  DEV void set_bdf()
  {
    double alpha0, alpha0_hat, xi_inv, xistar_inv, hsum, A1c, A2, A3, A4, A5, A6, C, CPrime, CPrimePrime;
    LL(0) = LL(1) = xi_inv = xistar_inv = 1.0;
    _Pragma("unroll") for (int i = 2; i <= QMAX; i++) if (i <= q) LL(i) = 0.0;
    alpha0 = alpha0_hat = -1.0;
    hsum = h;
    double lq = 1.0;                                 /* LL(q) */
    if (q > 1) {
      _Pragma("unroll") for (int j = 2; j < QMAX; j++) if (j < q) {
        hsum += TAU(j - 1);
        xi_inv = h / hsum;
        alpha0 -= 1.0 / j;
        _Pragma("unroll") for (int i = QMAX; i >= 1; i--) if (i <= j) LL(i) += LL(i - 1) * xi_inv;
      }
      alpha0 -= 1.0 / q;
      xistar_inv = -LL(1) - alpha0;
      double tauqm1 = 0.0, tauq = 0.0;
      _Pragma("unroll") for (int j = 1; j <= QMAX; j++) { if (j == q - 1) tauqm1 = TAU(j); }
      hsum += tauqm1;
      xi_inv = h / hsum;
      alpha0_hat = -LL(1) - xi_inv;
      _Pragma("unroll") for (int i = QMAX; i >= 1; i--) if (i <= q) LL(i) += LL(i - 1) * xistar_inv;
      (void)tauq;
    }
    _Pragma("unroll") for (int j = 1; j <= QMAX; j++) if (j == q) lq = LL(j);
    A1c = 1.0 - alpha0_hat + alpha0;
    A2 = 1.0 + q * A1c;
    TQ(2) = fabs(alpha0 * (A2 / A1c));
    TQ(5) = fabs((A2) / (lq * xi_inv / xistar_inv));
    if (qwait == 1) {
      double tauq2 = 0.0;
      _Pragma("unroll") for (int j = 1; j <= LMAX; j++) if (j == q) tauq2 = TAU(j);
      C = xistar_inv / lq;
      A3 = alpha0 + 1.0 / q;
      A4 = alpha0_hat + xi_inv;
      CPrime = A3 / (1.0 - A4 + A3);
      TQ(1) = fabs(CPrime / C);
      hsum += tauq2;
      xi_inv = h / hsum;
      A5 = alpha0 - (1.0 / (q + 1));
      A6 = alpha0_hat - xi_inv;
      CPrimePrime = A2 / (1.0 - A6 + A5);
      TQ(3) = fabs(CPrimePrime * xi_inv * (q + 2) * A5);
    }
    TQ(4) = K_CORTES * TQ(2);
  }
  DEV void set_coeffs()
  {
    set_bdf();
    rl1 = 1.0 / LL(1);
    gamma = h * rl1;
    if (nst == 0) gammap = gamma;
    gamrat = (nst > 0) ? gamma / gammap : 1.0;
  }

  /* ---- dense linear algebra on the strided shared-memory matrix (gefa / gesl) ---- */
  DEV int gefa()
  {
    _Pragma("unroll") for (int k = 0; k < NEQ - 1; k++) {
      int lp = k;
      double big = fabs(MAT(k, k));
      _Pragma("unroll") for (int i = k + 1; i < NEQ; i++) { const double v = fabs(MAT(i, k)); if (v > big) { big = v; lp = i; } }
      piv[k] = lp;
      const double pval = MAT(lp, k);
      if (pval == 0.0) return k + 1;
      const bool swap = (lp != k);
      if (swap) { MAT(lp, k) = MAT(k, k); MAT(k, k) = pval; }
      const double mult = -1.0 / pval;
      _Pragma("unroll") for (int i = k + 1; i < NEQ; i++) MAT(i, k) = MAT(i, k) * mult;
      _Pragma("unroll") for (int j = k + 1; j < NEQ; j++) {
        const double a_kj = MAT(lp, j);
        if (swap) { MAT(lp, j) = MAT(k, j); MAT(k, j) = a_kj; }
        if (a_kj != 0.0) { _Pragma("unroll") for (int i = k + 1; i < NEQ; i++) MAT(i, j) = MAT(i, j) + a_kj * MAT(i, k); }
      }
    }
    piv[NEQ - 1] = NEQ - 1;
    if (MAT(NEQ - 1, NEQ - 1) == 0.0) return NEQ;
    return 0;
  }
  /* solves M x = b in place; b is staged in shared memory so the pivot row may be indexed at run time */
  DEV void gesl(double (&b)[A1(NEQ)])
  {
    FORI BV(i) = b[i];
    _Pragma("unroll") for (int k = 0; k < NEQ - 1; k++) {
      const int lp = piv[k];
      const double mult = BV(lp);
      if (lp != k) { BV(lp) = BV(k); BV(k) = mult; }
      _Pragma("unroll") for (int i = k + 1; i < NEQ; i++) BV(i) = BV(i) + mult * MAT(i, k);
    }
    _Pragma("unroll") for (int k = NEQ - 1; k >= 0; k--) {
      const double bk = BV(k) / MAT(k, k);
      BV(k) = bk;
      const double mult = -bk;
      _Pragma("unroll") for (int i = 0; i < k; i++) BV(i) = BV(i) + mult * MAT(i, k);
    }
    FORI b[i] = BV(i);
  }

  /* CVDenseSetup (cvdense.c:368-412) */
  DEV int lsetup(int convfail)
  {
    const double dgamma = fabs((gamma / gammap) - 1.0);
    const bool jbad = (nst == 0) || (nst > nstlj + K_MSBJ) || ((convfail == CV_FAIL_BAD_J) && (dgamma < K_DGMAX_J)) || (convfail == CV_FAIL_OTHER);
    const double mg = -gamma;
#if HAS_JAC
    if (jbad) {
      c_nje++; nstlj = nst; jcur = 1;
      double jnz[A1(NNZ)];
      jacobi_f(tn, z0, pv, jnz);
      _Pragma("unroll") for (int k = 0; k < NNZ; k++) SJ(k) = jnz[k];
    } else jcur = 0;
    /* M = I - gamma*J: zero pattern first, then the stored non-zeros (DenseScale + DenseAddI order) */
    _Pragma("unroll") for (int k = 0; k < NMAT; k++) SM(k) = 0.0 * mg;
    ODES_SCATTER_J(mg)
    FORI MAT(i, i) = MAT(i, i) + 1.0;
#else
    if (jbad) {
      /* difference-quotient Jacobian, cvdense.c:479-537; vtemp1 = acor is the scratch f vector */
      c_nje++; nstlj = nst; jcur = 1;
      const double srur = rsqrt_(UROUND);
      const double fnorm = wrms(ftemp);
      const double minInc = (fnorm != 0.0) ? (K_MIN_INC_MULT * fabs(h) * UROUND * NEQ * fnorm) : 1.0;
      double yv[A1(NEQ)];
      FORI yv[i] = Z(0, i);
      _Pragma("unroll") for (int j = 0; j < NEQ; j++) {
        const double yjsaved = yv[j];
        const double inc = fmax(srur * fabs(yjsaved), minInc / EW(j));
        yv[j] += inc;
        ode_f(tn, yv, pv, acor);
        yv[j] = yjsaved;
        const double inc_inv = 1.0 / inc;
        FORI SJ(j * NEQ + i) = inc_inv * (acor[i] - ftemp[i]);
      }
    } else jcur = 0;
    _Pragma("unroll") for (int k = 0; k < NMAT; k++) SM(k) = SJ(k) * mg;
    FORI MAT(i, i) = MAT(i, i) + 1.0;
#endif
    return gefa() > 0 ? 1 : 0;
  }
  DEV void lsolve(double (&b)[A1(NEQ)])
  {
    gesl(b);
    if (gamrat != 1.0) { const double c = 2.0 / (1.0 + gamrat); FORI b[i] = c * b[i]; }
  }

  /* ---- CVHin (cvode.c:1514-1636) ---- */
  DEV bool hin(double tout)
  {
    const double tdiff = tout - tn;
    if (tdiff == 0.0) return false;
    const int sign = (tdiff > 0.0) ? 1 : -1;
    const double tdist = fabs(tdiff);
    const double tround = UROUND * fmax(fabs(tn), fabs(tout));
    if (tdist < 2.0 * tround) return false;
    const double hlb = K_HLB_FACTOR * tround;
    /* CVUpperBoundH0 */
    double hub_inv = 0.0;
    FORI {
      double t1 = reltol * fabs(Z(0, i)) + abstol;
      t1 = 1.0 / t1; t1 = 1.0 / t1;                               /* efun then N_VInv */
      t1 = K_HUB_FACTOR * fabs(Z(0, i)) + t1;
      const double r = fabs(Z(1, i)) / t1;
      if (fabs(r) > hub_inv) hub_inv = fabs(r);
    }
    double hub = K_HUB_FACTOR * tdist;
    if (hub * hub_inv > 1.0) hub = 1.0 / hub_inv;
    double hg = rsqrt_(hlb * hub);
    if (hub < hlb) { if (sign == -1) hg = -hg; h = hg; return true; }
    int count = 0; double hnew = 0.0;
    for (;;) {
      const double hgs = hg * sign;
      /* CVYddNorm */
      FORI y[i] = hgs * Z(1, i) + Z(0, i);
      f(tn + hgs, y, tempv);
      FORI tempv[i] = tempv[i] - Z(1, i);
      { const double c = 1.0 / hgs; FORI tempv[i] = c * tempv[i]; }
      const double yddnrm = wrms(tempv);
      hnew = (yddnrm * hub * hub > 2.0) ? rsqrt_(2.0 / yddnrm) : rsqrt_(hg * hub);
      count++;
      if (count >= K_MAX_ITERS) break;
      const double hrat = hnew / hg;
      if ((hrat > 0.5) && (hrat < 2.0)) break;
      if ((count >= 2) && (hrat > 2.0)) { hnew = hg; break; }
      hg = hnew;
    }
    double h0 = K_H_BIAS * hnew;
    if (h0 < hlb) h0 = hlb;
    if (h0 > hub) h0 = hub;
    if (sign == -1) h0 = -h0;
    h = h0;
    return true;
  }

  /* ---- nonlinear solver: CVnlsNewton + CVNewtonIteration (cvode.c:2234-2371) ----
     returns 0 solved, -1 recoverable convergence failure, -2 setup failed unrecoverably */
  DEV int nls(int nflag)
  {
    int convfail = ((nflag == FIRST_CALL) || (nflag == PREV_ERR_FAIL)) ? CV_NO_FAILURES : CV_FAIL_OTHER;
    bool callSetup = (nflag == PREV_CONV_FAIL) || (nflag == PREV_ERR_FAIL) || (nst == 0) ||
                     (nst >= nstlp + K_MSBP) || (fabs(gamrat - 1.0) > K_DGMAX);
    for (;;) {
      f(tn, z0, ftemp);
      if (callSetup) {
        const int ier = lsetup(convfail);
        c_nsetups++;
        callSetup = false;
        gamrat = crate = 1.0;
        gammap = gamma;
        nstlp = nst;
        if (ier > 0) return -1;
      }
      FORI { acor[i] = 0.0; y[i] = Z(0, i); }
      /* Newton iteration */
      int m = 0, ret; double del, delp = 0.0;
      for (;;) {
        if (rl1 == 1.0) { FORI tempv[i] = Z(1, i) + acor[i]; } else { FORI tempv[i] = rl1 * Z(1, i) + acor[i]; }
        if (gamma == 1.0) { FORI tempv[i] = ftemp[i] - tempv[i]; } else { FORI tempv[i] = gamma * ftemp[i] - tempv[i]; }
        lsolve(tempv);
        c_nni++;
        del = wrms(tempv);
        FORI { acor[i] = acor[i] + tempv[i]; y[i] = Z(0, i) + acor[i]; }
        if (m > 0) crate = fmax(K_CRDOWN * crate, del / delp);
        const double dcon = del * fmin(1.0, crate) / TQ(4);
        if (dcon <= 1.0) { acnrm = (m == 0) ? del : wrms(acor); jcur = 0; ret = 0; break; }
        m++;
        if ((m == K_NLS_MAXCOR) || ((m >= 2) && (del > K_RDIV * delp))) { ret = jcur ? -1 : 99; break; }
        delp = del;
        f(tn, y, ftemp);
      }
      if (ret != 99) return ret;
      callSetup = true;
      convfail = CV_FAIL_BAD_J;
    }
  }

  /* ---- one internal step, CVStep (cvode.c:1657-1710); returns 0 or a CV_* failure flag ---- */
  DEV int step()
  {
    const double saved_t = tn;
    int ncf = 0, nef = 0, nflag = FIRST_CALL;
    double dsm = 0.0;
    if ((nst > 0) && (hprime != h)) adjust_params();
    for (;;) {
      predict();
      set_coeffs();
      const int nl = nls(nflag);
      if (nl != 0) {
        /* CVHandleNFlag */
        c_ncfn++;
        restore(saved_t);
        if (nl == -2) return CV_LSETUP_FAIL;
        ncf++;
        etamax = 1.0;
        if ((fabs(h) <= 0.0) || (ncf == K_MXNCF)) return CV_CONV_FAILURE;        /* hmin = 0 */
        eta = K_ETACF;                                                          /* max(ETACF, hmin/|h|) */
        nflag = PREV_CONV_FAIL;
        rescale();
        continue;
      }
      /* CVDoErrorTest */
      dsm = acnrm / TQ(2);
      if (dsm <= 1.0) break;
      nef++; c_netf++;
      nflag = PREV_ERR_FAIL;
      restore(saved_t);
      if ((fabs(h) <= 0.0) || (nef == K_MXNEF)) return CV_ERR_FAILURE;
      etamax = 1.0;
      if (nef <= K_MXNEF1) {
        eta = 1.0 / (rpower_r(K_BIAS2 * dsm, 1.0 / L) + K_ADDON);
        eta = fmax(K_ETAMIN, fmax(eta, 0.0));
        if (nef >= K_SMALL_NEF) eta = fmin(eta, K_ETAMXF);
        rescale();
        continue;
      }
      if (q > 1) {
        eta = K_ETAMIN;
        adjust_order(-1);
        L = q; q--; qwait = L;
        rescale();
        continue;
      }
      eta = K_ETAMIN;
      h *= eta;
      hscale = h;
      qwait = K_LONG_WAIT;
      f(tn, z0, tempv);
      FORI Z(1, i) = h * tempv[i];
    }
    /* CVCompleteStep */
    nst++; c_nst++;
    hu = h;
    _Pragma("unroll") for (int i = QMAX; i >= 2; i--) if (i <= q) TAU(i) = TAU(i - 1);
    if ((q == 1) && (nst > 1)) TAU(2) = TAU(1);
    TAU(1) = h;
    _Pragma("unroll") for (int j = 0; j <= QMAX; j++) if (j <= q) {
      const double lj = LL(j);
      if (lj == 1.0) { FORI Z(j, i) = Z(j, i) + acor[i]; }
      else { FORI Z(j, i) = lj * acor[i] + Z(j, i); }
    }
    qwait--;
    if ((qwait == 1) && (q != QMAX)) { FORI Z(QMAX, i) = acor[i]; saved_tq5 = TQ(5); }
    /* CVPrepareNextStep */
    if (etamax == 1.0) { qwait = qwait > 2 ? qwait : 2; qprime = q; hprime = h; eta = 1.0; }
    else {
      const double etaq = 1.0 / (rpower_r(K_BIAS2 * dsm, 1.0 / L) + K_ADDON);
      if (qwait != 0) { eta = etaq; qprime = q; }
      else {
        qwait = 2;
        double etaqm1 = 0.0, etaqp1 = 0.0;
        if (q > 1) {
          double znq[A1(NEQ)]; zn_get(q, znq);
          const double ddn = wrms(znq) / TQ(1);
          etaqm1 = 1.0 / (rpower_r(K_BIAS1 * ddn, 1.0 / q) + K_ADDON);
        }
        if (q != QMAX) {
          double hr = h / TAU(2), pw = 1.0;
          _Pragma("unroll") for (int i = 1; i <= LMAX; i++) if (i <= L) pw *= hr;
          const double cquot = (TQ(5) / saved_tq5) * pw;
          FORI tempv[i] = (-cquot) * Z(QMAX, i) + acor[i];
          const double dup = wrms(tempv) / TQ(3);
          etaqp1 = 1.0 / (rpower_r(K_BIAS3 * dup, 1.0 / (L + 1)) + K_ADDON);
        }
        const double etam = fmax(etaqm1, fmax(etaq, etaqp1));
        if (etam < K_THRESH) { eta = 1.0; qprime = q; }
        else if (etam == etaq) { eta = etaq; qprime = q; }
        else if (etam == etaqm1) { eta = etaqm1; qprime = q - 1; }
        else { eta = etaqp1; qprime = q + 1; FORI Z(QMAX, i) = acor[i]; }
      }
      /* CVSetEta (hmax_inv = 0) */
      if (eta < K_THRESH) { eta = 1.0; hprime = h; }
      else { eta = fmin(eta, etamax); eta /= fmax(1.0, 0.0); hprime = h * eta; }
    }
    etamax = (nst <= K_SMALL_NST) ? K_ETAMX2 : K_ETAMX3;
    /* acor is rescaled to the local error estimate in CVODE; it is not read again before being reset */
    return CV_SUCCESS;
  }

  /* CVodeGetDky with k = 0 (cvode.c:1228-1279): Horner on the Nordsieck array */
  DEV bool get_dky0(double t, double (&dky)[A1(NEQ)])
  {
    double tfuzz = K_FUZZ_FACTOR * UROUND * (fabs(tn) + fabs(hu));
    if (hu < 0.0) tfuzz = -tfuzz;
    const double tp = tn - hu - tfuzz, tn1 = tn + tfuzz;
    if ((t - tp) * (t - tn1) > 0.0) return false;
    const double s = (t - tn) / h;
    _Pragma("unroll") for (int j = QMAX; j >= 0; j--) {
      if (j == q) { FORI dky[i] = Z(j, i); }
      else if (j < q) {
        if (s == 1.0) { FORI dky[i] = Z(j, i) + dky[i]; }
        else if (s == -1.0) { FORI dky[i] = Z(j, i) - dky[i]; }
        else { FORI dky[i] = s * dky[i] + Z(j, i); }
      }
    }
    return true;
  }

  /* CVodeMalloc / CVodeReInit state (cvode.c:416-452, :588-617) + CVDenseInit */
  DEV void reinit(double t0, const double (&y0)[A1(NEQ)])
  {
    tn = t0;
    q = 1; L = 2; qwait = L; etamax = K_ETAMX1;
    hu = 0.0;
    FORI Z(0, i) = y0[i];
    nst = 0; nstlp = 0; nstlj = 0;
  }

  /* CVode(), CV_NORMAL or CV_NORMAL_TSTOP (cvode.c:875-1210) */
  DEV int solve(double tout, double (&yout)[A1(NEQ)], double &tret, int mxstep)
  {
    if (nst == 0) {
      if (!ewt_set(z0)) return CV_ILL_INPUT;
      f(tn, z0, z1);
      h = 0.0;
      if (!hin(tout)) return CV_ILL_INPUT;
#if USE_TSTOP
      if ((tstop - tn) * h < 0.0) return CV_ILL_INPUT;
      if ((tn + h - tstop) * h > 0.0) h = tstop - tn;
#endif
      hscale = h; hprime = h;
      FORI Z(1, i) = h * Z(1, i);
    }
    if (nst > 0) {
#if USE_TSTOP
      if ((tstop - tn) * h < 0.0) return CV_ILL_INPUT;
#endif
      if ((tn - tout) * h >= 0.0) {
        tret = tout;
        return get_dky0(tout, yout) ? CV_SUCCESS : CV_ILL_INPUT;
      }
#if USE_TSTOP
      {
        const double troundoff = K_FUZZ_FACTOR * UROUND * (fabs(tn) + fabs(h));
        if (fabs(tn - tstop) <= troundoff) {
          if (!get_dky0(tstop, yout)) return CV_ILL_INPUT;
          tret = tstop;
          return CV_TSTOP_RETURN;
        }
        if ((tn + hprime - tstop) * h > 0.0) { hprime = tstop - tn; eta = hprime / h; }
      }
#endif
    }
    int nstloc = 0;
    for (;;) {
      if (nst > 0 && !ewt_set(z0)) { tret = tn; FORI yout[i] = Z(0, i); return CV_ILL_INPUT; }
      if (nstloc >= mxstep) { tret = tn; FORI yout[i] = Z(0, i); return CV_TOO_MUCH_WORK; }
      if (UROUND * wrms(z0) > 1.0) { tret = tn; FORI yout[i] = Z(0, i); return CV_TOO_MUCH_ACC; }
      const int kflag = step();
      if (kflag != CV_SUCCESS) { tret = tn; FORI yout[i] = Z(0, i); return kflag; }
      nstloc++;
#if USE_TSTOP
      {
        const double troundoff = K_FUZZ_FACTOR * UROUND * (fabs(tn) + fabs(h));
        if (fabs(tn - tstop) <= troundoff) { (void)get_dky0(tstop, yout); tret = tstop; return CV_TSTOP_RETURN; }
        if ((tn + hprime - tstop) * h > 0.0) { hprime = tstop - tn; eta = hprime / h; }
      }
#endif
      if ((tn - tout) * h >= 0.0) { tret = tout; (void)get_dky0(tout, yout); return CV_SUCCESS; }
    }
  }
};

/* One trajectory: reset, set varied values, integrate over the output grid with the reference's
   output-time event handling (src/odeSolver.c:313-330, src/integratorInstance.c:1088-1237). */
DEV void integrate_set(const OdesArgs &a, const size_t set, double *sm, unsigned int sstride)
{
  Solver s;
#if !ODES_MAT_LOCAL
  s.sm = sm; s.sstride = sstride;
  s.pv.p = sm + (size_t)PVOFF * sstride; s.pv.s = sstride;
#else
  (void)sm; (void)sstride;
  s.pv.p = s.mat + PVOFF; s.pv.s = 1u;
#endif
  s.reltol = a.reltol; s.abstol = a.abstol;
  s.c_nst = s.c_nfe = s.c_nsetups = s.c_nje = s.c_nni = s.c_ncfn = s.c_netf = 0;
  s.jcur = 0; s.gammap = 1.0; s.gamrat = 1.0; s.crate = 1.0; s.saved_tq5 = 1.0; s.acnrm = 0.0;
  s.clear_history();

  const size_t stride = (size_t)a.stride;
  double val[A1(NEQ)];                 /* data->value[0..neq) */
  int trigger[A1(NEVENTS)];
  model_init(val, s.pv, a.vary, stride, set);
  _Pragma("unroll") for (int e = 0; e < NEVENTS; e++) trigger[e] = c_trigger0[e];
  store_row0(a.out + set, stride, val, s.pv);
  if (a.fired) a.fired[set] = 0u;
  {
    /* static assigned values (rules not recomputed before the ODEs) for this trajectory's inputs */
    double atmp[A1(NASS)];
    rules_all(a.tpoints[0], val, s.pv, atmp);
  }

  double t = a.tpoints[0];
  int status = 0, valid = 0, iout = 1;
  for (; iout <= a.nout; iout++) {
    const double tout = a.tpoints[iout];
#if NEQ > 0
    if (!valid) { s.reinit(t, val); valid = 1; }
#if USE_TSTOP
    s.tstop = tout;
#endif
    const int flag = s.solve(tout, s.y, t, a.mxstep);
    if (flag < 0) { status = flag; break; }
    FORI val[i] = s.y[i];
#else
    t = tout;
#endif
    unsigned int firedmask = 0u;
#if NEVENTS > 0
    {
      int re = 0;
      firedmask = event_f(t, val, s.pv, trigger, re);
      if (re) valid = 0;
    }
#endif
    store_row(a.out + ((size_t)iout * NSTORE) * stride + set, stride, t, val, s.pv);
    if (a.fired) a.fired[(size_t)iout * stride + set] = firedmask;
#if HALT_ON_EVENT
    if (firedmask) { iout++; break; }
#endif
  }
  /* rows never reached (solver failure or halt on event) are marked not-a-number */
  for (; iout <= a.nout; iout++) {
    _Pragma("unroll") for (int k = 0; k < NSTORE; k++) a.out[((size_t)iout * NSTORE + k) * stride + set] = __longlong_as_double(0x7ff8000000000000LL);
    if (a.fired) a.fired[(size_t)iout * stride + set] = 0u;
  }
  if (a.status) a.status[set] = status;
  if (a.stats) {
    a.stats[0 * stride + set] = s.c_nst; a.stats[1 * stride + set] = s.c_nfe; a.stats[2 * stride + set] = s.c_nsetups;
    a.stats[3 * stride + set] = s.c_nje; a.stats[4 * stride + set] = s.c_nni; a.stats[5 * stride + set] = s.c_ncfn;
    a.stats[6 * stride + set] = s.c_netf;
  }
}

#ifndef ODES_HOST_EMULATION
extern "C" __global__ void __launch_bounds__(ODES_BLOCK, ODES_MINBLOCKS) odes_bdf_kernel(const OdesArgs a)
{
  extern __shared__ double odes_smem[];
  const size_t set = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (set >= (size_t)a.nsets) return;
  integrate_set(a, set, odes_smem + threadIdx.x, blockDim.x);
}
#endif
