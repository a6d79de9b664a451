/*
 * bdf_kernel.cuh -- batched variable-order (1..5) variable-step BDF integrator,
 * one trajectory per lane, written for sm_100a (B200).
 *
 * This text is concatenated after the model section produced by
 * ODEModel_generateCUDASource (emit.c) and compiled at run time with NVRTC
 * (-arch=sm_100a).  The model section supplies NEQ/NVAL/... and the device
 * functions model_init, ode_f, jacobi_f, event_f, store_row, store_row0.
 *
 * What it replaces in the reference: the per-design-point call chain
 *   IntegratorInstance_cvodeOneStep -> CVode -> CVStep -> CVnlsNewton ->
 *   CVDenseSetup/CVDenseSolve -> gefa/gesl, plus IntegratorInstance_updateData
 *   (src/cvodeSolver.c:81-240, src/integratorInstance.c:1029-1237, vendored
 *   CVODE 2.3.0 cvode.c:875-2713, cvdense.c:368-441, smalldense.c:58-160).
 * The controller (error test, step/order selection, Newton convergence test,
 * Jacobian/LU reuse rules, initial step) uses CVODE's constants, decision order
 * and floating-point operation order (no FMA contraction, IEEE division and
 * square root), so each trajectory takes the same steps as the reference's
 * CVODE integration of the same parameter set.
 *
 * Data layout.  Inputs, flags and counters are structure-of-arrays with the batch index fastest --
 *   vary[j*stride + set], status[set], stats[k*stride + set], fired[row*stride + set].
 * The trajectories are stored trajectory-major, out[(set*(nout+1) + row)*NSTORE + k]: lanes pull set indices
 * dynamically, so the lanes of a warp do not hold neighbouring sets and a batch-fastest row would be written 8
 * bytes at a time into 32 different sectors (measured: 5.5x the algorithmic DRAM traffic); a lane's row is NSTORE
 * contiguous doubles instead, i.e. whole sectors, and it is the layout the reference keeps per design point.
 * On chip every per-trajectory vector lives in SHARED MEMORY, element-major / lane-minor
 * (word k of lane t at smem[k*ODES_BLOCK + t], bank-conflict free for any run-time k):
 * the Nordsieck array zn[0..5], error weights, acor, y, the f / linear-system vector, the
 * Newton matrix M = I - gamma*J (LU in place), the saved sparse Jacobian, the per-trajectory
 * constants and the BDF coefficient arrays.  Registers hold only scalars.  That makes every
 * vector operation a short ROLLED loop over the state dimension with run-time indexing, so
 * the whole solver is a few thousand instructions and stays in the instruction cache: the
 * fully unrolled register-resident variant (v1-v4 of this file) was 15-20k instructions and
 * spent 40-70 % of its issue slots waiting for instruction fetch (profiles/).
 *
 * Warp scheduling.  A trajectory is a little state machine (ST_* below) and the kernel body is ONE
 * warp-uniform loop whose pass runs the blocks REFILL, START, PRE, F, SETUP, NEWTON, COMPLETE,
 * OUTPUT in this fixed order, each under a per-lane predicate, with the warp reconverged between
 * blocks.  A lane enters the pass at the block its state names and flows through as many
 * consecutive blocks as it can, so lanes stay in phase although every trajectory has its own step
 * size, order, Newton iteration count and retry history.  Expensive rare blocks (linear-solver
 * setup, refill) run when enough lanes want them or one has waited long enough.  Lanes are
 * persistent: a finished lane pulls the next set index from a global counter.  When a lane
 * executes a block never changes its results.
 *
 * Build knobs (defined by the host before this text):
 *   ODES_BLOCK        threads per block
 *   ODES_MINBLOCKS    resident blocks per SM for __launch_bounds__
 *   ODES_MAT_LOCAL    1: keep the per-trajectory vectors in local memory instead of shared memory
 *   ODES_COL_UNROLL   unroll factor of the loops over the state dimension
 *   ODES_SETUP_MIN / ODES_SETUP_WAIT     vote thresholds of the SETUP block
 *   ODES_REFILL_MIN / ODES_REFILL_WAIT   vote thresholds of the REFILL block
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
#ifndef ODES_COL_UNROLL
#define ODES_COL_UNROLL 2
#endif
#ifndef ODES_SETUP_MIN
#define ODES_SETUP_MIN 10           /* measured on MAPK (B200, left-looking LU): 4 245 ms, 6 233, 8 227, 10 224, 12 225, 16 229 */
#endif
#ifndef ODES_SETUP_WAIT
#define ODES_SETUP_WAIT 3
#endif
#ifndef ODES_REFILL_MIN
#define ODES_REFILL_MIN 3
#endif
#ifndef ODES_REFILL_WAIT
#define ODES_REFILL_WAIT 20
#endif
#ifndef ODES_LOCKSTEP
#define ODES_LOCKSTEP 0
#endif
/* cvodeSettings_t.StepResolution ("number of internal integration steps between output steps ... useful for fine-grained
   event detection", reference integratorSettings.h:61-63; the reference never reads the field): the grid the solver is
   driven over -- and on which events are examined, assignments applied and the solver re-initialised, exactly as at the
   reference's output times -- has ODES_STEPRES points per print step; a.nout counts those, rows are KEPT at the print
   steps only, and the event mask of a kept row collects the events fired since the row before. */
#ifndef ODES_STEPRES
#define ODES_STEPRES 1
#endif
/* StepResolution < 0: events are LOCATED inside the internal steps (the procedure is described at the
   warp kernel's integrate_warp and mirrored by oracle/driver.c: run_rootfind) */
#ifndef ODES_ROOTFIND
#define ODES_ROOTFIND 0
#endif
#define ODES_ROOTS (ODES_ROOTFIND && NEVENTS > 0 && NEQ > 0)
#define ODES_ROWS(a) ((size_t)((a).nout / ODES_STEPRES + 1))
#define ODES_ROW(iout) ((size_t)((iout) / ODES_STEPRES))
#define ODES_KEEP(iout) (ODES_STEPRES == 1 || (iout) % ODES_STEPRES == 0)
#ifndef ODES_GESL_UNROLL
#define ODES_GESL_UNROLL 4            /* measured on MAPK (B200): 8 (full) 4.07 M, 4 4.17 M trajectories/s */
#endif
/* ODES_LOCKSTEP 2 re-aligns the warps of a block before every block of the pass as well */
#ifndef ODES_ALIGN_MASK
#define ODES_ALIGN_MASK 7          /* which block boundaries re-align: 1 PRE, 2 F, 4 NEWTON, 8 COMPLETE */
#endif
#if ODES_LOCKSTEP >= 2 && !defined(ODES_HOST_EMULATION)
#define ODES_BLOCK_ALIGN(bit) do { if (ODES_ALIGN_MASK & (bit)) __syncthreads(); } while (0)
#else
#define ODES_BLOCK_ALIGN(bit) ((void)0)
#endif

/* warp-level primitives; the host emulation of the per-trajectory program has no warp */
#ifdef ODES_HOST_EMULATION
#define ODES_WARP_SYNC() ((void)0)
#define ODES_WARP_ALL(p) (p)
#define ODES_WARP_ANY(p) (p)
#define ODES_WARP_COUNT(p) ((p) ? 1 : 0)
static inline unsigned long long odes_take_sets(unsigned long long *counter, bool want, int n) { (void)n; if (!want) return 0; return (*counter)++; }
#define ODES_NOINLINE static
#else
#define ODES_WARP_SYNC() __syncwarp()
#define ODES_WARP_ALL(p) __all_sync(0xffffffffu, (p))
#define ODES_WARP_ANY(p) __any_sync(0xffffffffu, (p))
#define ODES_WARP_COUNT(p) __popc(__ballot_sync(0xffffffffu, (p)))
/* warp-aggregated hand-out of n consecutive set indices to the lanes with want == true */
__device__ __forceinline__ unsigned long long odes_take_sets(unsigned long long *counter, bool want, int n)
{
  const unsigned int mask = __ballot_sync(0xffffffffu, want);
  const unsigned int lane = threadIdx.x & 31u;
  const int leader = __ffs(mask) - 1;
  unsigned long long base = 0;
  if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)n);
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + (unsigned long long)__popc(mask & ((1u << lane) - 1u));
}
#define ODES_NOINLINE __device__ __noinline__
#endif

/* lane states: the block a trajectory needs next */
#define ST_START 0      /* (re)start of the solver at t from y[]: CVodeReInit + first-call block */
#define ST_PRE 1        /* prologue of a step attempt: predict, coefficients */
#define ST_F 2          /* right-hand side at (tn, y) */
#define ST_SETUP 3      /* Newton matrix: Jacobian (maybe), M = I - gamma J, LU */
#define ST_NEWTON 4     /* one Newton iteration: solve, norms, convergence test, error test */
#define ST_COMPLETE 5   /* accepted step: history update, next step size and order */
#define ST_OUTPUT 6     /* output rows covered by the last step */
#define ST_DONE 7       /* no trajectory (idle lane) */
#define ST_FINISH 8     /* trajectory ended (grid done, failure, halt on event): write status, counters, NaN rows */
#define ST_ROOT 9       /* event location (ODES_ROOTFIND): triggers at the end of the step just completed, bisection on the dense output */
/* why the right-hand side is being evaluated */
#define FW_ATTEMPT 0    /* f(tn, predicted z0) at the start of the nonlinear solve */
#define FW_NEWTON 1     /* f(tn, y) inside the Newton iteration */
#define FW_RESTART 2    /* f(tn, zn[0]) to reload zn[1] at order 1 after repeated error-test failures */

#ifndef ODES_ADAMS
#define ODES_ADAMS 0          /* 1: Adams-Moulton (orders 1..12) instead of BDF (1..5); warp-per-trajectory kernel only */
#endif
#ifndef ODES_FUNCTIONAL
#define ODES_FUNCTIONAL 0      /* 1: functional (fixed-point) iteration instead of Newton; warp-per-trajectory kernel only */
#endif
#if ODES_ADAMS
#define QMAX 12
#define LMAX 13
#else
#define QMAX 5
#define LMAX 6
#endif
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

#ifndef NSENS
#define NSENS 0
#endif
struct OdesArgs {
  const double *vary;          /* [NVARY][stride]                       */
  double *out;                 /* [nsets][nout+1][NSTORE]               */
  int *status;                 /* [stride]                              */
  int *stats;                  /* [7][stride]  nst,nfe,nsetups,nje,nni,ncfn,netf */
  unsigned int *fired;         /* [nout+1][stride] event bitmasks       */
  const double *tpoints;       /* [nout+1] output grid                  */
  unsigned long long *counter; /* next set index to hand out (zeroed before the launch) */
  double *scratch;             /* [ODES_GLOBAL_WORDS][lanes of the grid] (ODES_GLOBAL_LA) */
  unsigned long long stride;
  int nsets, nout, mxstep, pad;       /* pad != 0: events are not processed (IntegratorInstance_integrateOneStepWithoutEventProcessing) */
  double reltol, abstol;
  /* completion signalling for the host pipeline (ODESBatch_runHost): sets are handed out in index order, so groups of
     2^groupShift consecutive sets finish roughly in order; the lane that completes a group publishes it to a flag array
     in mapped host memory and the host copies that group's results while the kernel is still integrating the rest */
  unsigned int *groupDone;     /* [groups] device counters, zeroed before the launch; NULL: no signalling */
  unsigned int *hostFlags;     /* [groups] mapped pinned host memory */
  int groupShift, pad2;        /* pad2 != 0: the initial sensitivities are read from row 0 of sens */
  double *sens;                /* [nsets][nout+1][NSENS][NEQ] forward sensitivities (warp kernel with NSENS > 0), or NULL */
};
#ifndef ODES_HOST_EMULATION
__device__ __forceinline__ void odes_signal_done(const OdesArgs &a, size_t set)
{
  if (!a.groupDone) return;
  __threadfence();                                            /* this trajectory's rows, status and counters first */
  const unsigned int g = (unsigned int)(set >> a.groupShift);
  const unsigned int first = g << a.groupShift, left = (unsigned int)a.nsets - first, full = 1u << a.groupShift;
  const unsigned int size = left < full ? left : full;
  if (atomicAdd(&a.groupDone[g], 1u) + 1u == size) {
    __threadfence_system();
    *((volatile unsigned int *)a.hostFlags + g) = 1u;
  }
}
#else
static inline void odes_signal_done(const OdesArgs &, size_t) {}
#endif

#define ODES_STR_(x) #x
#define ODES_STR(x) ODES_STR_(x)
/* loop over the state dimension (columns of the Nordsieck array) */
#define FORC _Pragma(ODES_STR(unroll ODES_COL_UNROLL)) for (int i = 0; i < NEQ; i++)
/* the same loop in the blocks every step runs (predictor, Newton update, step completion): these are chains of dependent
   FP64 operations per column (6-9 cycles each, profiles/r02_ncu_v14), so more columns in flight hide more latency */
#ifndef ODES_HOT_UNROLL
#define ODES_HOT_UNROLL ODES_COL_UNROLL
#endif
#define FORH _Pragma(ODES_STR(unroll ODES_HOT_UNROLL)) for (int i = 0; i < NEQ; i++)
/* fully unrolled loop, only where model code needs compile-time indices */
#define FORI _Pragma("unroll") for (int i = 0; i < NEQ; i++)
#define NMAT (NEQ * NEQ)
#if HAS_JAC
#define NSJ NNZ
#else
#define NSJ NMAT
#endif
/* Per-lane storage.
   SHARED MEMORY (word k of lane t at smem[k*ODES_BLOCK + t], conflict-free for any run-time k): the Nordsieck
   array, error weights, acor, y, the f / linear-system vector and the BDF coefficient arrays -- everything a step
   touches element by element, many times.
   LINEAR-ALGEBRA WORDS: the Newton matrix M = I - gamma*J (LU in place, column-major, columns padded to NP rows),
   the saved Jacobian and the per-trajectory constants.  They are only ever moved in chunks of 8 doubles
   (la_load8 / la_store8: a matrix column, 8 Jacobian entries, 8 constants) at warp-converged points, which lets
   them live in one of three places (ODES_LA):
     2  TENSOR MEMORY (Blackwell TMEM, 256 KB per SM next to the 228 KB of shared memory): a warp owns 32 TMEM
        lanes, lane t of the warp holds trajectory t's words in its columns, and one tcgen05.ld/st 32x32b.x16 moves
        8 doubles per lane (measured here: ~32 cycles latency for a dependent load, tools/micro/tmem_probe.cu).
        Halves the shared memory per lane -> twice the resident warps, and a matrix column costs one instruction
        instead of eight shared-memory loads.
     0  shared memory (models whose words do not fit the 512 TMEM columns)
     1  a global scratch that stays L2-resident (largest models). */
#ifndef ODES_LA
#define ODES_LA 0
#endif
#ifdef ODES_HOST_EMULATION
#undef ODES_LA
#define ODES_LA 0
#endif
/* Form of the dense linear algebra (gefa / gesl):
     0  element-wise on shared-memory words (run-time row indices are plain addresses)
     1  right-looking, one matrix column at a time in registers, rolled loops (run-time pivot rows picked with selects)
     2  LEFT-LOOKING with compile-time indices: column j is built from the saved Jacobian in registers, receives the
        eliminations of the columns before it (each skipped when its multiplier row entry is zero in every lane of the
        warp -- the structural zeros of a reaction network), is searched for its pivot, scaled and stored ONCE; row
        exchanges are applied to the stored multipliers as well (LAPACK's layout instead of LINPACK's), so that the
        solves permute the right-hand side up front and then run without any run-time row index.  Every matrix
        element still receives the reference's operations in the reference's order (smalldense.c:58-160).
   Form 2 needs the analytic Jacobian (the column build is fused into it) and is meant for small systems. */
#ifndef ODES_LU2_MAXN
#define ODES_LU2_MAXN 10
#endif
#ifndef ODES_LU_FORM
#if ODES_LA == 2 && HAS_JAC && NEQ >= 2 && NEQ <= ODES_LU2_MAXN
#define ODES_LU_FORM 2
#elif ODES_LA != 0
#define ODES_LU_FORM 1
#else
#define ODES_LU_FORM 0
#endif
#endif
#if ODES_LU_FORM == 2 && !(HAS_JAC && NEQ >= 2 && NEQ <= 32)
#undef ODES_LU_FORM
#define ODES_LU_FORM 0
#endif
#if defined(ODES_HOST_EMULATION) && defined(ODES_EMU_COUNT_SWAPS)
extern "C" { long odes_emu_swaps = 0; }
#define ODES_EMU_NOTE_SWAP (odes_emu_swaps++)
#else
#define ODES_EMU_NOTE_SWAP ((void)0)
#endif
#define ZOFF 0                                    /* zn[0..5][NEQ]                         */
#define EWOFF (ZOFF + LMAX * NEQ)                 /* error weights                         */
#define ACOFF (EWOFF + NEQ)                       /* acor: accumulated Newton correction   */
#define YOFF (ACOFF + NEQ)                        /* y: argument of f; data->value[0..neq) between steps */
#ifndef ODES_ALIAS_FY
#define ODES_ALIAS_FY 0
#endif
/* f(t,y); right-hand side / solution of the linear system; scratch.  With ODES_ALIAS_FY (set by the host where it buys a
   resident block, analytic Jacobian only) f SHARES the words of y: the
   argument is dead once f has been evaluated (f_conv loads all of y before it stores f; the Newton update reads the solved
   correction of a column before it writes the column's new y; every other writer of y -- predictor, restarts, the initial-step
   search, the dense output -- comes when the last f is no longer needed).  Only the difference-quotient Jacobian perturbs y
   while it still holds f(y).  One vector less per lane: a six-equation model fits three blocks per SM. */
#if HAS_JAC && ODES_ALIAS_FY
#define FTOFF YOFF
#else
#define FTOFF (YOFF + NEQ)
#endif
#define TAUOFF (FTOFF + NEQ)                      /* tau[0..6], tq[0..5], l[0..5]          */
#define LAOFF (TAUOFF + (LMAX + 1) + 6 + LMAX)    /* linear-algebra words when they are in shared memory */
/* linear-algebra word space (doubles); NP, NSJP, NPVP (multiples of 8) come from the model section */
#define LAM 0                                     /* M: column j at [j*NP, j*NP + NP)      */
#define LAJ (NEQ * NP)                            /* saved Jacobian                        */
#define LAPV (LAJ + NSJP)                         /* per-trajectory constants              */
#define ODES_LA_WORDS (LAPV + NPVP)
#define NCH (NP / 8)
#define NSJCH (NSJP / 8)
#define NPVCH (NPVP / 8)
/* the buffer the per-trajectory constants are moved in: their own padded region, or -- packed (ODES_LA_PACKED, small models
   whose words then fit half the tensor-memory columns: a third resident block per SM) -- the chunks of the saved Jacobian,
   behind its NNZ entries */
#if ODES_LA_PACKED && HAS_JAC
#define PVB_WORDS NSJP
#define PVB_CH NSJCH
#define PVB_BASE LAJ
#define PVB_OFF NNZ
#else
#define PVB_WORDS NPVP
#define PVB_CH NPVCH
#define PVB_BASE LAPV
#define PVB_OFF 0
#endif
#if ODES_LA == 0
#define ODES_SMEM_WORDS (LAOFF + ODES_LA_WORDS)
#else
#define ODES_SMEM_WORDS LAOFF
#endif
#define ODES_GLOBAL_WORDS ODES_LA_WORDS

#define V(off, i) sm[((off) + (i)) * ODES_SSTRIDE]
/* inside a column loop: c points at word i of this lane, so every vector element of column i is at a
   compile-time offset from it (one address computation per iteration instead of one per access) */
#define COLPTR double *const c = sm + i * ODES_SSTRIDE
#define C(off) c[(off) * ODES_SSTRIDE]
#define CZ(j) C(ZOFF + (j) * NEQ)
#define ZN(j, i) V(ZOFF + (j) * NEQ, i)
#define EW(i) V(EWOFF, i)
#define ACOR(i) V(ACOFF, i)
#define YV(i) V(YOFF, i)
#define FT(i) V(FTOFF, i)
#define TAU(j) V(TAUOFF, j)
#define TQ(j) V(TAUOFF + (LMAX + 1), j)
#define LL(j) V(TAUOFF + (LMAX + 1) + 6, j)

#ifndef ODES_WARP
#define ODES_WARP 0
#endif
#if !ODES_WARP
/* where this lane's linear-algebra words are */
struct La {
  double *sm;                  /* ODES_LA 0: this lane's shared-memory words                     */
  double *gm; unsigned int gs; /* ODES_LA 1: this lane's global scratch words and their stride   */
  unsigned int taddr;          /* ODES_LA 2: TMEM address (lane base of the warp, first column)  */
};
#if ODES_LA == 2
DEV void la_load8(const La &la, int w, double (&v)[8])
{
  unsigned int r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(la.taddr + 2u * (unsigned int)w));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  _Pragma("unroll") for (int k = 0; k < 8; k++) v[k] = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
}
DEV void la_store8(const La &la, int w, const double (&v)[8])
{
  unsigned int r[16];
  _Pragma("unroll") for (int k = 0; k < 8; k++) { r[2 * k] = (unsigned int)__double2loint(v[k]); r[2 * k + 1] = (unsigned int)__double2hiint(v[k]); }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
               :: "r"(la.taddr + 2u * (unsigned int)w), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                  "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
/* stores must have landed before the same words are loaded again */
DEV void la_fence(const La &) { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
#elif ODES_LA == 1
DEV void la_load8(const La &la, int w, double (&v)[8]) { _Pragma("unroll") for (int k = 0; k < 8; k++) v[k] = la.gm[(size_t)(w + k) * la.gs]; }
DEV void la_store8(const La &la, int w, const double (&v)[8]) { _Pragma("unroll") for (int k = 0; k < 8; k++) la.gm[(size_t)(w + k) * la.gs] = v[k]; }
DEV void la_fence(const La &) {}
#else
DEV void la_load8(const La &la, int w, double (&v)[8]) { const double *p = la.sm + (LAOFF + w) * ODES_SSTRIDE; _Pragma("unroll") for (int k = 0; k < 8; k++) v[k] = p[k * ODES_SSTRIDE]; }
DEV void la_store8(const La &la, int w, const double (&v)[8]) { double *p = la.sm + (LAOFF + w) * ODES_SSTRIDE; _Pragma("unroll") for (int k = 0; k < 8; k++) p[k * ODES_SSTRIDE] = v[k]; }
DEV void la_fence(const La &) {}
#endif
/* vectors of N8 chunks */
template <int N8> DEV void la_load(const La &la, int w, double (&v)[N8 * 8 > 0 ? N8 * 8 : 1])
{
  _Pragma("unroll") for (int c = 0; c < N8; c++) { double t[8]; la_load8(la, w + 8 * c, t); _Pragma("unroll") for (int k = 0; k < 8; k++) v[8 * c + k] = t[k]; }
}
template <int N8> DEV void la_store(const La &la, int w, const double (&v)[N8 * 8 > 0 ? N8 * 8 : 1])
{
  _Pragma("unroll") for (int c = 0; c < N8; c++) { double t[8]; _Pragma("unroll") for (int k = 0; k < 8; k++) t[k] = v[8 * c + k]; la_store8(la, w + 8 * c, t); }
}
/* commits a lane's new values where take is set and keeps the stored ones elsewhere (a TMEM store is warp-wide) */
template <int N8> DEV void la_commit(const La &la, int w, const double (&v)[N8 * 8 > 0 ? N8 * 8 : 1], bool take)
{
  _Pragma("unroll") for (int c = 0; c < N8; c++) {
    double t[8]; la_load8(la, w + 8 * c, t);
    _Pragma("unroll") for (int k = 0; k < 8; k++) if (take) t[k] = v[8 * c + k];
    la_store8(la, w + 8 * c, t);
  }
  la_fence(la);
}

#endif /* !ODES_WARP */

/* 1.0/j for the small integers of the BDF coefficient formulas (j <= 7): compile-time constants are the
   correctly rounded quotients, i.e. what the reference's run-time division ONE/j yields */
#ifdef ODES_HOST_EMULATION
static const double c_inv_int[8] = {0.0, 1.0, 0.5, 1.0 / 3.0, 0.25, 0.2, 1.0 / 6.0, 1.0 / 7.0};
#else
__constant__ double c_inv_int[8] = {0.0, 1.0, 0.5, 1.0 / 3.0, 0.25, 0.2, 1.0 / 6.0, 1.0 / 7.0};
#endif
DEV double inv_int(int j) { return c_inv_int[j]; }

/* RPowerR(base, 1/n) of sundialsmath.c, i.e. pow(base, 1.0/n) with n = 1..7 (step-size controller,
   cvode.c:2497, :2612-2676), 0 for base <= 0.  The controller's decisions hang on these values: with CUDA's pow
   (<= 2 ulp, different from the host libm in ~30 % of the calls) 1.5-2.3 % of MAPK trajectories took a different step
   sequence than the reference; with a correctly rounded root (glibc misrounds ~8e-4 of the calls) it was still
   1.4e-4 of MAPK and 2 % of the much more sensitive huang96 trajectories.  So the kernel evaluates the function the
   reference evaluates: glibc's pow, restated in glibc_pow.inc (odes_glibc_pow), bit-identical to libm. */
ODES_NOINLINE double root_pow(double x, int n)
{
  if (x <= 0.0) return 0.0;
  if (!(x > 1e-300 && x < 1e300)) return pow(x, 1.0 / n);
  return odes_glibc_pow(x, n <= 7 ? inv_int(n) : 1.0 / n);     /* n up to 13 with Adams-Moulton (orders 1..12) */
}
ODES_NOINLINE double rsqrt_(double x) { return x <= 0.0 ? 0.0 : sqrt(x); }
/* The step-size factors are eta = 1 / (RPowerR(x, 1/n) + ADDON) and are only looked at above THRESH = 1.5 (CVSetEta,
   cvode.c:2689-2700; the order selection compares the largest of three, :2640-2660).  For x >= c_eta_small[n] =
   (2/3 + 1e-6)^n (1 + 1e-9) the root is above 2/3 + 1e-6 whatever the last bit of pow, so eta < 1.5 for certain and the
   pow call is skipped -- the decisions (eta = 1, or "not the largest") are the reference's. */
#ifdef ODES_HOST_EMULATION
static const double c_eta_small[8] =
#else
__constant__ double c_eta_small[8] =
#endif
  {1e300, 1e300, 0.44444577822322356, 0.2962976299279273, 0.1975320495829148, 0.1316882305873261, 0.08779228541311465, 0.05852827806769518};
DEV bool eta_below_thresh(double x, int n) { return n >= 2 && n <= 7 && x >= c_eta_small[n]; }
#if !ODES_WARP
#ifndef ODES_POW_SHARE
#define ODES_POW_SHARE 0      /* measured on MAPK (B200, 10^6 sets) once the certain-below-THRESH requests are skipped: shared 224.6 ms, per lane 222.8 ms */
#endif
/* RPowerR for up to three (x, n) requests per lane, warp-converged.  pow is a pure function, so it does not matter WHICH
   lane evaluates a request: the requests of the warp (about one per completing lane plus two for the few lanes that
   select a new order) are compacted into a per-warp staging array and evaluated by all 32 lanes together -- one pass
   through the 300 instructions of pow per warp instead of three passes with 20, 9 and 3 lanes. */
#if defined(ODES_HOST_EMULATION) || !ODES_POW_SHARE
DEV void pow_service(const int nreq, double (&x)[3], const int (&n)[3])
{
  _Pragma("unroll") for (int r = 0; r < 3; r++) if (r < nreq) x[r] = root_pow(x[r], n[r]);
}
#else
DEV void pow_service(const int nreq, double (&x)[3], const int (&n)[3])
{
  __shared__ double stage_x[ODES_BLOCK / 32][96];
  __shared__ int stage_n[ODES_BLOCK / 32][96];
  const unsigned int lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  const unsigned int b0 = __ballot_sync(0xffffffffu, (nreq & 1) != 0), b1 = __ballot_sync(0xffffffffu, (nreq & 2) != 0);
  const unsigned int below = (1u << lane) - 1u;
  const int off = __popc(b0 & below) + 2 * __popc(b1 & below), total = __popc(b0) + 2 * __popc(b1);
  if (total == 0) return;
  _Pragma("unroll") for (int r = 0; r < 3; r++) if (r < nreq) { stage_x[w][off + r] = x[r]; stage_n[w][off + r] = n[r]; }
  __syncwarp();
  _Pragma("unroll 1") for (int base = 0; base < total; base += 32) {
    const int idx = base + (int)lane;
    if (idx < total) stage_x[w][idx] = root_pow(stage_x[w][idx], stage_n[w][idx]);
  }
  __syncwarp();
  _Pragma("unroll") for (int r = 0; r < 3; r++) if (r < nreq) x[r] = stage_x[w][off + r];
  __syncwarp();
}
#endif
/* N_VWrmsNorm(x, ewt) with x the vector at word offset off (nvector_serial.c:591) */
ODES_NOINLINE double wrms_vec(const double *sm, int off)
{
  double sum = 0.0;
  const double *x = sm + off * ODES_SSTRIDE;
  FORC { const double p = x[i * ODES_SSTRIDE] * EW(i); sum += p * p; }
  return rsqrt_(sum / NEQ);
}

/* The one instance of the model's right-hand side: FT <- f(t, YV) for the lanes with act set.  Called with the
   warp converged (the constants come through la_load, which is warp-wide); lanes without act evaluate the
   expressions on whatever their y holds and discard the result.  A real function: all callers (hot F block,
   first call, initial-step search, difference-quotient Jacobian) share these instructions. */
ODES_NOINLINE void f_conv(double *sm, La la, double t, bool act)
{
  double yl[A1(NEQ)], fl[A1(NEQ)], pvl[A1(PVB_WORDS)];
  la_load<PVB_CH>(la, PVB_BASE, pvl);
  PvRef pv; pv.p = pvl + PVB_OFF; pv.s = 1u;
  FORI yl[i] = YV(i);
  ode_f(t, yl, pv, fl);
  if (act) { FORI FT(i) = fl[i]; }
}

struct Solver {
  double *sm;                  /* this lane's shared-memory words */
  La la;                       /* this lane's linear-algebra words */
#if ODES_MAT_LOCAL
  double mat[ODES_SMEM_WORDS];
#endif
  double h, hprime, eta, hscale, tn, rl1, gamma, gammap, gamrat, crate, acnrm, etamax, hu, saved_tq5;
  double reltol, abstol, tstop;
  int q, qprime, qwait, L, jcur;
  int nst, nstlp, nstlj;
  /* state of the step machine: first attempt of a step, steps in this CVode call, CVStep's failure
     counters and nflag, Newton iteration counter, block state, pending setup request */
  int fresh, nstloc, ncf, nef, nflag, nl_m;
  int st, fwhy, callSetup, convfail, waited;
  int pend, ydone;             /* pending history transforms fused into the next predictor (1 restore, 2 rescale); y already interpolated */
  double saved_t, nl_delp, dsm, resc_eta;
  unsigned int swapmask;       /* LU form 2: bit k set when step k of the factorisation exchanged rows */
  int c_nst, c_nfe, c_nsetups, c_nje, c_nni, c_ncfn, c_netf;      /* totals over re-inits */
#if NEQ <= 16
  unsigned long long pivpack;                                      /* LU pivot rows, 4 bits each */
  DEV void piv_clear() { pivpack = 0ull; }
  DEV void piv_set(int k, int p) { pivpack |= (unsigned long long)p << (4 * k); }
  DEV int piv_get(int k) const { return (int)((pivpack >> (4 * k)) & 15ull); }
#else
  int pivarr[A1(NEQ)];
  DEV void piv_clear() {}
  DEV void piv_set(int k, int p) { pivarr[k] = p; }
  DEV int piv_get(int k) const { return pivarr[k]; }
#endif

  /* converged: right-hand side for the lanes with act */
  DEV void f(double t, bool act) { f_conv(sm, la, t, act); if (act) c_nfe++; }

  /* N_VWrmsNorm(x, ewt) with x a vector at word offset off (nvector_serial.c:591) */
  DEV double wrms(int off) const { return wrms_vec(sm, off); }
  /* CVEwtSetSV on zn[0] (cvode.c:3542-3549); false when a weight would be non-positive */
  DEV bool ewt_set()
  {
    double mn = 1.0;
    FORH { COLPTR; const double w = reltol * fabs(CZ(0)) + abstol; if (i == 0 || w < mn) mn = w; C(EWOFF) = 1.0 / w; }
    return mn > 0.0;
  }

  DEV void clear_history()
  {
    _Pragma("unroll 1") for (int k = 0; k < LAOFF; k++) V(0, k) = 0.0;
  }

  /* CVRescale (cvode.c:1890-1902): zn[j] *= eta^j, h = hscale*eta.  The step size changes now; the history rows
     are scaled inside the next predictor's column loop (same values: nothing reads them in between). */
  DEV void rescale()
  {
    flush_pending_rescale();
    resc_eta = eta; pend |= 2;
    h = hscale * eta;
    hscale = h;
  }
  /* CVRestore (cvode.c:2448-2458): undo the predictor.  Deferred like the rescale when the next thing that touches
     the history is the predictor of the retry (the usual case); restore_now() where the rows are needed at once. */
  DEV void restore(double t0) { tn = t0; pend |= 1; }
  /* The history loops below load the rows above the current order as signed zeros instead of predicating every
     operation on q: x + (-0.0) == x and x - (+0.0) == x hold exactly for every x (zeros of either sign included),
     so the unconditional Pascal triangle on a masked column does to rows 0..q exactly what the reference's loops over
     j <= q do, the masked rows stay zero, and only the stores are predicated. */
  DEV void history_column(double *const c, const int qq, const bool pr, const bool ps, const bool pp,
                          const double f1, const double f2, const double f3, const double f4, const double f5)
  {
    const bool p2 = qq >= 2, p3 = qq >= 3, p4 = qq >= 4, p5 = qq >= 5;
    double z0 = CZ(0), z1 = CZ(1), z2 = CZ(2), z3 = CZ(3), z4 = CZ(4), z5 = CZ(5);
    if (pr) {                                                 /* CVRestore: z[j-1] -= z[j] */
      z2 = p2 ? z2 : 0.0; z3 = p3 ? z3 : 0.0; z4 = p4 ? z4 : 0.0; z5 = p5 ? z5 : 0.0;
      _Pragma("unroll") for (int k = 1; k <= QMAX; k++) {
        if (k <= 5) z4 = z4 - z5;
        if (k <= 4) z3 = z3 - z4;
        if (k <= 3) z2 = z2 - z3;
        if (k <= 2) z1 = z1 - z2;
        if (k <= 1) z0 = z0 - z1;
      }
    }
    if (ps) { z1 = f1 * z1; z2 = f2 * z2; z3 = f3 * z3; z4 = f4 * z4; z5 = f5 * z5; }     /* CVRescale (rows above q: garbage or zero, not stored) */
    if (pp) {                                                 /* CVPredict: z[j-1] += z[j] */
      z2 = p2 ? z2 : -0.0; z3 = p3 ? z3 : -0.0; z4 = p4 ? z4 : -0.0; z5 = p5 ? z5 : -0.0;
      _Pragma("unroll") for (int k = 1; k <= QMAX; k++) {
        if (k <= 5) z4 = z4 + z5;
        if (k <= 4) z3 = z3 + z4;
        if (k <= 3) z2 = z2 + z3;
        if (k <= 2) z1 = z1 + z2;
        if (k <= 1) z0 = z0 + z1;
      }
    }
    CZ(0) = z0; CZ(1) = z1;
    if (p2) CZ(2) = z2;
    if (p3) CZ(3) = z3;
    if (p4) CZ(4) = z4;
    if (p5 && ps) CZ(5) = z5;
  }
  DEV void apply_pending()
  {
    if (!pend) return;
    const int qq = q;
    const bool pr = (pend & 1) != 0, ps = (pend & 2) != 0;
    const double f1 = resc_eta, f2 = f1 * resc_eta, f3 = f2 * resc_eta, f4 = f3 * resc_eta, f5 = f4 * resc_eta;
    FORC { COLPTR; history_column(c, qq, pr, ps, false, f1, f2, f3, f4, f5); }
    pend = 0;
  }
  DEV void restore_now(double t0) { restore(t0); apply_pending(); }
  DEV void flush_pending_rescale() { if (pend & 2) apply_pending(); }
  /* CVPredict (cvode.c:1912-1920): Pascal triangle on the columns of zn, preceded by the pending restore / rescale
     of the same columns */
  DEV void predict()
  {
    tn += h;
    const int qq = q;
    const bool pr = (pend & 1) != 0, ps = (pend & 2) != 0;
    const double f1 = resc_eta, f2 = f1 * resc_eta, f3 = f2 * resc_eta, f4 = f3 * resc_eta, f5 = f4 * resc_eta;
    FORH { COLPTR; history_column(c, qq, pr, ps, true, f1, f2, f3, f4, f5); }
    pend = 0;
  }

  /* CVIncreaseBDF (cvode.c:1799-1838) */
  DEV void increase_bdf()
  {
    double alpha0, alpha1, prod, xi, xiold, hsum, A1c;
    _Pragma("unroll 1") for (int i = 0; i <= QMAX; i++) LL(i) = 0.0;
    LL(2) = alpha1 = prod = xiold = 1.0;
    alpha0 = -1.0;
    hsum = hscale;
    _Pragma("unroll 1") for (int j = 1; j < q; j++) {
      hsum += TAU(j + 1);
      xi = hsum / hscale;
      prod *= xi;
      alpha0 -= inv_int(j + 1);
      alpha1 += 1.0 / xi;
      _Pragma("unroll 1") for (int i = j + 2; i >= 2; i--) LL(i) = LL(i) * xiold + LL(i - 1);
      xiold = xi;
    }
    A1c = (-alpha0 - alpha1) / prod;
    /* zn[L] = A1 * zn[qmax]; zn[j] += l[j] * zn[L], j = 2..q   (L = q+1 <= 5 here) */
    FORC {
      const double zl = A1c * ZN(QMAX, i);
      ZN(L, i) = zl;
      _Pragma("unroll 1") for (int j = 2; j <= q; j++) ZN(j, i) = LL(j) * zl + ZN(j, i);
    }
  }
  /* CVDecreaseBDF (cvode.c:1849-1876) */
  DEV void decrease_bdf()
  {
    double hsum, xi;
    _Pragma("unroll 1") for (int i = 0; i <= QMAX; i++) LL(i) = 0.0;
    LL(2) = 1.0;
    hsum = 0.0;
    _Pragma("unroll 1") for (int j = 1; j <= q - 2; j++) {
      hsum += TAU(j);
      xi = hsum / hscale;
      _Pragma("unroll 1") for (int i = j + 2; i >= 2; i--) LL(i) = LL(i) * xi + LL(i - 1);
    }
    FORC {
      const double zq = ZN(q, i);
      _Pragma("unroll 1") for (int j = 2; j < q; j++) ZN(j, i) = (-LL(j)) * zq + ZN(j, i);
    }
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
  /* CVSetBDF + CVSetTqBDF (cvode.c:2086-2148) */
  DEV void set_bdf()
  {
    double alpha0, alpha0_hat, xi_inv, xistar_inv, hsum, A1c, A2, A3, A4, A5, A6, C, CPrime, CPrimePrime;
    LL(0) = LL(1) = xi_inv = xistar_inv = 1.0;
    _Pragma("unroll 1") for (int i = 2; i <= q; i++) LL(i) = 0.0;
    alpha0 = alpha0_hat = -1.0;
    hsum = h;
    if (q > 1) {
      _Pragma("unroll 1") for (int j = 2; j < q; j++) {
        hsum += TAU(j - 1);
        xi_inv = h / hsum;
        alpha0 -= inv_int(j);
        _Pragma("unroll 1") for (int i = j; i >= 1; i--) LL(i) += LL(i - 1) * xi_inv;
      }
      alpha0 -= inv_int(q);
      xistar_inv = -LL(1) - alpha0;
      hsum += TAU(q - 1);
      xi_inv = h / hsum;
      alpha0_hat = -LL(1) - xi_inv;
      _Pragma("unroll 1") for (int i = q; i >= 1; i--) LL(i) += LL(i - 1) * xistar_inv;
    }
    const double lq = LL(q);
    A1c = 1.0 - alpha0_hat + alpha0;
    A2 = 1.0 + q * A1c;
    TQ(2) = fabs(alpha0 * (A2 / A1c));
    TQ(5) = fabs((A2) / (lq * xi_inv / xistar_inv));
    if (qwait == 1) {
      C = xistar_inv / lq;
      A3 = alpha0 + inv_int(q);
      A4 = alpha0_hat + xi_inv;
      CPrime = A3 / (1.0 - A4 + A3);
      TQ(1) = fabs(CPrime / C);
      hsum += TAU(q);
      xi_inv = h / hsum;
      A5 = alpha0 - inv_int(q + 1);
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

  /* ---- dense linear algebra (gefa / gesl, smalldense.c:58-160), one matrix column (NP doubles) at a time in
     registers; all three routines run with the warp converged, lanes without act keep their data ---- */
  DEV void mcol_load(int j, double (&c)[A1(NP)]) const { la_load<NCH>(la, LAM + j * NP, c); }
  DEV void mcol_store(int j, const double (&c)[A1(NP)]) const { la_store<NCH>(la, LAM + j * NP, c); }

#if ODES_LU_FORM == 2
  /* rows j and lp of the multiplier columns 0..j-1 change places (the exchange of step j applied to the stored L) */
  DEV void swap_stored_rows(const int j, const int lp, const bool doit) const
  {
    _Pragma("unroll 1") for (int k = 0; k < j; k++) {
      double m[A1(NP)];
      mcol_load(k, m);
      if (doit) {
        double mj = 0.0, ml = 0.0;
        FORI { if (i == j) mj = m[i]; if (i == lp) ml = m[i]; }
        FORI { if (i == j) m[i] = ml; if (i == lp) m[i] = mj; }
      }
      mcol_store(k, m);
    }
    la_fence(la);
  }
  /* M = I - gamma*J built column by column from the saved Jacobian and factored on the way (left-looking gefa);
     returns 0, or j+1 when column j has no pivot.  Warp-converged; lanes without act keep their stored matrix. */
  DEV int factor_columns(const bool act0, const double (&jnz)[A1(NSJP)], const double mg)
  {
    bool act = act0;
    int ier = 0;
    unsigned int swm = 0u;
    bool scratched = false;
    if (act) piv_clear();
#ifndef ODES_LU2_ROLLJ
#define ODES_LU2_ROLLJ 1
#endif
#if ODES_LU2_ROLLJ
    /* the column loop stays rolled (j is a run-time value, the eliminations inside are unrolled over k and predicated on
       k < j): a third of the instructions of the fully unrolled form -- the kernel is bound by instruction delivery, and
       measured on MAPK the smaller text wins although it executes more */
    _Pragma("unroll 1") for (int j = 0; j < NEQ; j++) {
#else
    _Pragma("unroll") for (int j = 0; j < NEQ; j++) {
#endif
      double c[A1(NP)];
      {
        const double z = 0.0 * mg;
        _Pragma("unroll") for (int i = 0; i < NP; i++) c[i] = z;
        scatter_col(j, c, jnz, mg);
        FORI if (i == j) c[i] = c[i] + 1.0;
      }
      /* the row exchanges of the steps before (rare: through this lane's acor words, which are zero during a setup) */
      if (j > 0 && ODES_WARP_ANY(act && swm != 0u)) {
        if (act && swm != 0u) {
          FORI ACOR(i) = c[i];
          _Pragma("unroll 1") for (int k = 0; k < j; k++) {
            const int lp = piv_get(k);
            if (lp != k) { const double t = ACOR(lp); ACOR(lp) = ACOR(k); ACOR(k) = t; }
          }
          FORI c[i] = ACOR(i);
          scratched = true;
        }
        ODES_WARP_SYNC();
      }
      /* eliminations of the steps before, each with its stored multipliers */
      _Pragma("unroll") for (int k = 0; k < NEQ - 1; k++) {
        if (k < j) {
          const double a_kj = c[k];
          if (ODES_WARP_ANY(act && a_kj != 0.0)) {
            double m[A1(NP)];
            mcol_load(k, m);
            if (a_kj != 0.0) { _Pragma("unroll") for (int i = 0; i < NEQ; i++) if (i > k) c[i] = c[i] + a_kj * m[i]; }
          }
        }
      }
      if (j < NEQ - 1) {
        /* pivot: the first row of maximal magnitude at or below the diagonal */
        int lp = j;
        double cj = 0.0;
        FORI if (i == j) cj = c[i];
        double pval = cj, big = fabs(cj);
        _Pragma("unroll") for (int i = 0; i < NEQ; i++) if (i > j) { const double v = fabs(c[i]); if (v > big) { big = v; lp = i; pval = c[i]; } }
        if (act) {
          piv_set(j, lp);
          if (pval == 0.0) { ier = j + 1; act = false; }
        }
        const bool sw = act && (lp != j);
        if (ODES_WARP_ANY(sw)) {
          if (sw) {
            ODES_EMU_NOTE_SWAP;
            _Pragma("unroll") for (int i = 0; i < NEQ; i++) { if (i > j && i == lp) c[i] = cj; if (i == j) c[i] = pval; }
            swm |= 1u << j;
          }
          if (j > 0) swap_stored_rows(j, lp, sw);
        }
        const double mult = -1.0 / pval;
        _Pragma("unroll") for (int i = 0; i < NEQ; i++) if (i > j) c[i] = c[i] * mult;
      } else if (act) {
        piv_set(NEQ - 1, NEQ - 1);
        if (c[NEQ - 1] == 0.0) ier = NEQ;
      }
      /* the finished column replaces the stored one in the lanes that factor */
      {
        double o[A1(NP)];
        mcol_load(j, o);
        _Pragma("unroll") for (int i = 0; i < NP; i++) if (act0) o[i] = c[i];
        mcol_store(j, o);
        la_fence(la);
      }
    }
    if (act0) { swapmask = swm; if (scratched) { FORC ACOR(i) = 0.0; } }
    return ier;
  }
  /* solves M x = b in place, b = FT, for the lanes with act: b is permuted up front, then both triangular solves run on
     registers with compile-time indices, one stored column per step */
  DEV void gesl(bool act)
  {
    if (ODES_WARP_ANY(act && swapmask != 0u)) {
      if (act && swapmask != 0u) {
        _Pragma("unroll 1") for (int k = 0; k < NEQ - 1; k++) {
          const int lp = piv_get(k);
          if (lp != k) { const double t = FT(lp); FT(lp) = FT(k); FT(k) = t; }
        }
      }
      ODES_WARP_SYNC();
    }
    double b[A1(NEQ)];
    FORI b[i] = FT(i);
    _Pragma("unroll") for (int k = 0; k < NEQ - 1; k++) {
      double c[A1(NP)];
      mcol_load(k, c);
      const double mult = b[k];
      _Pragma("unroll") for (int i = 0; i < NEQ; i++) if (i > k) b[i] = b[i] + mult * c[i];
    }
    _Pragma("unroll") for (int k = NEQ - 1; k >= 0; k--) {
      double c[A1(NP)];
      mcol_load(k, c);
      const double bk = b[k] / c[k];
      b[k] = bk;
      const double mult = -bk;
      _Pragma("unroll") for (int i = 0; i < NEQ; i++) if (i < k) b[i] = b[i] + mult * c[i];
    }
    if (act) { FORI FT(i) = b[i]; }
  }
#elif ODES_LU_FORM == 1
  /* LU with partial pivoting of the lanes with act; returns 0, or k+1 when column k has no pivot.  Column-at-a-time
     form: for tensor memory (whole columns are what tcgen05.ld moves) and for the global scratch (a column is 8
     independent loads in flight; the element-wise form below measured 51 k vs 82 k trajectories/s on huang96) */
  DEV int gefa(bool act)
  {
    int ier = 0;
    if (act) piv_clear();
    _Pragma("unroll 1") for (int k = 0; k < NEQ - 1; k++) {
      /* column k: pivot search, row exchange, multipliers (they stay in registers for the update) */
      double m[A1(NP)];
      mcol_load(k, m);
      int lp = k; bool swap = false;
      if (act) {
        double big = 0.0, pval = 0.0, mk = 0.0;
        FORI if (i >= k) { const double v = fabs(m[i]); if (i == k || v > big) { big = v; lp = i; } }
        FORI { if (i == lp) pval = m[i]; if (i == k) mk = m[i]; }
        piv_set(k, lp);
        if (pval == 0.0) { ier = k + 1; act = false; }
        else {
          swap = (lp != k);
          FORI { if (swap && i == lp) m[i] = mk; if (i == k) m[i] = pval; }
          const double mult = -1.0 / pval;
          FORI if (i > k) m[i] = m[i] * mult;
        }
      }
      mcol_store(k, m);
      _Pragma("unroll 1") for (int j = k + 1; j < NEQ; j++) {
        double c[A1(NP)];
        mcol_load(j, c);
        if (act) {
          double a_kj = 0.0, ck = 0.0;
          FORI { if (i == lp) a_kj = c[i]; if (i == k) ck = c[i]; }
          if (swap) { FORI { if (i == lp) c[i] = ck; if (i == k) c[i] = a_kj; } }
          if (a_kj != 0.0) { FORI if (i > k) c[i] = c[i] + a_kj * m[i]; }
        }
        mcol_store(j, c);
      }
      la_fence(la);
    }
    {
      double c[A1(NP)];
      mcol_load(NEQ - 1, c);
      if (act) { piv_set(NEQ - 1, NEQ - 1); if (c[NEQ - 1] == 0.0) ier = NEQ; }
    }
    return ier;
  }
  /* solves M x = b in place, b = FT, for the lanes with act.  b lives in registers (static indices, the run-time
     pivot row is picked with selects); the k loops stay rolled, each body is one column load and NEQ
     independent multiply-adds */
  DEV void gesl(bool act)
  {
    double b[A1(NEQ)];
    FORI b[i] = FT(i);
    _Pragma(ODES_STR(unroll ODES_GESL_UNROLL)) for (int k = 0; k < NEQ - 1; k++) {
      double c[A1(NP)];
      mcol_load(k, c);
      const int lp = piv_get(k);
      double mult = 0.0, bk = 0.0;
      FORI { if (i == lp) mult = b[i]; if (i == k) bk = b[i]; }
      if (lp != k) { FORI { if (i == lp) b[i] = bk; if (i == k) b[i] = mult; } }
      FORI if (i > k) b[i] = b[i] + mult * c[i];
    }
    _Pragma(ODES_STR(unroll ODES_GESL_UNROLL)) for (int k = NEQ - 1; k >= 0; k--) {
      double c[A1(NP)];
      mcol_load(k, c);
      double bk = 0.0, ckk = 1.0;
      FORI if (i == k) { bk = b[i]; ckk = c[i]; }
      bk = bk / ckk;
      FORI if (i == k) b[i] = bk;
      const double mult = -bk;
      FORI if (i < k) b[i] = b[i] + mult * c[i];
    }
    if (act) { FORI FT(i) = b[i]; }
  }

#else
  /* The same two routines for linear-algebra words in shared memory: the run-time pivot row is simply indexed and
     the loops stay rolled. */
#define MELT(i, j) sm[(LAOFF + LAM + (j) * NP + (i)) * ODES_SSTRIDE]
  DEV int gefa(bool act)
  {
    if (!act) return 0;
    piv_clear();
    _Pragma("unroll 1") for (int k = 0; k < NEQ - 1; k++) {
      int lp = k;
      double big = fabs(MELT(k, k));
      _Pragma("unroll 4") for (int i = k + 1; i < NEQ; i++) { const double v = fabs(MELT(i, k)); if (v > big) { big = v; lp = i; } }
      piv_set(k, lp);
      const double pval = MELT(lp, k);
      if (pval == 0.0) return k + 1;
      const bool swap = (lp != k);
      if (swap) { MELT(lp, k) = MELT(k, k); MELT(k, k) = pval; }
      const double mult = -1.0 / pval;
      _Pragma("unroll 4") for (int i = k + 1; i < NEQ; i++) MELT(i, k) = MELT(i, k) * mult;
      _Pragma("unroll 1") for (int j = k + 1; j < NEQ; j++) {
        const double a_kj = MELT(lp, j);
        if (swap) { MELT(lp, j) = MELT(k, j); MELT(k, j) = a_kj; }
        if (a_kj != 0.0) { _Pragma("unroll 4") for (int i = k + 1; i < NEQ; i++) MELT(i, j) = MELT(i, j) + a_kj * MELT(i, k); }
      }
    }
    piv_set(NEQ - 1, NEQ - 1);
    if (MELT(NEQ - 1, NEQ - 1) == 0.0) return NEQ;
    return 0;
  }
  DEV void gesl(bool act)
  {
    if (!act) return;
    _Pragma("unroll 1") for (int k = 0; k < NEQ - 1; k++) {
      const int lp = piv_get(k);
      const double mult = FT(lp);
      if (lp != k) { FT(lp) = FT(k); FT(k) = mult; }
      _Pragma("unroll 4") for (int i = k + 1; i < NEQ; i++) FT(i) = FT(i) + mult * MELT(i, k);
    }
    _Pragma("unroll 1") for (int k = NEQ - 1; k >= 0; k--) {
      const double bk = FT(k) / MELT(k, k);
      FT(k) = bk;
      const double mult = -bk;
      _Pragma("unroll 4") for (int i = 0; i < k; i++) FT(i) = FT(i) + mult * MELT(i, k);
    }
  }
#endif

  /* CVDenseSetup (cvdense.c:368-412) for the lanes with act; returns 1 when the matrix is singular */
  DEV int lsetup(bool act, int cf)
  {
    const double dgamma = fabs((gamma / gammap) - 1.0);
    const bool jbad = act && ((nst == 0) || (nst > nstlj + K_MSBJ) || ((cf == CV_FAIL_BAD_J) && (dgamma < K_DGMAX_J)) || (cf == CV_FAIL_OTHER));
    const double mg = -gamma;
    if (jbad) { c_nje++; nstlj = nst; jcur = 1; } else if (act) jcur = 0;
#if HAS_JAC
    double jnz[A1(NSJP)];
    la_load<NSJCH>(la, LAJ, jnz);
    if (ODES_WARP_ANY(jbad)) {
      double yl[A1(NEQ)], jnew[A1(NNZ)];
#if ODES_LA_PACKED
      PvRef pv; pv.p = jnz + PVB_OFF; pv.s = 1u;          /* the constants came with the Jacobian's chunks */
#else
      double pvl[A1(PVB_WORDS)];
      la_load<PVB_CH>(la, PVB_BASE, pvl);
      PvRef pv; pv.p = pvl + PVB_OFF; pv.s = 1u;
#endif
      FORI yl[i] = ZN(0, i);
      if (jbad) {
        jacobi_f(tn, yl, pv, jnew);
        _Pragma("unroll") for (int k = 0; k < NNZ; k++) jnz[k] = jnew[k];
      }
      ODES_WARP_SYNC();
      la_store<NSJCH>(la, LAJ, jnz);
    }
#if ODES_LU_FORM == 2
    la_fence(la);
    return factor_columns(act, jnz, mg) > 0 ? 1 : 0;
#else
    /* M = I - gamma*J column by column: zero pattern, the stored non-zeros (DenseScale order), +1 on the diagonal */
    _Pragma("unroll 1") for (int j = 0; j < NEQ; j++) {
      double c[A1(NP)];
      mcol_load(j, c);
      if (act) {
        const double z = 0.0 * mg;
        _Pragma("unroll") for (int i = 0; i < NP; i++) c[i] = z;
        scatter_col(j, c, jnz, mg);
        FORI if (i == j) c[i] = c[i] + 1.0;
      }
      mcol_store(j, c);
    }
#endif
#else
    /* difference-quotient Jacobian, cvdense.c:479-537: y is z0 here; ACOR holds f(tn, z0) during the sweep.
       The saved Jacobian is dense: column j at LAJ + j*NP */
    const double srur = rsqrt_(UROUND);
    double minInc = 1.0;
    if (jbad) {
      const double fnorm = wrms(FTOFF);
      minInc = (fnorm != 0.0) ? (K_MIN_INC_MULT * fabs(h) * UROUND * NEQ * fnorm) : 1.0;
      FORC ACOR(i) = FT(i);
    }
    const bool anybad = ODES_WARP_ANY(jbad);
    _Pragma("unroll 1") for (int j = 0; j < NEQ; j++) {
      double jc[A1(NP)];
      la_load<NCH>(la, LAJ + j * NP, jc);
      if (anybad) {
        double yjsaved = 0.0, inc = 1.0;
        if (jbad) { yjsaved = YV(j); inc = fmax(srur * fabs(yjsaved), minInc / EW(j)); YV(j) = yjsaved + inc; }
        ODES_WARP_SYNC();
        f_conv(sm, la, tn, jbad);
        if (jbad) {
          YV(j) = yjsaved;
          const double inc_inv = 1.0 / inc;
          FORI jc[i] = inc_inv * (FT(i) - ACOR(i));
        }
        ODES_WARP_SYNC();
        la_store<NCH>(la, LAJ + j * NP, jc);
      }
      double c[A1(NP)];
      mcol_load(j, c);
      if (act) { FORI c[i] = jc[i] * mg; FORI if (i == j) c[i] = c[i] + 1.0; }
      mcol_store(j, c);
    }
    if (jbad) { FORC { FT(i) = ACOR(i); ACOR(i) = 0.0; } }
#endif
#if ODES_LU_FORM != 2
    la_fence(la);
    return gefa(act) > 0 ? 1 : 0;
#endif
  }

  /* ---- CVHin (cvode.c:1514-1636), warp-converged: lanes with act compute their initial step; the at most
     K_MAX_ITERS right-hand sides of the search are evaluated together.  returns false where CVHin fails ---- */
  DEV bool hin(bool act, double tout)
  {
    bool ok = true, looping = false;
    int sign = 1, count = 0;
    double hlb = 0.0, hub = 0.0, hg = 0.0, hnew = 0.0;
    if (act) {
      const double tdiff = tout - tn;
      sign = (tdiff > 0.0) ? 1 : -1;
      const double tdist = fabs(tdiff);
      const double tround = UROUND * fmax(fabs(tn), fabs(tout));
      if (tdiff == 0.0 || tdist < 2.0 * tround) ok = false;
      else {
        hlb = K_HLB_FACTOR * tround;
        /* CVUpperBoundH0 */
        double hub_inv = 0.0;
        FORC {
          double t1 = reltol * fabs(ZN(0, i)) + abstol;
          t1 = 1.0 / t1; t1 = 1.0 / t1;                               /* efun then N_VInv */
          t1 = K_HUB_FACTOR * fabs(ZN(0, i)) + t1;
          const double r = fabs(ZN(1, i)) / t1;
          if (fabs(r) > hub_inv) hub_inv = fabs(r);
        }
        hub = K_HUB_FACTOR * tdist;
        if (hub * hub_inv > 1.0) hub = 1.0 / hub_inv;
        hg = rsqrt_(hlb * hub);
        if (hub < hlb) { if (sign == -1) hg = -hg; h = hg; }
        else looping = true;
      }
    }
    _Pragma("unroll 1") for (int it = 0; it < K_MAX_ITERS; it++) {
      ODES_WARP_SYNC();
      if (!ODES_WARP_ANY(looping)) break;
      const double hgs = hg * sign;
      /* CVYddNorm */
      if (looping) { FORC YV(i) = hgs * ZN(1, i) + ZN(0, i); }
      ODES_WARP_SYNC();
      f(tn + hgs, looping);
      if (looping) {
        FORC FT(i) = FT(i) - ZN(1, i);
        { const double c = 1.0 / hgs; FORH FT(i) = c * FT(i); }
        const double yddnrm = wrms(FTOFF);
        hnew = (yddnrm * hub * hub > 2.0) ? rsqrt_(2.0 / yddnrm) : rsqrt_(hg * hub);
        count++;
        bool stop = (count >= K_MAX_ITERS);
        if (!stop) {
          const double hrat = hnew / hg;
          if ((hrat > 0.5) && (hrat < 2.0)) stop = true;
          else if ((count >= 2) && (hrat > 2.0)) { hnew = hg; stop = true; }
          else hg = hnew;
        }
        if (stop) {
          double h0 = K_H_BIAS * hnew;
          if (h0 < hlb) h0 = hlb;
          if (h0 > hub) h0 = hub;
          if (sign == -1) h0 = -h0;
          h = h0;
          looping = false;
        }
      }
    }
    ODES_WARP_SYNC();
    return ok;
  }

  /* ---- the blocks of one step attempt: CVStep's retry loop (cvode.c:1657-1710) with CVnlsNewton /
     CVNewtonIteration (cvode.c:2234-2371), CVDoErrorTest, CVCompleteStep and CVPrepareNextStep cut
     at the points where the right-hand side or the linear-solver setup is needed ---- */

  /* PRE: head of CVode's step loop + CVStep prologue + predictor + coefficients + setup decision.
     returns 0 or a CV_* failure flag */
  DEV int blk_pre(int mxstep)
  {
    if (fresh) {
      /* CVode loop head (cvode.c:1094-1130) */
      if (nst > 0 && !ewt_set()) return CV_ILL_INPUT;
      if (nstloc >= mxstep) return CV_TOO_MUCH_WORK;
      /* "too much accuracy" test UROUND*||zn[0]||_wrms > 1 (cvode.c:1119): the square root is only taken when the
         mean square is not provably small (mean square < 1e30  =>  UROUND*sqrt(mean square) < 0.23) */
      {
        double sum = 0.0;
        FORH { const double pz = ZN(0, i) * EW(i); sum += pz * pz; }
        if (!(sum / NEQ < 1e30) && UROUND * rsqrt_(sum / NEQ) > 1.0) return CV_TOO_MUCH_ACC;
      }
      /* CVStep prologue */
      saved_t = tn; ncf = 0; nef = 0; nflag = FIRST_CALL;
      if ((nst > 0) && (hprime != h)) adjust_params();
      fresh = 0;
    }
    predict();
    set_coeffs();
    convfail = ((nflag == FIRST_CALL) || (nflag == PREV_ERR_FAIL)) ? CV_NO_FAILURES : CV_FAIL_OTHER;
    callSetup = (nflag == PREV_CONV_FAIL) || (nflag == PREV_ERR_FAIL) || (nst == 0) ||
                (nst >= nstlp + K_MSBP) || (fabs(gamrat - 1.0) > K_DGMAX);
    /* start of the Newton iteration: acor = 0, y = predicted zn[0] */
    FORH { COLPTR; C(YOFF) = CZ(0); C(ACOFF) = 0.0; }
    nl_m = 0; nl_delp = 0.0;
    st = ST_F; fwhy = FW_ATTEMPT;
    return 0;
  }
  DEV void after_f()
  {
    if (fwhy == FW_RESTART) { FORC ZN(1, i) = h * FT(i); st = ST_PRE; }      /* order-1 restart: zn[1] = h f(tn, zn[0]) */
    else st = (fwhy == FW_ATTEMPT && callSetup) ? ST_SETUP : ST_NEWTON;
  }
  /* CVHandleNFlag for a recoverable failure (cvode.c:2400-2440); returns 0 or CV_CONV_FAILURE */
  DEV int conv_fail()
  {
    c_ncfn++;
    restore(saved_t);
    ncf++;
    etamax = 1.0;
    if ((fabs(h) <= 0.0) || (ncf == K_MXNCF)) return CV_CONV_FAILURE;          /* hmin = 0 */
    eta = K_ETACF;                                                            /* max(ETACF, hmin/|h|) */
    nflag = PREV_CONV_FAIL;
    rescale();
    st = ST_PRE;
    return 0;
  }
  /* SETUP: CVDenseSetup + bookkeeping of CVnlsNewton (cvode.c:2262-2275), warp-converged, for the lanes with act */
  DEV int blk_setup(bool act)
  {
    const int ier = lsetup(act, convfail);
    if (!act) return 0;
    c_nsetups++;
    callSetup = 0;
    gamrat = crate = 1.0;
    gammap = gamma;
    nstlp = nst;
    waited = 0;
    st = ST_NEWTON;
    if (ier > 0) return conv_fail();
    return 0;
  }
  /* NEWTON: one pass of CVNewtonIteration's loop body (cvode.c:2320-2371) with CVDenseSolve (cvdense.c:423-441),
     then CVDoErrorTest when it converged.  returns 0 or a CV_* failure flag */
  DEV int blk_newton(bool act)
  {
    /* b = gamma*f(tn,y) - (rl1*zn[1] + acor)   (N_VLinearSum with a == 1 multiplies by 1.0: same value) */
    if (act) { FORH { COLPTR; const double tv = rl1 * CZ(1) + C(ACOFF); C(FTOFF) = gamma * C(FTOFF) - tv; } }
    ODES_WARP_SYNC();
    gesl(act);                                     /* warp-converged: the matrix columns come through la_load */
    if (!act) return 0;
    if (gamrat != 1.0) { const double c = 2.0 / (1.0 + gamrat); FORC FT(i) = c * FT(i); }
    c_nni++;
    double del;
    {
      double sum = 0.0;
      FORH {
        COLPTR;
        const double bi = C(FTOFF), p = bi * C(EWOFF);
        sum += p * p;
        const double ac = C(ACOFF) + bi;
        C(ACOFF) = ac;
        C(YOFF) = CZ(0) + ac;
      }
      del = rsqrt_(sum / NEQ);
    }
    if (nl_m > 0) crate = fmax(K_CRDOWN * crate, del / nl_delp);
    const double dcon = del * fmin(1.0, crate) / TQ(4);
    if (dcon > 1.0) {
      nl_m++;
      if ((nl_m == K_NLS_MAXCOR) || ((nl_m >= 2) && (del > K_RDIV * nl_delp))) {
        if (jcur) return conv_fail();
        /* out-of-date Jacobian: same predicted values, new setup (cvode.c:2283-2290) */
        FORC { YV(i) = ZN(0, i); ACOR(i) = 0.0; }
        nl_m = 0; nl_delp = 0.0;
        callSetup = 1; convfail = CV_FAIL_BAD_J;
        st = ST_F; fwhy = FW_ATTEMPT;
        return 0;
      }
      nl_delp = del;
      st = ST_F; fwhy = FW_NEWTON;
      return 0;
    }
    acnrm = (nl_m == 0) ? del : wrms(ACOFF);
    jcur = 0;
    /* CVDoErrorTest (cvode.c:2470-2526) */
    dsm = acnrm / TQ(2);
    if (dsm <= 1.0) { st = ST_COMPLETE; return 0; }
    nef++; c_netf++;
    nflag = PREV_ERR_FAIL;
    restore(saved_t);
    if ((fabs(h) <= 0.0) || (nef == K_MXNEF)) return CV_ERR_FAILURE;
    etamax = 1.0;
    st = ST_PRE;
    if (nef <= K_MXNEF1) {
      eta = 1.0 / (root_pow(K_BIAS2 * dsm, L) + K_ADDON);
      eta = fmax(K_ETAMIN, fmax(eta, 0.0));
      if (nef >= K_SMALL_NEF) eta = fmin(eta, K_ETAMXF);
      rescale();
      return 0;
    }
    if (q > 1) {
      eta = K_ETAMIN;
      apply_pending();                               /* the order change works on the restored history */
      adjust_order(-1);
      L = q; q--; qwait = L;
      rescale();
      return 0;
    }
    /* order 1 after repeated failures: reload the first derivative through the F block (rare) */
    eta = K_ETAMIN;
    h *= eta;
    hscale = h;
    qwait = K_LONG_WAIT;
    apply_pending();
    FORC YV(i) = ZN(0, i);
    st = ST_F; fwhy = FW_RESTART;
    return 0;
  }
  /* COMPLETE: CVCompleteStep + CVPrepareNextStep (cvode.c:2540-2713), warp-converged for the lanes with act (the
     pow calls of the step-size / order selection are shared by the warp, see pow_service) */
  DEV void blk_complete(const bool act, double tout)
  {
    double px[3] = {0.0, 0.0, 0.0};
    int pn[3] = {1, 1, 1}, nreq = 0;
    bool want_q = false, want_m = false, want_p = false, select = false;
    if (act) {
    nst++; c_nst++;
    hu = h;
    _Pragma("unroll 1") for (int i = q; i >= 2; i--) TAU(i) = TAU(i - 1);
    if ((q == 1) && (nst > 1)) TAU(2) = TAU(1);
    TAU(1) = h;
    /* zn[j] += l[j]*acor, j = 0..q  (l[j] == 1 multiplies by 1.0: same value as the reference's special case) */
    qwait--;
    const bool save5 = (qwait == 1) && (q != QMAX);
    {
      const double l0 = LL(0), l1 = LL(1), l2 = LL(2), l3 = LL(3), l4 = LL(4), l5 = LL(5);
      const int qq = q;
      /* when this step reaches the next output time, CVodeGetDky (cvode.c:1228-1279) is evaluated right here on the
         freshly updated column (Horner with s = (tout - tn)/h); later outputs of the same step use get_dky0 */
#if USE_TSTOP
      const bool outnow = false;
#else
      const bool outnow = (tn - tout) * h >= 0.0;
#endif
      const double sh = (tout - tn) / h;
      const bool p2 = qq >= 2, p3 = qq >= 3, p4 = qq >= 4, p5 = qq >= 5;
      FORH {
        COLPTR;
        const double ac = C(ACOFF);
        const double z0 = l0 * ac + CZ(0), z1 = l1 * ac + CZ(1), z2 = l2 * ac + CZ(2), z3 = l3 * ac + CZ(3),
                     z4 = l4 * ac + CZ(4), z5 = l5 * ac + CZ(5);
        CZ(0) = z0; CZ(1) = z1;
        if (p2) CZ(2) = z2;
        if (p3) CZ(3) = z3;
        if (p4) CZ(4) = z4;
        if (p5) CZ(5) = z5;
        if (save5) CZ(QMAX) = ac;
        if (outnow) {
          double d = z1;                                        /* q == 1 */
          if (p5) d = z5;
          if (p4) d = (qq == 4) ? z4 : sh * d + z4;
          if (p3) d = (qq == 3) ? z3 : sh * d + z3;
          if (p2) d = (qq == 2) ? z2 : sh * d + z2;
          if (p2) d = sh * d + z1;
          C(YOFF) = sh * d + z0;
        }
      }
      ydone = outnow ? 1 : 0;
    }
    if (save5) saved_tq5 = TQ(5);
    /* CVPrepareNextStep (cvode.c:2577-2713): the arguments of the pow calls first */
    if (etamax == 1.0) { qwait = qwait > 2 ? qwait : 2; qprime = q; hprime = h; eta = 1.0; }
    else {
      select = (qwait == 0);
      px[0] = K_BIAS2 * dsm; pn[0] = L;
      want_q = !eta_below_thresh(px[0], L);
      nreq = want_q ? 1 : 0;
      if (select) {
        if (q > 1) {
          const double ddn = wrms(ZOFF + q * NEQ) / TQ(1);
          const double xm = K_BIAS1 * ddn;
          want_m = !eta_below_thresh(xm, q);
          if (want_m) { px[nreq] = xm; pn[nreq] = q; nreq++; }
        }
        if (q != QMAX) {
          double hr = h / TAU(2), pw = 1.0;
          _Pragma("unroll 1") for (int i = 1; i <= L; i++) pw *= hr;
          const double cquot = (TQ(5) / saved_tq5) * pw;
          double sum = 0.0;
          FORC { COLPTR; const double tv = (-cquot) * CZ(QMAX) + C(ACOFF), p = tv * C(EWOFF); sum += p * p; }
          const double dup = rsqrt_(sum / NEQ) / TQ(3);
          const double xp = K_BIAS3 * dup;
          want_p = !eta_below_thresh(xp, L + 1);
          if (want_p) { px[nreq] = xp; pn[nreq] = L + 1; nreq++; }
        }
      }
    }
    }
    ODES_WARP_SYNC();
    pow_service(nreq, px, pn);
    if (!act) return;
    if (etamax != 1.0) {
      /* a factor known to be below THRESH is represented by 0: it is neither taken nor the largest of the three */
      int r = 0;
      const double etaq = want_q ? 1.0 / (px[r++] + K_ADDON) : 0.0;
      if (!select) { eta = etaq; qprime = q; }
      else {
        qwait = 2;
        const double etaqm1 = want_m ? 1.0 / (px[r++] + K_ADDON) : 0.0;
        const double etaqp1 = want_p ? 1.0 / (px[r++] + K_ADDON) : 0.0;
        const double etam = fmax(etaqm1, fmax(etaq, etaqp1));
        if (etam < K_THRESH) { eta = 1.0; qprime = q; }
        else if (etam == etaq) { eta = etaq; qprime = q; }
        else if (etam == etaqm1) { eta = etaqm1; qprime = q - 1; }
        else { eta = etaqp1; qprime = q + 1; FORC ZN(QMAX, i) = ACOR(i); }
      }
      /* CVSetEta (hmax_inv = 0) */
      if (eta < K_THRESH) { eta = 1.0; hprime = h; }
      else { eta = fmin(eta, etamax); eta /= fmax(1.0, 0.0); hprime = h * eta; }
    }
    etamax = (nst <= K_SMALL_NST) ? K_ETAMX2 : K_ETAMX3;
    /* acor is rescaled to the local error estimate in CVODE; it is not read again before being reset */
    nstloc++;
    fresh = 1;
    st = ST_PRE;
  }

  /* CVodeGetDky with k = 0 (cvode.c:1228-1279): Horner on the columns of the Nordsieck array, result in y */
  DEV bool get_dky0(double t)
  {
    double tfuzz = K_FUZZ_FACTOR * UROUND * (fabs(tn) + fabs(hu));
    if (hu < 0.0) tfuzz = -tfuzz;
    const double tp = tn - hu - tfuzz, tn1 = tn + tfuzz;
    if ((t - tp) * (t - tn1) > 0.0) return false;
    const double s = (t - tn) / h;
    const int qq = q;
    FORC {
      COLPTR;
      double d = CZ(1);                                       /* q == 1 */
      if (qq >= 5) d = CZ(5);
      if (qq >= 4) d = (qq == 4) ? CZ(4) : s * d + CZ(4);
      if (qq >= 3) d = (qq == 3) ? CZ(3) : s * d + CZ(3);
      if (qq >= 2) d = (qq == 2) ? CZ(2) : s * d + CZ(2);
      if (qq >= 2) d = s * d + CZ(1);
      C(YOFF) = s * d + CZ(0);
    }
    return true;
  }

  /* fresh solver memory for a new trajectory (CVodeCreate/CVodeMalloc defaults) */
  DEV void init_state()
  {
    c_nst = c_nfe = c_nsetups = c_nje = c_nni = c_ncfn = c_netf = 0;
    jcur = 0; gammap = 1.0; gamrat = 1.0; crate = 1.0; saved_tq5 = 1.0; acnrm = 0.0; dsm = 0.0;
    nst = 0; fresh = 1; nstloc = 0; ncf = 0; nef = 0; nflag = FIRST_CALL; saved_t = 0.0; nl_m = 0; nl_delp = 0.0;
    fwhy = FW_ATTEMPT; callSetup = 0; convfail = CV_NO_FAILURES; waited = 0; pend = 0; ydone = 0; resc_eta = 1.0;
    h = 0.0; hprime = 0.0; hscale = 0.0; hu = 0.0; eta = 1.0; tn = 0.0; q = 1; L = 2; qprime = 1; qwait = 2;
    etamax = K_ETAMX1; rl1 = 1.0; gamma = 0.0; tstop = 0.0; nstlp = 0; nstlj = 0;
    piv_clear(); swapmask = 0u;
    clear_history();
  }

  /* CVodeMalloc / CVodeReInit state (cvode.c:416-452, :588-617) + CVDenseInit; y0 = YV */
  DEV void reinit(double t0)
  {
    tn = t0;
    q = 1; L = 2; qwait = L; etamax = K_ETAMX1;
    hu = 0.0;
    FORC ZN(0, i) = YV(i);
    nst = 0; nstlp = 0; nstlj = 0;
    fresh = 1;
  }

  /* first-call block of CVode() (cvode.c:925-1010), warp-converged, for the lanes with act: weights, f(t0,y0),
     initial step, tstop clamp.  returns CV_SUCCESS or CV_ILL_INPUT */
  DEV int first_call(bool act, double tout)
  {
    int rc = CV_SUCCESS;
    if (act && !ewt_set()) { rc = CV_ILL_INPUT; act = false; }
    ODES_WARP_SYNC();
    f(tn, act);                                      /* y == zn[0] here */
    if (act) { FORC ZN(1, i) = FT(i); h = 0.0; }
    const bool ok = hin(act, tout);
    if (!act) return rc;
    if (!ok) return CV_ILL_INPUT;
#if USE_TSTOP
    if ((tstop - tn) * h < 0.0) return CV_ILL_INPUT;
    if ((tn + h - tstop) * h > 0.0) h = tstop - tn;
#endif
    hscale = h; hprime = h;
    FORC ZN(1, i) = h * ZN(1, i);
    nstloc = 0;
    return CV_SUCCESS;
  }

  /* entry of a later CVode() call (nst > 0), cvode.c:1040-1090: returns 1 when tout (or tstop) is
     already covered by the last step (interpolate without stepping), 0 to go on stepping, <0 failure */
  DEV int next_call(double tout, double &tret)
  {
    nstloc = 0;
#if USE_TSTOP
    if ((tstop - tn) * h < 0.0) return CV_ILL_INPUT;
#endif
    if ((tn - tout) * h >= 0.0) { tret = tout; return 1; }
#if USE_TSTOP
    {
      const double troundoff = K_FUZZ_FACTOR * UROUND * (fabs(tn) + fabs(h));
      if (fabs(tn - tstop) <= troundoff) { tret = tstop; return 1; }
      if ((tn + hprime - tstop) * h > 0.0) { hprime = tstop - tn; eta = hprime / h; }
    }
#endif
    return 0;
  }

  /* after a completed step (cvode.c:1170-1206): 1 when tout/tstop has been reached */
  DEV int after_step(double tout, double &tret)
  {
#if USE_TSTOP
    {
      const double troundoff = K_FUZZ_FACTOR * UROUND * (fabs(tn) + fabs(h));
      if (fabs(tn - tstop) <= troundoff) { tret = tstop; return 1; }
      if ((tn + hprime - tstop) * h > 0.0) { hprime = tstop - tn; eta = hprime / h; }
    }
#endif
    if ((tn - tout) * h >= 0.0) { tret = tout; return 1; }
    return 0;
  }
};

/* development instrumentation (-DODES_LANE_PROFILE=1): per block of the pass, how often it ran and with how many lanes;
   the last block of the grid prints the totals */
#ifndef ODES_LANE_PROFILE
#define ODES_LANE_PROFILE 0
#endif
#if ODES_LANE_PROFILE && !defined(ODES_HOST_EMULATION)
__device__ unsigned long long odes_prof[40];
__device__ unsigned int odes_prof_done;
#define PROF_DECL unsigned int prof_runs[16], prof_lanes[16]; _Pragma("unroll") for (int k_ = 0; k_ < 16; k_++) { prof_runs[k_] = 0u; prof_lanes[k_] = 0u; }
#define PROF(k, pred) do { const unsigned int n_ = ODES_WARP_COUNT(pred); prof_runs[k] += 1u; prof_lanes[k] += n_; } while (0)
#define PROF_IF_ANY(k, pred) do { if (ODES_WARP_ANY(pred)) PROF(k, pred); } while (0)
#define PROF_FLUSH do { if ((threadIdx.x & 31u) == 0u) { _Pragma("unroll") for (int k_ = 0; k_ < 16; k_++) { atomicAdd(&odes_prof[k_], (unsigned long long)prof_runs[k_]); atomicAdd(&odes_prof[16 + k_], (unsigned long long)prof_lanes[k_]); } } } while (0)
#else
#define PROF_DECL
#define PROF(k, pred) ((void)0)
#define PROF_IF_ANY(k, pred) ((void)0)
#define PROF_FLUSH ((void)0)
#endif
/* One lane integrates one trajectory at a time: reset, set varied values, integrate over the output grid with
   the reference's output-time event handling (src/odeSolver.c:313-330, src/integratorInstance.c:1088-1237).
   Lanes are persistent: a lane whose trajectory is finished takes the next set index from a global counter
   (warp-aggregated atomicAdd), so a warp does not idle behind its slowest trajectory (step counts differ by more
   than 2x between parameter sets).  See "warp scheduling" at the top of the file for the control structure.
   Blocks that move linear-algebra words (REFILL, START, F, SETUP, NEWTON, OUTPUT) are entered by the whole warp
   under a warp-uniform condition and predicate per lane inside, because those moves are warp-wide. */
DEV void integrate_lanes(const OdesArgs &a, double *smem_lane, La la)
{
  Solver s;
#if !ODES_MAT_LOCAL
  s.sm = smem_lane;
#else
  (void)smem_lane;
  s.sm = s.mat;
#endif
  double *const sm = s.sm;
  la.sm = sm;
  s.la = la;
  s.reltol = a.reltol; s.abstol = a.abstol;
  const size_t stride = (size_t)a.stride;
  int trigger[A1(NEVENTS)];
  _Pragma("unroll") for (int e = 0; e < NEVENTS; e++) trigger[e] = 0;
  double t = 0.0, tout = 0.0;
  int status = 0, iout = 1, idle = 0;
  unsigned int firedacc = 0u;             /* events fired since the last kept row (StepResolution) */
#if ODES_ROOTS
  bool pending = false;                   /* an event has been located at troot <= tn and waits for the rows before it */
  double troot = 0.0, rlo = 0.0;          /* its time; lower end of the interval still to be searched */
#endif
  size_t set = 0;
  bool exhausted = false;
  s.st = ST_DONE;
  s.init_state();
  PROF_DECL

  for (;;) {
    ODES_WARP_SYNC();
    PROF(0, true); PROF(8, s.st == ST_DONE && !exhausted); PROF(9, s.st == ST_DONE && exhausted); PROF(10, s.st == ST_SETUP); PROF(11, s.st == ST_F); PROF(12, s.st == ST_PRE);
    /* ---- FINISH + REFILL: close finished trajectories, hand out new set indices ---- */
    if (s.st == ST_FINISH) {
      /* rows never reached (solver failure or halt on event) are marked not-a-number */
      for (size_t r = ODES_ROW(iout + ODES_STEPRES - 1); r < ODES_ROWS(a); r++) {
        _Pragma("unroll 1") for (int k = 0; k < NSTORE; k++) a.out[(set * ODES_ROWS(a) + r) * NSTORE + k] = __longlong_as_double(0x7ff8000000000000LL);
        if (a.fired) a.fired[r * stride + set] = 0u;
      }
      iout = a.nout + 1;
      if (a.status) a.status[set] = status;
      if (a.stats) {
        a.stats[0 * stride + set] = s.c_nst; a.stats[1 * stride + set] = s.c_nfe; a.stats[2 * stride + set] = s.c_nsetups;
        a.stats[3 * stride + set] = s.c_nje; a.stats[4 * stride + set] = s.c_nni; a.stats[5 * stride + set] = s.c_ncfn;
        a.stats[6 * stride + set] = s.c_netf;
      }
      odes_signal_done(a, set);
      s.st = ST_DONE; idle = 0;
    }
    ODES_WARP_SYNC();
    {
      const bool want = (s.st == ST_DONE) && !exhausted;
#if ODES_LOCKSTEP && !defined(ODES_HOST_EMULATION)
      const int nwant_blk = __syncthreads_count(want);
      const int nwant = ODES_WARP_COUNT(want);
      if (nwant_blk > 0) {
        const bool run = (nwant_blk >= ODES_REFILL_MIN * (ODES_BLOCK / 32)) || (__syncthreads_count(want && idle >= ODES_REFILL_WAIT) > 0);
        if (run && nwant > 0) {
#else
      const int nwant = ODES_WARP_COUNT(want);
      if (nwant > 0) {
        const bool run = (nwant >= ODES_REFILL_MIN) || ODES_WARP_ANY(want && idle >= ODES_REFILL_WAIT) || ODES_WARP_ALL(s.st == ST_DONE);
        if (run) {
#endif
          const unsigned long long next = odes_take_sets(a.counter, want, nwant);
          const bool take = want && next < (unsigned long long)a.nsets;
          if (want && !take) exhausted = true;
          double pvl[A1(PVB_WORDS)];
          _Pragma("unroll") for (int k = 0; k < PVB_WORDS; k++) pvl[k] = 0.0;
          if (take) {
            set = (size_t)next;
            s.init_state();
            _Pragma("unroll") for (int e = 0; e < NEVENTS; e++) trigger[e] = c_trigger0[e];
            t = a.tpoints[0]; tout = t; status = 0; iout = 1;
            double yl[A1(NEQ)];
            FORI yl[i] = 0.0;
            PvRef pv; pv.p = pvl + PVB_OFF; pv.s = 1u;
            model_init(yl, pv, a.vary, stride, set);
            store_row0(a.out + set * ODES_ROWS(a) * NSTORE, 1, yl, pv);
            firedacc = 0u;
#if ODES_ROOTS
            pending = false;
#endif
            if (a.fired) a.fired[set] = 0u;
            /* static assigned values (rules not recomputed before the ODEs) for this trajectory's inputs */
            double atmp[A1(NASS)];
            rules_all(a.tpoints[0], yl, pv, atmp);
            FORI YV(i) = yl[i];
            s.st = (a.nout >= 1) ? ST_START : ST_FINISH;
          }
          ODES_WARP_SYNC();
          la_commit<PVB_CH>(la, PVB_BASE, pvl, take);
        } else if (want) idle++;
        ODES_WARP_SYNC();
      }
    }
#if ODES_LOCKSTEP && !defined(ODES_HOST_EMULATION)
    /* block-level lock step: all warps of the block start every pass together, so warps that share a scheduler
       run through the same instructions at about the same time and share their instruction fetches */
    if (__syncthreads_and(s.st == ST_DONE && exhausted)) break;
#else
    if (ODES_WARP_ALL(s.st == ST_DONE && exhausted)) break;
#endif
#if NEQ > 0
    /* ---- START: solver (re)start at t from y[] (first call, or after an event assignment) ---- */
    if (ODES_WARP_ANY(s.st == ST_START)) {
      const bool act = (s.st == ST_START);
      PROF(1, act);
      if (act) {
#if ODES_ROOTS
        pending = false;
#endif
        s.reinit(t);
        tout = a.tpoints[iout];
#if USE_TSTOP
        s.tstop = tout;
#endif
      }
      const int fl = s.first_call(act, tout);
      if (act) { if (fl < 0) { status = fl; s.st = ST_FINISH; } else s.st = ST_PRE; }
      ODES_WARP_SYNC();
    }
    /* ---- PRE ---- */
    ODES_BLOCK_ALIGN(1);
    PROF_IF_ANY(2, s.st == ST_PRE);
    if (s.st == ST_PRE) {
      const int rc = s.blk_pre(a.mxstep);
      if (rc < 0) { status = rc; s.st = ST_FINISH; }
    }
    ODES_WARP_SYNC();
    /* ---- F, SETUP, NEWTON.  The pair F / NEWTON can run ODES_NEWTON_REPS times per pass, so that a lane whose first
       corrector iteration has not converged (one step attempt in five on MAPK) gets its second right-hand side and solve
       in the same pass instead of sitting out the PRE and COMPLETE blocks of the next one.  Measured on MAPK (B200, 10^6
       sets): 2 repetitions cut the passes per trajectory by 17 % (PRE runs with 25 instead of 21 lanes) but run F and NEWTON
       1.6 times as often at 16 instead of 25 lanes: 228.6 ms vs 224.2 ms with one, 239.1 ms with three -- one it stays. ---- */
#ifndef ODES_NEWTON_REPS
#define ODES_NEWTON_REPS 1
#endif
    _Pragma("unroll 1") for (int rep = 0; rep < ODES_NEWTON_REPS; rep++) {
    /* ---- F: the model's right-hand side ---- */
    ODES_BLOCK_ALIGN(2);
    if (ODES_WARP_ANY(s.st == ST_F)) {
      const bool act = (s.st == ST_F);
      PROF(3, act);
      s.f(s.tn, act);
      if (act) s.after_f();
      ODES_WARP_SYNC();
    }
    /* ---- SETUP: run when enough lanes want it, one of them has waited long enough, or nobody
       else can make progress; waiting lanes simply re-enter here in the next pass.  In lock step the vote is taken
       over the whole block, so its warps factor together and stay aligned ---- */
    if (rep == 0) {
      const bool want = (s.st == ST_SETUP);
#if ODES_LOCKSTEP && !defined(ODES_HOST_EMULATION)
      const int nwant = __syncthreads_count(want);
      if (nwant > 0) {
        const bool run = (nwant >= ODES_SETUP_MIN * (ODES_BLOCK / 32)) || (__syncthreads_count(want && s.waited >= ODES_SETUP_WAIT) > 0);
#else
      const int nwant = ODES_WARP_COUNT(want);
      if (nwant > 0) {
        const bool run = (nwant >= ODES_SETUP_MIN) || ODES_WARP_ANY(want && s.waited >= ODES_SETUP_WAIT) ||
                         ODES_WARP_ALL(want || s.st == ST_DONE);
#endif
        if (run) { PROF(4, want); const int rc = s.blk_setup(want); if (want && rc < 0) { status = rc; s.st = ST_FINISH; } }
        else if (want) s.waited++;
        ODES_WARP_SYNC();
      }
    }
    /* ---- NEWTON ---- */
    ODES_BLOCK_ALIGN(4);
    if (ODES_WARP_ANY(s.st == ST_NEWTON)) {
      const bool act = (s.st == ST_NEWTON);
      PROF(5, act);
      const int rc = s.blk_newton(act);
      if (act && rc < 0) { status = rc; s.st = ST_FINISH; }
      ODES_WARP_SYNC();
    }
    }
    /* ---- COMPLETE ---- */
    ODES_BLOCK_ALIGN(8);
    if (ODES_WARP_ANY(s.st == ST_COMPLETE)) {
      const bool act = (s.st == ST_COMPLETE);
      PROF(6, act);
      s.blk_complete(act, tout);
#if ODES_ROOTS
      if (act) { rlo = s.tn - s.hu; s.st = ST_ROOT; }
#else
      if (act && s.after_step(tout, t)) s.st = ST_OUTPUT;
#endif
      ODES_WARP_SYNC();
    }
#if ODES_ROOTS
    /* ---- ROOT: which triggers became true over the last step, and where (first crossing in (rlo, tn]) ---- */
    if (ODES_WARP_ANY(s.st == ST_ROOT)) {
      const bool act = (s.st == ST_ROOT);
      double pvl[A1(PVB_WORDS)];
      la_load<PVB_CH>(la, PVB_BASE, pvl);
      PvRef pv; pv.p = pvl + PVB_OFF; pv.s = 1u;
      if (act) {
        double yl[A1(NEQ)];
        unsigned int known = 0u;
        _Pragma("unroll") for (int e = 0; e < NEVENTS; e++) if (trigger[e]) known |= 1u << e;
        FORI yl[i] = ZN(0, i);
        const unsigned int cand = trigger_f(s.tn, yl, pv) & ~known;
        if (cand) {
          double lo = rlo, hi = s.tn;
          const double ttol = 100.0 * UROUND * (fabs(s.tn) + fabs(s.hu));
          _Pragma("unroll 1") for (int it = 0; it < 200 && (hi - lo) > ttol; it++) {
            const double mid = lo + 0.5 * (hi - lo);
            (void)s.get_dky0(mid);
            FORI yl[i] = YV(i);
            if (trigger_f(mid, yl, pv) & cand) hi = mid; else lo = mid;
          }
          troot = hi; pending = true;
        }
        s.ydone = 0;                                     /* y has served as scratch */
        s.st = (s.after_step(tout, t) || pending) ? ST_OUTPUT : ST_PRE;
      }
      ODES_WARP_SYNC();
    }
#endif
#else
    if (s.st != ST_DONE && s.st != ST_FINISH) { t = a.tpoints[iout]; s.st = ST_OUTPUT; }
#endif
    /* ---- OUTPUT: rows covered by the last step ---- */
    if (ODES_WARP_ANY(s.st == ST_OUTPUT)) {
      PROF(7, s.st == ST_OUTPUT);
      double pvl[A1(PVB_WORDS)];
      /* the per-trajectory constants are fetched only where the output-time code reads them: stored assigned values or
         constants, events, the steady-state test (three tensor-memory loads per pass otherwise wasted) */
#ifndef STORE_NEEDS_PV
#define STORE_NEEDS_PV 1
#endif
#if STORE_NEEDS_PV || NEVENTS > 0 || HALT_ON_STEADY
      la_load<PVB_CH>(la, PVB_BASE, pvl);
#else
      _Pragma("unroll") for (int k = 0; k < PVB_WORDS; k++) pvl[k] = 0.0;
#endif
      PvRef pv; pv.p = pvl + PVB_OFF; pv.s = 1u;
      bool wrote = false;
      while (s.st == ST_OUTPUT) {
        double yl[A1(NEQ)];
#if ODES_ROOTS
        {
          const bool reached = (s.tn - tout) * s.h >= 0.0;
          if (pending && (!reached || troot <= tout)) {
            /* the located event: state at troot from the dense output, the reference's event handling there */
            int re = 0;
            (void)s.get_dky0(troot);
            FORI yl[i] = YV(i);
            const unsigned int fm = a.pad ? 0u : event_f(troot, yl, pv, trigger, re);
            firedacc |= fm;
            if (fm) wrote = true;
            pending = false;
            if (re) { FORI YV(i) = yl[i]; t = troot; s.st = ST_START; }        /* restart from the located time */
            else { rlo = troot; s.st = ST_ROOT; }                              /* later crossings of the same step */
            break;
          }
          if (!reached) { s.st = ST_PRE; break; }
          t = tout;
        }
#endif
#if NEQ > 0
        if (!s.ydone) (void)s.get_dky0(t);
        s.ydone = 0;
        FORI yl[i] = YV(i);
#endif
        unsigned int firedmask = 0u;
        int restart = 0;
#if NEVENTS > 0
        if (!a.pad) firedmask = event_f(t, yl, pv, trigger, restart);          /* a.pad: event processing switched off for this launch */
        if (firedmask) wrote = true;
        if (restart) { FORI YV(i) = yl[i]; }
#endif
        firedacc |= firedmask;
        if (ODES_KEEP(iout)) {
          store_row(a.out + (set * ODES_ROWS(a) + ODES_ROW(iout)) * NSTORE, 1, t, yl, pv);
          if (a.fired) a.fired[ODES_ROW(iout) * stride + set] = firedacc;
          firedacc = 0u;
        }
        iout++;
#if HALT_ON_EVENT
        if (firedmask) { s.st = ST_FINISH; break; }
#endif
#if HALT_ON_STEADY
        if (steady_f(t, yl, pv)) { s.st = ST_FINISH; break; }       /* approximate steady state: the run ends with this row */
#endif
        if (iout > a.nout) { s.st = ST_FINISH; break; }
#if NEQ > 0
        if (restart) { s.st = ST_START; break; }    /* event assignment: the solver restarts from y[] */
        tout = a.tpoints[iout];
#if USE_TSTOP
        s.tstop = tout;
#endif
        const int r = s.next_call(tout, t);
        if (r < 0) { status = r; s.st = ST_FINISH; break; }
#if ODES_ROOTS
        if (r == 0 && !pending) s.st = ST_PRE;      /* a located event before the next output time is handled at the loop's head */
#else
        if (r == 0) s.st = ST_PRE;
#endif
#else
        t = a.tpoints[iout];
#endif
      }
      ODES_WARP_SYNC();
#if NEVENTS > 0
      /* event assignments may have changed per-trajectory constants or refreshed assigned values */
      la_commit<PVB_CH>(la, PVB_BASE, pvl, wrote);
#else
      (void)wrote;
#endif
    }
  }
  PROF_FLUSH;
}

#ifndef ODES_HOST_EMULATION
extern "C" __global__ void __launch_bounds__(ODES_BLOCK, ODES_MINBLOCKS) odes_bdf_kernel(const OdesArgs a)
{
  extern __shared__ double odes_smem[];
  La la;
  la.sm = odes_smem + threadIdx.x;
  const unsigned int slot = blockIdx.x * blockDim.x + threadIdx.x;
  la.gm = a.scratch + slot; la.gs = gridDim.x * blockDim.x;
  la.taddr = 0u;
#if ODES_LA == 2
  /* tensor memory: warp 0 allocates ODES_TMEM_COLS columns for the block, every warp then works on its own 32
     TMEM lanes (warp w of a CTA can reach lanes 32*(w%4) .. 32*(w%4)+31, hence ODES_BLOCK <= 128) */
  __shared__ unsigned int tmem_base;
  if ((threadIdx.x >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "l"((unsigned long long)__cvta_generic_to_shared(&tmem_base)), "n"(ODES_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  la.taddr = tmem_base + ((((threadIdx.x >> 5) & 3u) * 32u) << 16) + (threadIdx.x >> 7) * (unsigned int)(2 * ODES_LA_WORDS);
#endif
  integrate_lanes(a, odes_smem + threadIdx.x, la);
#if ODES_LANE_PROFILE
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&odes_prof_done, 1u) + 1u == gridDim.x) {
      __threadfence();
      static const char *const nm[16] = {"pass", "START", "PRE", "F", "SETUP", "NEWTON", "COMPLETE", "OUTPUT", "idle:refill", "idle:exhausted", "wait:setup", "in:F", "in:PRE", "-", "-", "-"};
      for (int k = 0; k < 13; k++) printf("PROF %-16s runs %12llu lanes %14llu avg %.2f\n", nm[k], odes_prof[k], odes_prof[16 + k], odes_prof[k] ? (double)odes_prof[16 + k] / (double)odes_prof[k] : 0.0);
      for (int k = 0; k < 32; k++) odes_prof[k] = 0ull;
      odes_prof_done = 0u;
    }
  }
#endif
#if ODES_LA == 2
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "n"(ODES_TMEM_COLS));
#endif
}
#endif
#endif /* !ODES_WARP */
