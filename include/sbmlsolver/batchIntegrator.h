/*
 * batchIntegrator.h -- the NEW batched entry of the library (C ABI).
 *
 * It replaces the serial design-point loop of the reference,
 *     for i in designpoints: setVariableValue x nvary; integrateOneStep until
 *     timeCourseCompleted; harvest; reset          (src/odeSolver.c:313-330,
 *                                                   examples/ParameterScanner.c:136-163)
 * by one kernel launch per device in which every trajectory is integrated by
 * a CVODE-style variable-order BDF (vendored SUNDIALS 2.1.1:
 * Win32/WinCVODE/sundials/cvode/source/cvode.c:875-2713, cvdense.c:368-441).
 * Per set the semantics are exactly `reset; setVariableValue x nvary;
 * integrate` including the output-time event handling of
 * src/integratorInstance.c:1029-1237.
 *
 * All pointers are plain host pointers unless named "device".  No torch types.
 * Every function returns 1 on success and 0 on failure (a SolverError is
 * pushed, codes 15XXXX) unless stated otherwise.  There is no CPU fallback:
 * without a CUDA device every launch entry fails with
 * SOLVER_ERROR_CUDA_UNAVAILABLE.
 */
#ifndef ODES_B200_BATCHINTEGRATOR_H
#define ODES_B200_BATCHINTEGRATOR_H
#include <stddef.h>
#include "sbmlsolver/integratorInstance.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct odesBatch odesBatch_t;

/* per-trajectory work counters, CVODE naming (cvode_statistics.txt of the reference) */
enum { ODES_STAT_NST = 0, ODES_STAT_NFE, ODES_STAT_NSETUPS, ODES_STAT_NJE, ODES_STAT_NNI,
       ODES_STAT_NCFN, ODES_STAT_NETF, ODES_NUM_STATS };

/* One-shot drop-in for the batch loop.  values: [nsets][nvary] row-major, the
   varySettings_t.params layout (src/sbmlsolver/odeSolver.h:53-65).  out: NULL or
   [nsets][nout+1][nvalues] (per set: time-major rows of all model values, as the
   reference stores them in cvodeResults_t).  status: NULL or [nsets] CVODE flags
   (0 ok, -4 too much work, -5 too much accuracy, -6 error test failures, -7
   convergence failures, ... numbering of cvode.h:712-725 of the vendored tree).
   Returns the number of failed sets, or -1 if nothing could be launched. */
int IntegratorInstance_integrateBatch(integratorInstance_t *ii, int nsets, int nvary,
                                      variableIndex_t *const *vi, const double *values,
                                      double *out, int *status);

/* Plan API (device-resident use; what the one-shot call is built from). */
odesBatch_t *ODESBatch_create(odeModel_t *om, cvodeSettings_t *opt, int nvary, const int *varyIndex,
                              int nstore, const int *storeIndex /* NULL: all values */, int device);
void ODESBatch_free(odesBatch_t *);
int ODESBatch_reserve(odesBatch_t *, int capacity);
int ODESBatch_getCapacity(const odesBatch_t *);
int ODESBatch_getNout(const odesBatch_t *);        /* number of stored rows = PrintStep+1 */
int ODESBatch_getNstore(const odesBatch_t *);
/* values after reset with the model's own parameters (what every trajectory starts from before its
   varied entries are set, src/cvodeData.c:156-190) and the initial trigger flags (:377-379) */
void ODESBatch_getBaseValues(const odesBatch_t *, double *values /* [nvalues] or NULL */, int *triggers /* [nevents] or NULL */);
const char *ODESBatch_getSource(const odesBatch_t *);   /* generated CUDA source */
const char *ODESBatch_getCompileLog(const odesBatch_t *);
int ODESBatch_getKernelInfo(const odesBatch_t *, int *regsPerThread, int *sharedBytesPerBlock,
                            int *localBytesPerThread, int *threadsPerBlock, int *blocksPerSM);

/* inputs: host row-major [nsets][nvary] -> device SoA [nvary][capacity] */
int ODESBatch_setValues(odesBatch_t *, int nsets, const double *values);
/* synthetic inputs generated on the device: value = base[j] * 2^u, u ~ U(lo[j], hi[j]),
   Philox4x32-10 keyed by seed, counter = (firstSet + set, j)   */
int ODESBatch_generateValues(odesBatch_t *, int nsets, unsigned long long seed, long long firstSet,
                             const double *base, const double *lo, const double *hi, int log10Scale);
/* the batch's inputs back on the host as rows [nsets][nvary] */
int ODESBatch_getValues(odesBatch_t *, int nsets, double *values);
double *ODESBatch_deviceValues(odesBatch_t *);     /* [nvary][capacity] */
double *ODESBatch_deviceResults(odesBatch_t *);    /* [capacity][nout+1][nstore] */

/* run: asynchronous launch on the plan's stream, then explicit synchronisation */
int ODESBatch_integrate(odesBatch_t *, int nsets);
int ODESBatch_synchronize(odesBatch_t *);
float ODESBatch_getLastKernelMs(const odesBatch_t *);   /* CUDA events on the launching stream */
int ODESBatch_getNumLaunches(const odesBatch_t *);
/* stopwatch on the launching stream: Start records an event, Stop records + synchronises a second one
   and returns the milliseconds in between (-1 on failure) */
int ODESBatch_timerStart(odesBatch_t *);
float ODESBatch_timerStop(odesBatch_t *);

/* outputs: device -> host.  results: [nsets][nout+1][nstore] (one contiguous time-major block per trajectory,
   the layout of the reference's cvodeResults_t per design point). */
int ODESBatch_getResults(odesBatch_t *, int nsets, double *results);
int ODESBatch_getTrajectory(odesBatch_t *, int set, double *traj /* [nout+1][nstore] */);
int ODESBatch_getStatus(odesBatch_t *, int nsets, int *status);
int ODESBatch_getStats(odesBatch_t *, int nsets, int *stats /* [ODES_NUM_STATS][nsets] */);
int ODESBatch_getEventLog(odesBatch_t *, int nsets, unsigned int *fired /* [nout+1][nsets] bitmasks */);
/* forward sensitivities (plans created with opt->Sensitivity = 1; parameters from CvodeSettings_setSensParams, default
   all constants): what the reference keeps in cvodeResults_t.sensitivity[i][j][k] (src/cvodeData.c:555-567, filled by
   IntegratorInstance_getForwardSens, src/sensSolver.c:108-140), here sens[set][k][j][i] = d y_i(t_k) / d p_j.
   Rows not reached are not-a-number like the state rows. */
int ODESBatch_getNsens(const odesBatch_t *);
int ODESBatch_getSensIndex(const odesBatch_t *, int is);     /* value[] index of sensitivity parameter `is` */
int ODESBatch_getDifferenceQuotients(const odesBatch_t *);    /* bit 0: difference-quotient sensitivity right-hand sides, bit 1: difference-quotient Jacobian */
int ODESBatch_getSensitivities(odesBatch_t *, int nsets, double *sens /* [nsets][nout+1][nsens][neq] */);
double *ODESBatch_deviceSensitivities(odesBatch_t *);

/* reaction fluxes of the stored rows (the reference: src/odeSolver.c:463-471, :519-531), evaluated by
   odes_flux_kernel over the device-resident results.  Plans of models with reactions whose om->m is set carry the
   emitted kinetic laws; the plan must store all model values (storeIndex == NULL). */
int ODESBatch_getNflux(const odesBatch_t *);                 /* number of reactions, 0: no flux function in this plan */
int ODESBatch_computeFluxes(odesBatch_t *, int nsets);       /* asynchronous, on the plan's stream, after ODESBatch_integrate */
int ODESBatch_getFluxes(odesBatch_t *, int nsets, double *flux /* [nsets][nout+1][nflux] */);
double *ODESBatch_deviceFluxes(odesBatch_t *);
float ODESBatch_getLastFluxMs(const odesBatch_t *);          /* CUDA events around the last flux launch */

/* end-to-end helper with host buffers (values [nsets][nvary] in, results [nsets][nout+1][nstore] and status out).
   chunk >= 0: launches of `chunk` sets (0: up to 2^22), the results of finished groups of sets are copied back while
   the kernel runs; chunk < 0: launches of -chunk sets, the copy of launch k-1 overlaps launch k */
int ODESBatch_runHost(odesBatch_t *, int nsets, const double *values, double *results, int *status,
                      int chunk);
/* ... with the forward sensitivities [nsets][nout+1][nsens][neq] streamed back as well (sens may be NULL) */
int ODESBatch_runHostSens(odesBatch_t *, int nsets, const double *values, double *results, double *sens, int *status,
                          int chunk);

/* ... and with the reaction fluxes [nsets][nout+1][nflux] (flux may be NULL) */
int ODESBatch_runHostEx(odesBatch_t *, int nsets, const double *values, double *results, double *sens, double *flux,
                        int *status, int chunk);

/* the same over all GPUs of the box: device k integrates the index range [k n / G, (k+1) n / G) of the caller's arrays
   (no collective; one plan and one host thread per device).  ndevices <= 0: all devices when every one gets at least
   2^15 sets (ODES_DEVICES overrides).  Returns the number of devices used, 0 on failure.  IntegratorInstance_integrateBatch
   and Model_odeSolverBatch[Flat] run through this entry. */
int ODESBatch_runHostAllDevices(odeModel_t *om, cvodeSettings_t *opt, int nvary, const int *varyIndex, int nstore,
                                const int *storeIndex /* NULL: all values */, int nsets, const double *values,
                                double *results, int *status, int ndevices);

int ODESBatch_runHostAllDevicesSens(odeModel_t *om, cvodeSettings_t *opt, int nvary, const int *varyIndex, int nstore,
                                    const int *storeIndex, int nsets, const double *values, double *results,
                                    double *sens /* NULL or [nsets][nout+1][nsens][neq] */, int *status, int ndevices);
int ODESBatch_runHostAllDevicesEx(odeModel_t *om, cvodeSettings_t *opt, int nvary, const int *varyIndex, int nstore,
                                  const int *storeIndex, int nsets, const double *values, double *results,
                                  double *sens, double *flux /* NULL or [nsets][nout+1][nreactions] */, int *status, int ndevices);
/* one-shot entry with sensitivities: sens NULL or [nsets][nout+1][nsens][neq], nsens as IntegratorInstance_getNsens(ii) */
int IntegratorInstance_integrateBatchSens(integratorInstance_t *ii, int nsets, int nvary, variableIndex_t *const *vi,
                                          const double *values, double *out, double *sens, int *status);

/* The one-shot entries (IntegratorInstance_integrateBatch, Model_odeSolverBatch[Flat], ODESBatch_runHostAllDevices*) keep
   their plans -- compiled module, loaded kernel, streams, device buffers of every GPU used -- on the odeModel_t, keyed by
   the generated kernel text; tolerances, step limit, time grid and base values are refreshed per call.  The cache goes
   with ODEModel_free; ODEModel_freeKernelCache releases it (and its device memory) earlier. */
void ODEModel_freeKernelCache(odeModel_t *om);
/* work counters: [0] NVRTC compilations, [1] cubin-cache loads, [2] module loads, [3] device / pinned allocations,
   [4] plans created, [5] plan-cache hits, [6] kernel launches, [7] host registrations */
void ODES_getCounters(long long *counters /* [8] */);
/* pinned host buffers for callers that want full-speed host<->device copies */
void *ODES_hostAlloc(size_t bytes);
void ODES_hostFree(void *);

/* device utilities */
int ODES_deviceCount(void);
int ODES_deviceName(int device, char *buf, int len);
/* FP64 FMA micro-benchmark: returns measured TFLOP/s (2 flop per FMA), 0 on failure */
double ODES_measureFP64Peak(int device, int repeats);

#ifdef __cplusplus
}
#endif
#endif
