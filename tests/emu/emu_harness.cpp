/*
 * emu_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into libodes_b200.so).
 * Compiles the generated model section + bdf_kernel.cuh for the HOST so the
 * per-trajectory program of the CUDA kernel can be checked against the oracle
 * in the CPU-only test tier (the GPU tier checks the real kernel).  Usage:
 *   g++ -O1 -ffp-contract=off -shared -fPIC -DODES_MODEL_SOURCE='"full_tu.cu"' emu_harness.cpp   (full_tu.cu = ODESBatch_getSource())
 */
#include <cmath>
#include <cstddef>
#include <cstring>
using namespace std;
#define ODES_HOST_EMULATION 1
#define __device__
#define __forceinline__ inline
#define __constant__ static
#define __restrict__
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
#include ODES_MODEL_SOURCE


extern "C" int emu_run(const double *base, const int *trig0, const double *vary, double *out, int *status, int *stats,
                       unsigned int *fired, const double *tpoints, unsigned long long stride, int nsets, int nout,
                       int mxstep, double reltol, double abstol, double *sens)
{
  for (int i = 0; i < NVAL; i++) c_base[i] = base[i];
  for (int i = 0; i < NEVENTS; i++) c_trigger0[i] = trig0[i];
  OdesArgs a;
  a.vary = vary; a.out = out; a.status = status; a.stats = stats; a.fired = fired; a.tpoints = tpoints;
  a.sens = sens;
  a.groupDone = 0; a.hostFlags = 0; a.groupShift = 0; a.pad2 = 0;
  a.stride = stride; a.nsets = nsets; a.nout = nout; a.mxstep = mxstep; a.pad = 0; a.reltol = reltol; a.abstol = abstol;
  unsigned long long counter = 0;
  a.counter = &counter;
#if ODES_WARP
  a.scratch = 0;
  static WarpMem wm;
  integrate_warp(a, wm);                /* one emulated warp (32-wide lane values) pulls every set index in turn */
#else
  static double sm[ODES_SMEM_WORDS + 1];
  static double gscr[ODES_GLOBAL_WORDS + 1];
  a.scratch = gscr;
  La la; la.sm = sm; la.gm = gscr; la.gs = 1u; la.taddr = 0u;
  integrate_lanes(a, sm, la);           /* one emulated lane pulls every set index in turn */
#endif
  return 0;
}
