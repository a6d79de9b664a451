/*
 * batch.cu -- the device side of the C ABI: run-time compilation of the model
 * kernel (NVRTC -> sm_100a cubin -> driver module), device buffers in
 * structure-of-arrays layout, launches, result gathering.
 *
 * Reference seams this file replaces:
 *   src/compiler.c:428-509   Compiler_compile / CompiledCode_getFunction / CompiledCode_free
 *                            (g++ -shared + dlopen)  -> NVRTC + cuModuleLoadData
 *   src/odeModel.c:3191-3299 ODEModel_compileCVODEFunctions (assemble TU, compile, look up ode_f...)
 *   src/odeSolver.c:313-330  the serial design-point loop -> ODESBatch_integrate (one launch)
 * No CPU fallback exists: every entry that needs the device fails with
 * SOLVER_ERROR_CUDA_UNAVAILABLE when there is none.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvrtc.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sched.h>
#include <string>
#include <thread>
#include <vector>
#include "sbmlsolver/batchIntegrator.h"
#include "sbmlsolver/odeSolver.h"
#include "sbmlsolver/processAST.h"
#include "sbmlsolver/solverError.h"

extern "C" const char odes_bdf_kernel_source[];     /* bdf_kernel.cuh embedded as text (kernel_text.c) */

/* argument block of the kernel; must match struct OdesArgs in bdf_kernel.cuh */
struct OdesArgs {
  const double *vary; double *out; int *status; int *stats; unsigned int *fired; const double *tpoints;
  unsigned long long *counter; double *scratch; unsigned long long stride; int nsets, nout, mxstep, pad; double reltol, abstol;
  unsigned int *groupDone; unsigned int *hostFlags; int groupShift, pad2;
  double *sens;
};

/* --------------------------------------------------------------- driver API */
static struct {
  int ready;
  CUresult (*ModuleLoadData)(CUmodule *, const void *);
  CUresult (*ModuleUnload)(CUmodule);
  CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *);
  CUresult (*ModuleGetGlobal)(CUdeviceptr *, size_t *, CUmodule, const char *);
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int);
  CUresult (*FuncGetAttribute)(int *, CUfunction_attribute, CUfunction);
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **, void **);
  CUresult (*MemcpyHtoD)(CUdeviceptr, const void *, size_t);
  CUresult (*MemcpyDtoH)(void *, CUdeviceptr, size_t);
  CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int *, CUfunction, int, size_t);
} drv;

#define CUDA_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_CUDA_CALL_FAILED, "%s failed: %s", #call, cudaGetErrorString(e_)); return 0; } } while (0)
#define DRV_OK(call) do { CUresult r_ = (call); if (r_ != CUDA_SUCCESS) { \
    SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_CUDA_CALL_FAILED, "%s failed with CUresult %d", #call, (int)r_); return 0; } } while (0)
/* Host-to-device uploads from PAGEABLE caller memory with the synchronous copy calls return once the data is staged, not once
   the DMA has landed, and the kernels run on non-blocking streams that do not order against the NULL stream: wait for the
   NULL stream after every such group of uploads, before anything can be launched that reads them */
#define UPLOAD_FENCE() CUDA_OK(cudaStreamSynchronize(cudaStreamLegacy))

static int load_driver(void)
{
  if (drv.ready) return 1;
#define GET(field, name) do { void *p = NULL; cudaDriverEntryPointQueryResult q; \
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || !p) { \
      SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_CUDA_UNAVAILABLE, "CUDA driver entry point %s unavailable", name); return 0; } \
    *(void **)(&drv.field) = p; } while (0)
  GET(ModuleLoadData, "cuModuleLoadData");
  GET(ModuleUnload, "cuModuleUnload");
  GET(ModuleGetFunction, "cuModuleGetFunction");
  GET(ModuleGetGlobal, "cuModuleGetGlobal");
  GET(FuncSetAttribute, "cuFuncSetAttribute");
  GET(FuncGetAttribute, "cuFuncGetAttribute");
  GET(LaunchKernel, "cuLaunchKernel");
  GET(MemcpyHtoD, "cuMemcpyHtoD");
  GET(MemcpyDtoH, "cuMemcpyDtoH");
  GET(OccupancyMaxActiveBlocksPerMultiprocessor, "cuOccupancyMaxActiveBlocksPerMultiprocessor");
#undef GET
  drv.ready = 1;
  return 1;
}

extern "C" int ODES_deviceCount(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
extern "C" int ODES_deviceName(int device, char *buf, int len)
{
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) { cudaGetLastError(); return 0; }
  snprintf(buf, (size_t)len, "%s", p.name);
  return 1;
}
static int need_device(int device)
{
  if (ODES_deviceCount() <= device) {
    SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_CUDA_UNAVAILABLE,
                      "no CUDA device %d available: the batched integrator has no CPU fallback", device);
    return 0;
  }
  CUDA_OK(cudaSetDevice(device));
  CUDA_OK(cudaFree(0));                       /* primary context */
  return load_driver();
}

/* ----------------------------------------------------------- code-gen seam */
static unsigned long long fnv1a(const char *s, unsigned long long h)
{
  for (; *s; s++) { h ^= (unsigned char)*s; h *= 1099511628211ULL; }
  return h;
}
static std::string cache_dir(void)
{
  const char *env = getenv("ODES_B200_CACHE");
  if (env && *env) return env;
  Dl_info info;
  if (dladdr((void *)&cache_dir, &info) && info.dli_fname) {
    std::string p = info.dli_fname;
    size_t k = p.rfind('/');
    return (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/_cache";
  }
  return "/tmp/odes_b200_cache";
}

/* work counters of the library (ODES_getCounters): how often the expensive, once-per-plan operations ran.  The one-shot
   entries keep their plans and device buffers on the odeModel_t, so a second call with the same model, settings
   and varied set must not move any of the first four. */
#include <atomic>
#include <condition_variable>
#include <mutex>
enum { CNT_NVRTC = 0, CNT_CUBIN_LOAD, CNT_MODULE_LOAD, CNT_DEVICE_ALLOC, CNT_PLAN_CREATE, CNT_PLAN_HIT, CNT_LAUNCH, CNT_HOST_REGISTER, CNT_NUM };
static std::atomic<long long> g_counters[CNT_NUM];
extern "C" void ODES_getCounters(long long *out /* [8]: NVRTC compilations, cubin-cache loads, module loads, device/pinned allocations,
                                                  plans created, plan-cache hits, kernel launches, host registrations */)
{ for (int i = 0; i < CNT_NUM; i++) out[i] = g_counters[i].load(); }
static cudaError_t counted_malloc(void **p, size_t bytes) { g_counters[CNT_DEVICE_ALLOC]++; return cudaMalloc(p, bytes); }
template <class T> static cudaError_t counted_malloc(T **p, size_t bytes) { return counted_malloc((void **)p, bytes); }
#define cudaMalloc counted_malloc

static int env_int(const char *name, int dflt) { const char *v = getenv(name); return (v && *v) ? atoi(v) : dflt; }

/* Compiles a CUDA translation unit for sm_100a.  Same three-call shape as the reference
   (Compiler_compile / CompiledCode_getFunction / CompiledCode_free); extra NVRTC options are
   passed in a leading "//!opts: ..." line of the source. */
extern "C" compiled_code_t *Compiler_compile(const char *sourceCode)
{
  std::vector<std::string> opts;
  opts.push_back("--gpu-architecture=sm_100a");
  opts.push_back("-lineinfo");
  opts.push_back("--std=c++17");
  opts.push_back("-default-device");
  if (strncmp(sourceCode, "//!opts:", 8) == 0) {
    const char *e = strchr(sourceCode, '\n');
    std::string line(sourceCode + 8, e ? (size_t)(e - sourceCode - 8) : strlen(sourceCode + 8));
    size_t pos = 0;
    while (pos < line.size()) {
      while (pos < line.size() && line[pos] == ' ') pos++;
      size_t end = line.find(' ', pos);
      if (end == std::string::npos) end = line.size();
      if (end > pos) opts.push_back(line.substr(pos, end - pos));
      pos = end;
    }
  }
  compiled_code_t *code = (compiled_code_t *)calloc(1, sizeof *code);
  code->source = strdup(sourceCode);

  int major = 0, minor = 0;
  nvrtcVersion(&major, &minor);
  unsigned long long h = fnv1a(sourceCode, 1469598103934665603ULL);
  for (size_t i = 0; i < opts.size(); i++) h = fnv1a(opts[i].c_str(), h);
  h = h * 1099511628211ULL + (unsigned long long)(major * 100 + minor);
  char name[64]; snprintf(name, sizeof name, "/%016llx.cubin", h);
  std::string dir = cache_dir(), path = dir + name;

  if (!env_int("ODES_B200_NOCACHE", 0)) {
    FILE *f = fopen(path.c_str(), "rb");
    if (f) {
      fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
      if (n > 0) {
        code->cubin = malloc((size_t)n);
        if (fread(code->cubin, 1, (size_t)n, f) == (size_t)n) { g_counters[CNT_CUBIN_LOAD]++; code->cubinSize = (unsigned long)n; code->log = strdup("(loaded from cache)"); fclose(f); return code; }
        free(code->cubin); code->cubin = NULL;
      }
      fclose(f);
    }
  }

  nvrtcProgram prog;
  if (nvrtcCreateProgram(&prog, sourceCode, "odes_model.cu", 0, NULL, NULL) != NVRTC_SUCCESS) {
    SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_COMPILATION_FAILED, "nvrtcCreateProgram failed");
    CompiledCode_free(code); return NULL;
  }
  std::vector<const char *> copts;
  for (size_t i = 0; i < opts.size(); i++) copts.push_back(opts[i].c_str());
  g_counters[CNT_NVRTC]++;
  nvrtcResult res = nvrtcCompileProgram(prog, (int)copts.size(), copts.data());
  size_t logSize = 0; nvrtcGetProgramLogSize(prog, &logSize);
  code->log = (char *)calloc(logSize + 1, 1);
  if (logSize > 1) nvrtcGetProgramLog(prog, code->log);
  if (res != NVRTC_SUCCESS) {
    SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_COMPILATION_FAILED, "NVRTC compilation failed (%s):\n%.1500s", nvrtcGetErrorString(res), code->log);
    nvrtcDestroyProgram(&prog); CompiledCode_free(code); return NULL;
  }
  size_t n = 0; nvrtcGetCUBINSize(prog, &n);
  code->cubin = malloc(n); code->cubinSize = (unsigned long)n;
  nvrtcGetCUBIN(prog, (char *)code->cubin);
  nvrtcDestroyProgram(&prog);

  if (!env_int("ODES_B200_NOCACHE", 0)) {
    mkdir(dir.c_str(), 0777);
    std::string tmp = path + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (f) { fwrite(code->cubin, 1, n, f); fclose(f); rename(tmp.c_str(), path.c_str()); }
  }
  return code;
}

/* Loads the module on the current device at first use and returns the kernel (CUfunction). */
extern "C" void *CompiledCode_getFunction(compiled_code_t *code, const char *symbol)
{
  CUfunction fn = NULL;
  if (!code || !code->cubin) return NULL;
  if (!load_driver()) return NULL;
  if (!code->module) {
    CUmodule m;
    g_counters[CNT_MODULE_LOAD]++;
    CUresult r = drv.ModuleLoadData(&m, code->cubin);
    if (r != CUDA_SUCCESS) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_DLL_LOAD_FAILED, "cuModuleLoadData failed with CUresult %d", (int)r); return NULL; }
    code->module = m;
  }
  if (drv.ModuleGetFunction(&fn, (CUmodule)code->module, symbol) != CUDA_SUCCESS) {
    SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_GET_PROC_ADDRESS_FAILED, "kernel %s not found in compiled module", symbol);
    return NULL;
  }
  return (void *)fn;
}
extern "C" void CompiledCode_free(compiled_code_t *code)
{
  if (!code) return;
  if (code->module && drv.ready) drv.ModuleUnload((CUmodule)code->module);
  free(code->cubin); free(code->log); free(code->source); free(code);
}

/* ------------------------------------------------------------- static kernels */
/* Philox4x32-10 (Salmon et al. 2011), counter = (set, j), key = seed */
__device__ __forceinline__ void philox_round(unsigned int (&c)[4], unsigned int k0, unsigned int k1)
{
  const unsigned int hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const unsigned int hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const unsigned int n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ double philox_uniform(unsigned long long seed, unsigned long long set, unsigned int j)
{
  unsigned int c[4] = {(unsigned int)set, (unsigned int)(set >> 32), j, 0u};
  unsigned int k0 = (unsigned int)seed, k1 = (unsigned int)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
  const unsigned long long bits = ((unsigned long long)c[0] << 32) | c[1];
  return (double)(bits >> 11) * (1.0 / 9007199254740992.0);          /* [0,1) with 53 bits */
}
__global__ void odes_generate_values_kernel(double *vary, size_t stride, int nsets, int nvary, unsigned long long seed,
                                            long long firstSet, const double *base, const double *lo, const double *hi, int log10Scale)
{
  const size_t set = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (set >= (size_t)nsets) return;
  for (int j = 0; j < nvary; j++) {
    const double u = lo[j] + (hi[j] - lo[j]) * philox_uniform(seed, (unsigned long long)(firstSet + (long long)set), (unsigned int)j);
    vary[(size_t)j * stride + set] = log10Scale ? base[j] * pow(10.0, u) : base[j] * exp2(u);
  }
}
/* host layout [nsets][nvary] (varySettings rows) -> SoA [nvary][stride] */
__global__ void odes_rows_to_soa_kernel(const double *rows, double *vary, size_t stride, int nsets, int nvary)
{
  const size_t set = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (set >= (size_t)nsets) return;
  for (int j = 0; j < nvary; j++) vary[(size_t)j * stride + set] = rows[set * (size_t)nvary + j];
}
/* FP64 FMA throughput probe: 8 independent chains per thread */
__global__ void odes_fp64_peak_kernel(double *out, int iters)
{
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
extern "C" double ODES_measureFP64Peak(int device, int repeats)
{
  if (!need_device(device)) return 0.0;
  cudaDeviceProp p; if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return 0.0;
  const int threads = 256, blocks = p.multiProcessorCount * 8, iters = 1 << 16;
  double *out = NULL; cudaEvent_t e0, e1; float best = 1e30f;
  if (cudaMalloc(&out, (size_t)threads * blocks * sizeof(double)) != cudaSuccess) return 0.0;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int r = 0; r < repeats + 1; r++) {
    cudaEventRecord(e0);
    odes_fp64_peak_kernel<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  const double flop = 2.0 * 8.0 * (double)iters * (double)threads * (double)blocks;
  return flop / (best * 1e-3) / 1e12;
}

/* ------------------------------------------------------------------ the plan */
struct odesBatch {
  odeModel_t *om; cvodeSettings_t *opt;
  int device, nvary, nstore, nout, nvalues;
  std::vector<int> vary, store;
  compiled_code_t *code; CUfunction kernel;
  int block, minblocks, matLocal; size_t smemBytes;
  int capacity;
  double *d_vary, *d_out, *d_tpoints, *d_rows; int *d_status, *d_stats; unsigned int *d_fired; unsigned long long *d_counter; int numSMs, blocksPerSM, counterSlot; double *d_scratch; int warpMode, laMode, laWords, tmemCols, maxBlocksPerSM;
  size_t rowsCap;
  unsigned int *d_groupDone, *h_flags, *d_flags; int groupCap;   /* completion signalling of the host pipeline */
  cudaStream_t stream, copyIn, copyOut; cudaEvent_t e0, e1, t0, t1, evIn[2], evK[2], evOut[2];
  float lastMs; int launches;
  std::vector<double> base; std::vector<int> trigger0;
  int sensFromRow0;                               /* initial sensitivities supplied by the host (single-instance restarts) */
  int nflux; CUfunction fluxKernel; double *d_flux; size_t fluxCap; float lastFluxMs; cudaEvent_t f0, f1;   /* reaction fluxes of the stored rows: [capacity][nout+1][nflux] */
  odeSense_t *os; int nsens; double *d_sens;      /* forward sensitivities (opt->Sensitivity): [capacity][nout+1][nsens][neq] */
  int noEvents;                                   /* launches skip the emitted event function (single instance without event processing) */
  int rootFind;                                   /* StepResolution < 0: events located inside the steps (both kernels) */
  int stepRes;                                    /* cvodeSettings_t.StepResolution: solver / event grid points per print step */
  /* pinned staging ring for results that go to PAGEABLE caller memory (the reference-shaped entries hand in malloc'd arrays) */
  char *stage; size_t stageSlotBytes; int stageSlots; cudaEvent_t stageEv[8];
};

/* values after CvodeData_initialize with the model's own parameters (uniform for the whole batch) */
static int compute_base(odeModel_t *om, cvodeSettings_t *opt, std::vector<double> &base, std::vector<int> &trig)
{
  cvodeData_t *data = CvodeData_create(om);
  int store = opt->StoreResults;
  if (!data) return 0;
  opt->StoreResults = 0;
  CvodeData_initialize(data, opt, om, 0);
  opt->StoreResults = store;
  base.assign(data->value, data->value + data->nvalues);
  trig.assign(data->trigger, data->trigger + om->nevents);
  CvodeData_free(data);
  return 1;
}

extern "C" void ODESBatch_free(odesBatch_t *b)
{
  if (!b) return;
  if (ODES_deviceCount() > b->device) {
    cudaSetDevice(b->device);
    cudaFree(b->d_vary); cudaFree(b->d_out); cudaFree(b->d_tpoints); cudaFree(b->d_rows);
    cudaFree(b->d_status); cudaFree(b->d_stats); cudaFree(b->d_fired); cudaFree(b->d_counter); cudaFree(b->d_scratch); cudaFree(b->d_groupDone); if (b->h_flags) cudaFreeHost(b->h_flags);
    cudaFree(b->d_sens); cudaFree(b->d_flux);
    if (b->stage) cudaFreeHost(b->stage);
    for (int i = 0; i < 8; i++) if (b->stageEv[i]) cudaEventDestroy(b->stageEv[i]);
    if (b->f0) cudaEventDestroy(b->f0);
    if (b->f1) cudaEventDestroy(b->f1);
    if (b->stream) cudaStreamDestroy(b->stream);
    if (b->copyIn) cudaStreamDestroy(b->copyIn);
    if (b->copyOut) cudaStreamDestroy(b->copyOut);
    if (b->e0) cudaEventDestroy(b->e0);
    if (b->e1) cudaEventDestroy(b->e1);
    if (b->t0) cudaEventDestroy(b->t0);
    if (b->t1) cudaEventDestroy(b->t1);
    for (int i = 0; i < 2; i++) { if (b->evIn[i]) cudaEventDestroy(b->evIn[i]); if (b->evK[i]) cudaEventDestroy(b->evK[i]); if (b->evOut[i]) cudaEventDestroy(b->evOut[i]); }
  }
  CompiledCode_free(b->code);
  ODESense_free(b->os);
  delete b;
}

/* reaction fluxes: when the plan stores every model value and the ODE model still knows its SBML model,
   the kinetic laws are emitted as flux_row next to the model functions; odes_flux_kernel (end of the kernel text) maps
   the stored rows.  ODES_FLUX=0 leaves it out. */
static std::string flux_section(odesBatch_t *b)
{
  b->nflux = 0;
  if (!b->om->m || !env_int("ODES_FLUX", 1)) return "";
  if (b->nstore != b->nvalues) return "";
  for (int j = 0; j < b->nstore; j++) if (b->store[(size_t)j] != j) return "";
  ASTNode_t **kls = NULL;
  const int n = SBMLResults_fluxASTs(b->om->m, b->om, &kls);
  std::string text;
  if (n > 0) { char *t = ODEModel_generateCUDAFluxSection(n, kls); text = t; free(t); b->nflux = n; }
  for (int r = 0; r < n; r++) if (kls[r]) ASTNode_free(kls[r]);
  free(kls);
  return text;
}

static std::string build_source(odesBatch_t *b)
{
  char *model = ODEModel_generateCUDASourceSens(b->om, b->opt, b->nvary, b->vary.data(), b->nstore, b->store.data(), env_int("ODES_FOLD_CONSTANTS", 1) ? b->base.data() : NULL, b->os);
  const int neq = b->om->neq;
  const int usejac = b->opt->UseJacobian && b->om->jacobian;
  int npv = 0;
  { const char *q = strstr(model, "#define NPV "); if (q) npv = atoi(q + 12); }
  /* per-lane words (bdf_kernel.cuh): shared memory holds zn[6], ewt, acor, y, ftemp, tau/tq/l; the linear-algebra
     words (M with columns padded to NP rows, saved Jacobian, per-trajectory constants; all in chunks of 8) go to
     tensor memory when they fit its 512 columns (2 columns per double), else shared memory, else an L2-resident
     global scratch */
  int chipWords = 10 * neq + 19;
  const int np = 8 * ((neq + 7) / 8);
  const int nsjw = usejac ? b->om->sparsesize : neq * np;
  b->laWords = neq * np + 8 * ((nsjw + 7) / 8) + 8 * ((npv + 7) / 8);
  int tmemCols = 32; while (tmemCols < 2 * b->laWords) tmemCols *= 2;
  /* small models: when the saved Jacobian's non-zeros and the constants TOGETHER need fewer chunks than each padded on its
     own, and that halves the tensor-memory columns of a lane, they share their chunks (ODES_LA_PACKED in the kernel text):
     Goodwin 72 -> 64 words = 128 instead of 256 columns, which admits a third resident block per SM */
  int packed = 0;
  if (usejac) {
    const int pk = neq * np + 8 * ((nsjw + npv + 7) / 8);
    int pkCols = 32; while (pkCols < 2 * pk) pkCols *= 2;
    if (env_int("ODES_LA_PACKED", 1) && pkCols < tmemCols && pkCols <= 512) { packed = 1; b->laWords = pk; tmemCols = pkCols; }
  }
  int la = env_int("ODES_LA", -1);
  if (la < 0) la = (tmemCols <= 512 && neq > 0) ? 2 : ((size_t)(chipWords + b->laWords) * 8 * 64 <= 112 * 1024 ? 0 : 1);
  if (la == 2 && tmemCols > 512) la = 1;
  /* larger networks: one trajectory per WARP (bdf_warp_kernel.cuh) -- state in registers, row of the Newton matrix
     per lane; chosen when a lane's linear-algebra words no longer fit tensor memory */
  int warp = env_int("ODES_WARP", -1);
  /* crossover measured with tools/gpu_layout_crossover.py (chain models, 10^5 sets): n = 16 lane 1.09 M vs warp 0.69 M,
     n = 20 0.51 M vs 0.56 M, n = 24 0.26 M vs 0.48 M, n = 32 0.11 M vs 0.38 M trajectories/s; huang96 (n = 22) 82 k vs 174 k */
  /* without a Jacobian (CVDenseDQJac, one right-hand side per column): huang96 lane 62 k vs warp 296 k, MAPK lane 1.91 M vs
     warp 0.91 M (tools/gpu_dq_layout.py) -- the same crossover */
  if (warp < 0) warp = (la != 2 && neq >= 19 && neq <= 32) ? 1 : 0;
  if (b->nsens > 0) warp = 1;                   /* the simultaneous corrector lives in the warp-per-trajectory kernel */
  if (b->opt->CvodeMethod == 1 || b->opt->IterMethod == 1) warp = 1;       /* and so do Adams-Moulton and functional iteration */
  if (warp && !(neq >= 1 && neq <= 32)) warp = 0;
  b->warpMode = warp;
  if (warp) {
    const int wpb = env_int("ODES_WARPS_PER_BLOCK", 4);
    b->laMode = 0; b->tmemCols = 0; b->block = 32 * wpb; b->matLocal = 0; b->smemBytes = 0;
    /* measured on huang96 (B200): 8 warps per SM 113 k, 12 warps 145 k, 16 warps (128 registers, spills) 136 k trajectories/s;
       after the solver state left local memory (136 registers unconstrained): 12 warps 227 k, 16 warps (128 registers) 267 k */
    /* with sensitivities a lane carries 10 more vector elements per sensitivity: fewer resident warps, more registers */
    int slabs = 1;                            /* register slabs of the packed layout (bdf_warp_kernel.cuh: W_S) */
    if (b->nsens > 0) { const int nv = 1 + b->nsens; int g = (usejac && b->os->sensitivity) ? 32 / neq : 1; if (g < 1) g = 1; if (g > nv) g = nv; slabs = (nv + g - 1) / g; }   /* difference quotients: one vector per slab */
    b->minblocks = env_int("ODES_MINBLOCKS", wpb != 4 ? 1 : b->opt->CvodeMethod == 1 ? 3 : slabs == 1 ? 4 : slabs == 2 ? 3 : slabs <= 4 ? 2 : 1); b->maxBlocksPerSM = 32;
    char headw[640];
    snprintf(headw, sizeof headw, "//!opts: --fmad=%s -DODES_STEPRES=%d -DODES_ROOTFIND=%d -DODES_WARP=1 -DODES_ADAMS=%d -DODES_FUNCTIONAL=%d%s -DODES_WARPS_PER_BLOCK=%d -DODES_BLOCK=%d -DODES_MINBLOCKS=%d -DODES_WM_LAUNDER=%d -DODES_WDIV=%d -DODES_WLOCK=%d -DODES_WJE=%d -DODES_WDAG=%d%s\n",
             env_int("ODES_FMAD", 0) ? "true" : "false", b->stepRes, b->rootFind, b->opt->CvodeMethod == 1 ? 1 : 0, b->opt->IterMethod == 1 ? 1 : 0, env_int("ODES_WJAC_INLINE", -1) >= 0 ? (env_int("ODES_WJAC_INLINE", 0) ? " -DODES_WJAC_INLINE=1" : " -DODES_WJAC_INLINE=0") : "", wpb, b->block, b->minblocks, env_int("ODES_WM_LAUNDER", 0), env_int("ODES_WDIV", 2), env_int("ODES_WLOCK", (b->nsens > 0 && slabs > 4) ? 1 : 0), env_int("ODES_WJE", -1), env_int("ODES_WDAG", -1),
             env_int("ODES_MAXREG", 0) ? (std::string(" --maxrregcount=") + std::to_string(env_int("ODES_MAXREG", 0))).c_str() : "");
    std::string srcw = headw;
    srcw += model;
    srcw += flux_section(b);
    srcw += odes_bdf_kernel_source;
    free(model);
    return srcw;
  }
  b->laMode = la; b->tmemCols = la == 2 ? tmemCols : 0;
  b->block = env_int("ODES_BLOCK", la == 2 ? 128 : 64);
  /* blocks of more than 4 warps: warps w and w+4 share TMEM lanes and take separate column ranges */
  if (la == 2 && b->block > 128) { tmemCols = 32; while (tmemCols < 2 * b->laWords * ((b->block + 127) / 128)) tmemCols *= 2; if (tmemCols > 512) { b->block = 128; tmemCols = 32; while (tmemCols < 2 * b->laWords) tmemCols *= 2; } b->tmemCols = tmemCols; }
  /* With the analytic Jacobian f can share the words of y (bdf_kernel.cuh: ODES_ALIAS_FY), one vector less per lane.  It
     serialises a few loads behind stores (Goodwin, which gains no block from it: 164.5 -> 171.0 ms), so it is used where it
     buys a resident block: a six-equation model drops from 80.9 to 74.8 KB per block and fits three (repressilator, 10^6 sets x
     1000 output points: 575 -> 479 ms) */
  int alias = 0;
  if (usejac && la == 2 && env_int("ODES_ALIAS_FY", 1)) {
    const int blk = env_int("ODES_BLOCK", 128);
    const int fitU = (int)((228 * 1024) / ((size_t)chipWords * 8 * (size_t)blk + 1024)), fitA = (int)((228 * 1024) / ((size_t)(chipWords - neq) * 8 * (size_t)blk + 1024));
    const int lim = 512 / tmemCols, wantBlocks = env_int("ODES_THREADS_PER_SM", neq <= 3 ? 512 : neq <= 6 ? 384 : 256) / blk;
    const int useU = fitU < lim ? (fitU < wantBlocks ? fitU : wantBlocks) : (lim < wantBlocks ? lim : wantBlocks);
    const int useA = fitA < lim ? (fitA < wantBlocks ? fitA : wantBlocks) : (lim < wantBlocks ? lim : wantBlocks);
    if (useA > useU || env_int("ODES_ALIAS_FY", 1) > 1) { alias = 1; chipWords -= neq; }
  }
  const int words = chipWords + (la == 0 ? b->laWords : 0);
  b->matLocal = env_int("ODES_MAT_LOCAL", 0);
  while (!b->matLocal && (size_t)words * 8 * (size_t)b->block + 1024 > 227 * 1024 && b->block > 32) b->block /= 2;
  if ((size_t)words * 8 * (size_t)b->block + 1024 > 227 * 1024) b->matLocal = 1;
  b->smemBytes = b->matLocal ? 0 : (size_t)words * 8 * (size_t)b->block;
  int fit = b->smemBytes ? (int)((228 * 1024) / (b->smemBytes + 1024)) : 8;
  if (la == 2 && fit > 512 / tmemCols) fit = 512 / tmemCols;            /* resident blocks share the 512 TMEM columns */
  /* resident threads per SM asked for: 256 (two blocks of 128) -- a MAPK-sized lane fills the SM's shared and tensor memory
     with that; small systems whose lanes are half the size get a third block (384: Goodwin 5.13 -> 6.08 M trajectories/s at 10^6 sets) or,
     up to three equations, a fourth (512: events model 6.44 -> 7.40 M; profiles/r02_small_models.jsonl) */
  int want = env_int("ODES_THREADS_PER_SM", la != 2 ? 256 : neq <= 3 ? 512 : neq <= 6 ? 384 : 256) / b->block;
  b->minblocks = env_int("ODES_MINBLOCKS", fit < want ? (fit < 1 ? 1 : fit) : (want < 1 ? 1 : want));
  b->maxBlocksPerSM = (la == 2) ? 512 / tmemCols : 32;
  char head[640];
  snprintf(head, sizeof head, "//!opts: --fmad=%s -DODES_STEPRES=%d -DODES_ROOTFIND=%d -DODES_BLOCK=%d -DODES_MINBLOCKS=%d -DODES_MAT_LOCAL=%d -DODES_SETUP_MIN=%d -DODES_SETUP_WAIT=%d -DODES_REFILL_MIN=%d -DODES_REFILL_WAIT=%d -DODES_COL_UNROLL=%d -DODES_LA=%d -DODES_LA_PACKED=%d -DODES_ALIAS_FY=%d -DODES_TMEM_COLS=%d -DODES_LOCKSTEP=%d -DODES_GESL_UNROLL=%d -DODES_ALIGN_MASK=%d%s\n",
           env_int("ODES_FMAD", 0) ? "true" : "false", b->stepRes, b->rootFind, b->block, b->minblocks, b->matLocal,
           env_int("ODES_SETUP_MIN", 10), env_int("ODES_SETUP_WAIT", 3), env_int("ODES_REFILL_MIN", 3), env_int("ODES_REFILL_WAIT", 20), env_int("ODES_COL_UNROLL", 2), la, packed, alias, b->tmemCols, env_int("ODES_LOCKSTEP", 2), env_int("ODES_GESL_UNROLL", 4), env_int("ODES_ALIGN_MASK", 7),
           env_int("ODES_MAXREG", 0) ? (std::string(" --maxrregcount=") + std::to_string(env_int("ODES_MAXREG", 0))).c_str() : "");
  std::string src = head;
  /* development knob: further -D options for the kernel text (tools/gpu_sweep.py), e.g. ODES_EXTRA_DEFS="-DODES_LU_FORM=1" */
  if (getenv("ODES_EXTRA_DEFS") && *getenv("ODES_EXTRA_DEFS")) { src.erase(src.size() - 1); src += " "; src += getenv("ODES_EXTRA_DEFS"); src += "\n"; }
  src += model;
  src += flux_section(b);
  src += odes_bdf_kernel_source;
  free(model);
  return src;
}

static odesBatch_t *batch_prepare(odeModel_t *om, cvodeSettings_t *opt, int nvary, const int *varyIndex, int nstore, const int *storeIndex, int device, std::string &src);
extern "C" odesBatch_t *ODESBatch_create(odeModel_t *om, cvodeSettings_t *opt, int nvary, const int *varyIndex,
                                          int nstore, const int *storeIndex, int device)
{
  /* code generation + compilation need no device (NVRTC targets sm_100a explicitly) */
  std::string src;
  odesBatch_t *b = batch_prepare(om, opt, nvary, varyIndex, nstore, storeIndex, device, src);
  if (!b) return NULL;
  b->code = Compiler_compile(src.c_str());
  if (!b->code) { ODESBatch_free(b); return NULL; }
  g_counters[CNT_PLAN_CREATE]++;
  return b;
}
/* everything of a plan up to the generated kernel text (no compilation, no device) */
static odesBatch_t *batch_prepare(odeModel_t *om, cvodeSettings_t *opt, int nvary, const int *varyIndex,
                                  int nstore, const int *storeIndex, int device, std::string &src)
{
  if (!om || !opt) return NULL;
  if (om->hasCycle) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_ODE_MODEL_CYCLIC_DEPENDENCY_IN_RULES, "model rules contain a cycle"); return NULL; }
  if (opt->Indefinitely || !opt->TimePoints) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "batched integration needs a finite output grid"); return NULL; }
  if ((opt->CvodeMethod != 0 && opt->CvodeMethod != 1) || (opt->IterMethod != 0 && opt->IterMethod != 1)) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "CvodeMethod must be 0 (BDF) or 1 (Adams-Moulton), IterMethod 0 (Newton) or 1 (functional)"); return NULL; }
  if ((opt->CvodeMethod == 1 || opt->IterMethod == 1) && (opt->Sensitivity || om->neq < 1 || om->neq > 32 || !opt->UseJacobian)) {
    SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "Adams-Moulton and functional iteration run in the warp-per-trajectory kernel: 1..32 equations, analytic Jacobian available, no sensitivities");
    return NULL;
  }
  if (om->nevents > 32) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "at most 32 events are supported"); return NULL; }
  /* reference src/cvodeSolver.c:971-975: with DetectNegState the right-hand side reports a RECOVERABLE failure on a negative
     state and CVODE (>= 2.4) repeats the step with a quarter of the step size.  The vendored CVODE 2.3.0 this path is pinned
     to has no recoverable right-hand-side failures (its RhsFn returns void), so there is no reference arithmetic to reproduce:
     the switch is refused instead of being ignored */
  if (opt->DetectNegState) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "DetectNegState is not supported by the batched integrator (no recoverable right-hand-side failures); unset it with CvodeSettings_setDetectNegState(set, 0)"); return NULL; }
  if (opt->StepResolution < 0) {
    /* event location inside the steps: either layout, plain run without halting, sensitivities or the stop-time mode */
    const char *why = NULL;
    if (opt->HaltOnEvent || opt->SteadyState) why = "the halting modes";
    else if (opt->Sensitivity) why = "sensitivity analysis";
    else if (opt->SetTStop || om->npiecewise) why = "the stop-time mode (SetTStop, piecewise models)";
    if (why) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "StepResolution < 0 (events located inside the steps) cannot be combined with %s", why); return NULL; }
  }
  if (opt->StepResolution > 1 && (opt->HaltOnEvent || opt->SteadyState)) {
    SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "StepResolution > 1 keeps rows at the print steps only: it cannot be combined with HaltOnEvent or HaltOnSteadyState (the halting row may fall between print steps)");
    return NULL;
  }
  odesBatch_t *b = new odesBatch();
  b->om = om; b->opt = opt; b->device = device; b->nvary = nvary; b->nout = opt->PrintStep;
  b->stepRes = opt->StepResolution > 1 ? opt->StepResolution : 1;
  b->rootFind = (opt->StepResolution < 0 && om->nevents > 0) ? 1 : 0;
  b->nvalues = om->neq + om->nass + om->nconst;
  for (int j = 0; j < nvary; j++) {
    if (varyIndex[j] < 0 || varyIndex[j] >= b->nvalues) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "varied index %d out of range", varyIndex[j]); delete b; return NULL; }
    b->vary.push_back(varyIndex[j]);
  }
  if (storeIndex) { for (int j = 0; j < nstore; j++) b->store.push_back(storeIndex[j]); b->nstore = nstore; }
  else { for (int j = 0; j < b->nvalues; j++) b->store.push_back(j); b->nstore = b->nvalues; }
  if (opt->UseJacobian && !om->jacob && !om->jacobianFailed) ODEModel_constructJacobian(om);
  if (!compute_base(om, opt, b->base, b->trigger0)) { delete b; return NULL; }
  /* forward sensitivities (src/sensSolver.c:197-346): no error control on the sensitivity variables; the analytic sense_f
     where the Jacobian and the parametric matrix exist, else CVODES' centred difference quotients over f with the
     sensitivity parameters read from the perturbed array (src/sensSolver.c:312-321, data->use_p), and CVDENSE's
     difference-quotient Jacobian where there is none */
  b->os = NULL; b->nsens = 0; b->d_sens = NULL; b->sensFromRow0 = 0;
  if (opt->Sensitivity) {
    const char *why = NULL;
    if (opt->SensMethod < 0 || opt->SensMethod > 2) why = "SensMethod must be 0 (simultaneous), 1 (staggered) or 2 (staggered1)";
    else if (om->neq < 1 || om->neq > 32) why = "forward sensitivities are implemented for 1..32 equations (one trajectory per warp)";
    if (!why) {
      b->os = ODESense_create(om, opt);
      if (b->os && env_int("ODES_FORCE_SENS_DQ", 0)) b->os->sensitivity = 0;     /* test knob: a parametric matrix that could not be constructed */
      if (!b->os) why = "the sensitivity structures could not be created";
      else if (b->os->nsens < 1 || b->os->nsens > 32) why = "between 1 and 32 sensitivity parameters are supported";
      else if (!(opt->UseJacobian && om->jacobian && b->os->sensitivity)) {
        /* the reference's f overwrites value[index_sens[i]] with p[i] for EVERY sensitivity, a variable's included
           (src/cvodeSolver.c:981-984), which freezes that variable at its initial value inside f: not reproduced */
        for (int i = 0; i < b->os->nsens; i++) if (b->os->index_sensP[i] == -1) why = "difference-quotient sensitivities (no parametric matrix or no Jacobian) are implemented for parameters, not for initial values";
      }
    }
    if (why) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "%s", why); ODESense_free(b->os); delete b; return NULL; }
    b->nsens = b->os->nsens;
  }

  src = build_source(b);
  return b;
}

/* the grid the kernel is driven over: stepRes equal sub-steps per print step (stepRes == 1: the print grid itself) */
static int upload_time_grid(odesBatch_t *b, const double *tpoints, int nout)
{
  const int k = b->stepRes;
  std::vector<double> grid((size_t)nout * (size_t)k + 1);
  for (int i = 0; i < nout; i++)
    for (int j = 0; j < k; j++) grid[(size_t)i * (size_t)k + (size_t)j] = (j == 0) ? tpoints[i] : tpoints[i] + j * ((tpoints[i + 1] - tpoints[i]) / k);
  grid[(size_t)nout * (size_t)k] = tpoints[nout];
  CUDA_OK(cudaMemcpy(b->d_tpoints, grid.data(), sizeof(double) * grid.size(), cudaMemcpyHostToDevice));
  UPLOAD_FENCE();
  return 1;
}
static int ensure_loaded(odesBatch_t *b)
{
  if (b->kernel) { CUDA_OK(cudaSetDevice(b->device)); return 1; }
  if (!need_device(b->device)) return 0;
  b->kernel = (CUfunction)CompiledCode_getFunction(b->code, "odes_bdf_kernel");
  if (!b->kernel) return 0;
  CUdeviceptr sym; size_t bytes;
  DRV_OK(drv.ModuleGetGlobal(&sym, &bytes, (CUmodule)b->code->module, "c_base"));
  if (b->nvalues > 0) DRV_OK(drv.MemcpyHtoD(sym, b->base.data(), sizeof(double) * (size_t)b->nvalues));
  DRV_OK(drv.ModuleGetGlobal(&sym, &bytes, (CUmodule)b->code->module, "c_trigger0"));
  if (b->om->nevents > 0) DRV_OK(drv.MemcpyHtoD(sym, b->trigger0.data(), sizeof(int) * (size_t)b->om->nevents));
  UPLOAD_FENCE();
  if (b->warpMode) {
    unsigned int wmBytes = 0;
    DRV_OK(drv.ModuleGetGlobal(&sym, &bytes, (CUmodule)b->code->module, "c_warpmem_bytes"));
    DRV_OK(drv.MemcpyDtoH(&wmBytes, sym, sizeof wmBytes));
    b->smemBytes = (size_t)wmBytes * (size_t)(b->block / 32);
  }
  {
    /* the opt-in above 48 KB counts the kernel's static shared memory (task tables, staging of the pow requests) as well */
    int staticBytes = 0;
    DRV_OK(drv.FuncGetAttribute(&staticBytes, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, b->kernel));
    if (b->smemBytes + (size_t)staticBytes > 48 * 1024) DRV_OK(drv.FuncSetAttribute(b->kernel, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)b->smemBytes));
  }
  CUDA_OK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  CUDA_OK(cudaStreamCreateWithFlags(&b->copyIn, cudaStreamNonBlocking));
  CUDA_OK(cudaStreamCreateWithFlags(&b->copyOut, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreate(&b->e0)); CUDA_OK(cudaEventCreate(&b->e1));
  for (int i = 0; i < 2; i++) {
    CUDA_OK(cudaEventCreateWithFlags(&b->evIn[i], cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&b->evK[i], cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&b->evOut[i], cudaEventDisableTiming));
  }
  CUDA_OK(cudaMalloc(&b->d_counter, sizeof(unsigned long long) * 64));
  { cudaDeviceProp p; CUDA_OK(cudaGetDeviceProperties(&p, b->device)); b->numSMs = p.multiProcessorCount; }
  { int v = 1; DRV_OK(drv.OccupancyMaxActiveBlocksPerMultiprocessor(&v, b->kernel, b->block, b->smemBytes)); b->blocksPerSM = v < 1 ? 1 : v; }
  if (b->laMode == 2) {
    /* the occupancy query reports 1 block for kernels that allocate tensor memory although two are co-resident
       (measured: forcing 2 blocks per SM raises throughput 1.74 -> 2.16 M trajectories/s); size the persistent grid
       from the register, shared-memory and TMEM-column budgets instead */
    int regs = 0; DRV_OK(drv.FuncGetAttribute(&regs, CU_FUNC_ATTRIBUTE_NUM_REGS, b->kernel));
    const int byRegs = 65536 / (((regs + 7) / 8 * 8) * b->block);
    const int bySmem = (int)((228 * 1024) / (b->smemBytes + 1024 + 16));
    int v = byRegs < bySmem ? byRegs : bySmem;
    if (v > b->maxBlocksPerSM) v = b->maxBlocksPerSM;    /* more blocks would spin in tcgen05.alloc */
    b->blocksPerSM = v < 1 ? 1 : v;
  }
  if (env_int("ODES_BLOCKS_PER_SM", 0) > 0) b->blocksPerSM = env_int("ODES_BLOCKS_PER_SM", 0);
  if (b->laMode == 1) CUDA_OK(cudaMalloc(&b->d_scratch, sizeof(double) * (size_t)b->laWords * (size_t)b->numSMs * (size_t)b->blocksPerSM * (size_t)b->block));
  CUDA_OK(cudaMalloc(&b->d_tpoints, sizeof(double) * ((size_t)b->opt->PrintStep * (size_t)b->stepRes + 1)));
  if (!upload_time_grid(b, b->opt->TimePoints, b->nout)) return 0;
  return 1;
}

extern "C" int ODESBatch_reserve(odesBatch_t *b, int capacity)
{
  if (!ensure_loaded(b)) return 0;
  if (capacity <= b->capacity) return 1;
  capacity = (capacity + 31) & ~31;            /* rows start 256-byte aligned */
  cudaFree(b->d_vary); cudaFree(b->d_out); cudaFree(b->d_status); cudaFree(b->d_stats); cudaFree(b->d_fired); cudaFree(b->d_sens); b->d_sens = NULL;
  b->d_vary = b->d_out = NULL; b->d_status = b->d_stats = NULL; b->d_fired = NULL; b->capacity = 0;
  const size_t cap = (size_t)capacity;
  CUDA_OK(cudaMalloc(&b->d_vary, sizeof(double) * cap * (size_t)(b->nvary > 0 ? b->nvary : 1)));
  CUDA_OK(cudaMalloc(&b->d_out, sizeof(double) * cap * (size_t)(b->opt->PrintStep + 1) * (size_t)(b->nstore > 0 ? b->nstore : 1)));
  CUDA_OK(cudaMalloc(&b->d_status, sizeof(int) * cap));
  CUDA_OK(cudaMalloc(&b->d_stats, sizeof(int) * cap * ODES_NUM_STATS));
  CUDA_OK(cudaMalloc(&b->d_fired, sizeof(unsigned int) * cap * (size_t)(b->opt->PrintStep + 1)));
  if (b->nsens > 0) CUDA_OK(cudaMalloc(&b->d_sens, sizeof(double) * cap * (size_t)(b->opt->PrintStep + 1) * (size_t)b->nsens * (size_t)b->om->neq));
  b->capacity = capacity;
  return 1;
}
/* private: restart support for the single-instance driver (integrator.c) */
extern "C" int ODESBatch_setTimeGrid(odesBatch_t *b, int nout, const double *tpoints)
{
  if (!ensure_loaded(b)) return 0;
  if (nout > b->opt->PrintStep || nout < 0) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "time grid longer than the plan's"); return 0; }
  if (!upload_time_grid(b, tpoints, nout)) return 0;
  b->nout = nout;
  return 1;
}
/* private: initial sensitivities [nsens][neq] of set 0 for a continued course; NULL switches back to the 0 / 1 start */
extern "C" int ODESBatch_setInitialSensitivities(odesBatch_t *b, const double *s0)
{
  if (b->nsens <= 0) return 1;
  if (!ODESBatch_reserve(b, 1)) return 0;
  b->sensFromRow0 = s0 ? 1 : 0;
  if (s0) CUDA_OK(cudaMemcpy(b->d_sens, s0, sizeof(double) * (size_t)b->nsens * (size_t)b->om->neq, cudaMemcpyHostToDevice));
  UPLOAD_FENCE();
  return 1;
}
extern "C" int ODESBatch_setInitialTriggers(odesBatch_t *b, const int *trigger)
{
  if (!ensure_loaded(b)) return 0;
  if (b->om->nevents > 0) {
    CUdeviceptr sym; size_t bytes;
    DRV_OK(drv.ModuleGetGlobal(&sym, &bytes, (CUmodule)b->code->module, "c_trigger0"));
    DRV_OK(drv.MemcpyHtoD(sym, trigger, sizeof(int) * (size_t)b->om->nevents));
    UPLOAD_FENCE();
  }
  return 1;
}
extern "C" void ODESBatch_getBaseValues(const odesBatch_t *b, double *values, int *triggers)
{
  if (values) for (size_t i = 0; i < b->base.size(); i++) values[i] = b->base[i];
  if (triggers) for (size_t i = 0; i < b->trigger0.size(); i++) triggers[i] = b->trigger0[i];
}
extern "C" int ODESBatch_getCapacity(const odesBatch_t *b) { return b->capacity; }
extern "C" int ODESBatch_getNout(const odesBatch_t *b) { return b->nout + 1; }
extern "C" int ODESBatch_getNstore(const odesBatch_t *b) { return b->nstore; }
extern "C" int ODESBatch_getNsens(const odesBatch_t *b) { return b->nsens; }
extern "C" int ODESBatch_getSensIndex(const odesBatch_t *b, int is) { return (b->os && is >= 0 && is < b->nsens) ? b->os->index_sens[is] : -1; }
/* bit 0: the sensitivity right-hand sides are CVODES' difference quotients (no parametric matrix or no Jacobian,
   src/sensSolver.c:312-321); bit 1: the Jacobian is CVDENSE's difference quotient (src/cvodeSolver.c:620-623) */
extern "C" int ODESBatch_getDifferenceQuotients(const odesBatch_t *b)
{
  const int usejac = b->opt->UseJacobian && b->om->jacobian;
  return ((b->nsens > 0 && !(usejac && b->os->sensitivity)) ? 1 : 0) | (usejac ? 0 : 2);
}
extern "C" double *ODESBatch_deviceSensitivities(odesBatch_t *b) { return b->d_sens; }
extern "C" int ODESBatch_getSensitivities(odesBatch_t *b, int nsets, double *sens)
{
  if (b->nsens <= 0) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "the plan was created without sensitivity analysis"); return 0; }
  if (nsets > b->capacity || nsets < 0) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "nsets %d exceeds reserved capacity %d", nsets, b->capacity); return 0; }
  CUDA_OK(cudaSetDevice(b->device));
  const size_t per = (size_t)(b->nout + 1) * (size_t)b->nsens * (size_t)b->om->neq;      /* dense pitch of the current grid, like the kernels */
  CUDA_OK(cudaMemcpyAsync(sens, b->d_sens, sizeof(double) * per * (size_t)nsets, cudaMemcpyDeviceToHost, b->stream));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  return 1;
}
extern "C" int ODESBatch_getNflux(const odesBatch_t *b) { return b->nflux; }
extern "C" void ODESBatch_setEventProcessing(odesBatch_t *b, int on) { if (b) b->noEvents = on ? 0 : 1; }
extern "C" double *ODESBatch_deviceFluxes(odesBatch_t *b) { return b->d_flux; }
extern "C" float ODESBatch_getLastFluxMs(const odesBatch_t *b) { return b->lastFluxMs; }
/* fluxes of the sets [offset, offset + nsets) of the device-resident results, on `stream` */
static int launch_fluxes(odesBatch_t *b, int nsets, size_t offset, cudaStream_t stream)
{
  if (b->nflux <= 0) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "this plan has no flux function (it must store all model values of a model with reactions)"); return 0; }
  if (!ensure_loaded(b)) return 0;
  if (!b->fluxKernel) {
    b->fluxKernel = (CUfunction)CompiledCode_getFunction(b->code, "odes_flux_kernel");
    if (!b->fluxKernel) return 0;
    CUDA_OK(cudaEventCreate(&b->f0)); CUDA_OK(cudaEventCreate(&b->f1));
  }
  const size_t rowsPerSet = (size_t)(b->nout + 1);
  const size_t need = (size_t)b->capacity * rowsPerSet * (size_t)b->nflux;
  if (need > b->fluxCap) { cudaFree(b->d_flux); b->d_flux = NULL; b->fluxCap = 0; CUDA_OK(cudaMalloc(&b->d_flux, need * sizeof(double))); b->fluxCap = need; }
  if (nsets == 0) return 1;
  const int rowsPerBlock = 128;
  const unsigned int smem = (unsigned int)(sizeof(double) * (size_t)rowsPerBlock * (size_t)((b->nvalues | 1) + (b->nflux | 1)));
  if (smem > 48 * 1024) DRV_OK(drv.FuncSetAttribute(b->fluxKernel, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem));
  const double *out = b->d_out + offset * rowsPerSet * (size_t)b->nstore;
  double *flux = b->d_flux + offset * rowsPerSet * (size_t)b->nflux;
  unsigned long long nrows = (unsigned long long)nsets * rowsPerSet;
  int rps = (int)rowsPerSet;
  const double *tp = b->d_tpoints;
  void *params[] = { &out, &tp, &flux, &nrows, &rps };
  unsigned long long blocks = (nrows + rowsPerBlock - 1) / rowsPerBlock;
  const unsigned long long cap = (unsigned long long)b->numSMs * 16ull;          /* grid-stride beyond 16 resident blocks per SM */
  if (blocks > cap) blocks = cap;
  CUDA_OK(cudaEventRecord(b->f0, stream));
  g_counters[CNT_LAUNCH]++;
  DRV_OK(drv.LaunchKernel(b->fluxKernel, (unsigned)blocks, 1, 1, rowsPerBlock, 1, 1, smem, (CUstream)stream, params, NULL));
  CUDA_OK(cudaEventRecord(b->f1, stream));
  b->launches++;
  return 1;
}
extern "C" int ODESBatch_computeFluxes(odesBatch_t *b, int nsets)
{
  if (nsets > b->capacity || nsets < 0) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "nsets %d exceeds reserved capacity %d", nsets, b->capacity); return 0; }
  return launch_fluxes(b, nsets, 0, b->stream);
}
extern "C" int ODESBatch_getFluxes(odesBatch_t *b, int nsets, double *flux)
{
  if (b->nflux <= 0 || !b->d_flux) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "no fluxes have been computed on this plan"); return 0; }
  if (nsets > b->capacity || nsets < 0) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "nsets %d exceeds reserved capacity %d", nsets, b->capacity); return 0; }
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaMemcpyAsync(flux, b->d_flux, sizeof(double) * (size_t)nsets * (size_t)(b->nout + 1) * (size_t)b->nflux, cudaMemcpyDeviceToHost, b->stream));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  { float ms = 0; if (b->f0 && cudaEventElapsedTime(&ms, b->f0, b->f1) == cudaSuccess) b->lastFluxMs = ms; else cudaGetLastError(); }
  return 1;
}
extern "C" const char *ODESBatch_getSource(const odesBatch_t *b) { return b->code ? b->code->source : NULL; }
extern "C" const char *ODESBatch_getCompileLog(const odesBatch_t *b) { return b->code ? b->code->log : NULL; }
extern "C" double *ODESBatch_deviceValues(odesBatch_t *b) { return b->d_vary; }
extern "C" double *ODESBatch_deviceResults(odesBatch_t *b) { return b->d_out; }
extern "C" float ODESBatch_getLastKernelMs(const odesBatch_t *b) { return b->lastMs; }
extern "C" int ODESBatch_getNumLaunches(const odesBatch_t *b) { return b->launches; }

extern "C" int ODESBatch_getKernelInfo(const odesBatch_t *cb, int *regs, int *smem, int *local, int *threads, int *blocksPerSM)
{
  odesBatch_t *b = (odesBatch_t *)cb;
  if (!ensure_loaded(b)) return 0;
  int v = 0;
  if (regs) { DRV_OK(drv.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_NUM_REGS, b->kernel)); *regs = v; }
  if (local) { DRV_OK(drv.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, b->kernel)); *local = v; }
  if (smem) *smem = (int)b->smemBytes;
  if (threads) *threads = b->block;
  if (blocksPerSM) *blocksPerSM = b->blocksPerSM;
  return 1;
}

extern "C" int ODESBatch_setValues(odesBatch_t *b, int nsets, const double *values)
{
  if (!ODESBatch_reserve(b, nsets)) return 0;
  if (b->nvary == 0 || nsets == 0) return 1;
  const size_t n = (size_t)nsets * (size_t)b->nvary;
  if (n > b->rowsCap) { cudaFree(b->d_rows); b->d_rows = NULL; CUDA_OK(cudaMalloc(&b->d_rows, n * sizeof(double))); b->rowsCap = n; }
  CUDA_OK(cudaMemcpyAsync(b->d_rows, values, n * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  odes_rows_to_soa_kernel<<<(nsets + 255) / 256, 256, 0, b->stream>>>(b->d_rows, b->d_vary, (size_t)b->capacity, nsets, b->nvary);
  CUDA_OK(cudaGetLastError());
  return 1;
}

/* device SoA [nvary][capacity] -> host rows [nsets][nvary] (what setValues was given, or what generateValues made) */
extern "C" int ODESBatch_getValues(odesBatch_t *b, int nsets, double *values)
{
  if (!ensure_loaded(b)) return 0;
  if (nsets > b->capacity || nsets < 0) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "nsets %d exceeds reserved capacity %d", nsets, b->capacity); return 0; }
  if (b->nvary == 0 || nsets == 0) return 1;
  std::vector<double> soa((size_t)b->nvary * (size_t)nsets);
  CUDA_OK(cudaStreamSynchronize(b->stream));
  CUDA_OK(cudaMemcpy2D(soa.data(), sizeof(double) * (size_t)nsets, b->d_vary, sizeof(double) * (size_t)b->capacity,
                       sizeof(double) * (size_t)nsets, (size_t)b->nvary, cudaMemcpyDeviceToHost));
  for (int j = 0; j < b->nvary; j++) for (int s = 0; s < nsets; s++) values[(size_t)s * (size_t)b->nvary + (size_t)j] = soa[(size_t)j * (size_t)nsets + (size_t)s];
  return 1;
}

extern "C" int ODESBatch_generateValues(odesBatch_t *b, int nsets, unsigned long long seed, long long firstSet,
                                        const double *base, const double *lo, const double *hi, int log10Scale)
{
  if (!ODESBatch_reserve(b, nsets)) return 0;
  if (b->nvary == 0 || nsets == 0) return 1;
  double *d = NULL;
  CUDA_OK(cudaMalloc(&d, 3 * sizeof(double) * (size_t)b->nvary));
  CUDA_OK(cudaMemcpy(d, base, sizeof(double) * (size_t)b->nvary, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(d + b->nvary, lo, sizeof(double) * (size_t)b->nvary, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(d + 2 * b->nvary, hi, sizeof(double) * (size_t)b->nvary, cudaMemcpyHostToDevice));
  UPLOAD_FENCE();
  odes_generate_values_kernel<<<(nsets + 255) / 256, 256, 0, b->stream>>>(b->d_vary, (size_t)b->capacity, nsets, b->nvary, seed, firstSet,
                                                                         d, d + b->nvary, d + 2 * b->nvary, log10Scale);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(b->stream));
  cudaFree(d);
  return 1;
}

/* One launch = one persistent grid (resident blocks only); lanes pull set indices from a device counter.
   Each launch in flight uses its own counter slot (the pipelined host path has several in flight). */
static int launch(odesBatch_t *b, int nsets, size_t offset, cudaStream_t stream, int timed, int groupShift = -1)
{
  OdesArgs a;
  a.groupDone = NULL; a.hostFlags = NULL; a.groupShift = 0; a.pad2 = b->sensFromRow0;
  if (groupShift >= 0) { a.groupDone = b->d_groupDone; a.hostFlags = b->d_flags; a.groupShift = groupShift; }
  a.vary = b->d_vary + offset; a.out = b->d_out + offset * (size_t)(b->nout + 1) * (size_t)b->nstore; a.status = b->d_status + offset; a.stats = b->d_stats + offset;
  a.fired = b->d_fired + offset; a.tpoints = b->d_tpoints;
  a.sens = b->d_sens ? b->d_sens + offset * (size_t)(b->nout + 1) * (size_t)b->nsens * (size_t)b->om->neq : NULL;
  a.counter = b->d_counter + (b->counterSlot++ & 63);
  a.scratch = b->d_scratch;
  a.stride = (unsigned long long)b->capacity; a.nsets = nsets; a.nout = b->nout * b->stepRes; a.mxstep = b->opt->Mxstep; a.pad = b->noEvents;
  a.reltol = b->opt->RError; a.abstol = b->opt->Error;
  void *params[] = { &a };
  const int perBlock = b->warpMode ? b->block / 32 : b->block;       /* trajectories in flight per block */
  unsigned grid = (unsigned)((nsets + perBlock - 1) / perBlock);
  const unsigned resident = (unsigned)(b->numSMs * b->blocksPerSM);
  if (grid > resident) grid = resident;
  {
    /* small batches: lanes balance their load by pulling sets, which needs a few sets per lane -- beyond one block per SM
       the grid grows only while every lane still gets ODES_MIN_SETS_PER_LANE sets (measured on the events model, 10^5 sets:
       4 resident blocks per SM 4.2 M trajectories/s, 3 blocks 5.1 M; at 10^6 sets 7.4 M vs 6.4 M) */
    const unsigned long long per = (unsigned long long)perBlock * (unsigned long long)(env_int("ODES_MIN_SETS_PER_LANE", 2) > 0 ? env_int("ODES_MIN_SETS_PER_LANE", 2) : 1);
    unsigned cap = (unsigned)(((unsigned long long)nsets + per - 1) / per);
    if (cap < (unsigned)b->numSMs) cap = (unsigned)b->numSMs;
    if (grid > cap) grid = cap;
  }
  CUDA_OK(cudaMemsetAsync(a.counter, 0, sizeof(unsigned long long), stream));
  if (timed) CUDA_OK(cudaEventRecord(b->e0, stream));
  g_counters[CNT_LAUNCH]++;
  DRV_OK(drv.LaunchKernel(b->kernel, grid, 1, 1, (unsigned)b->block, 1, 1, (unsigned)b->smemBytes, (CUstream)stream, params, NULL));
  if (timed) CUDA_OK(cudaEventRecord(b->e1, stream));
  b->launches++;
  return 1;
}

extern "C" int ODESBatch_integrate(odesBatch_t *b, int nsets)
{
  if (nsets > b->capacity || nsets < 0) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "nsets %d exceeds reserved capacity %d", nsets, b->capacity); return 0; }
  if (!ensure_loaded(b)) return 0;
  if (nsets == 0) return 1;
  return launch(b, nsets, 0, b->stream, 1);
}
extern "C" int ODESBatch_synchronize(odesBatch_t *b)
{
  if (!b->stream) return 1;
  CUDA_OK(cudaStreamSynchronize(b->stream));
  if (b->launches) { float ms = 0; if (cudaEventElapsedTime(&ms, b->e0, b->e1) == cudaSuccess) b->lastMs = ms; else cudaGetLastError(); }
  return 1;
}

/* CUDA-event stopwatch on the launching stream (bench.py times K launches with it) */
extern "C" int ODESBatch_timerStart(odesBatch_t *b)
{
  if (!ensure_loaded(b)) return 0;
  if (!b->t0) { CUDA_OK(cudaEventCreate(&b->t0)); CUDA_OK(cudaEventCreate(&b->t1)); }
  CUDA_OK(cudaEventRecord(b->t0, b->stream));
  return 1;
}
extern "C" float ODESBatch_timerStop(odesBatch_t *b)
{
  float ms = -1.0f;
  if (!b->t0) return ms;
  if (cudaEventRecord(b->t1, b->stream) != cudaSuccess || cudaEventSynchronize(b->t1) != cudaSuccess ||
      cudaEventElapsedTime(&ms, b->t0, b->t1) != cudaSuccess) { cudaGetLastError(); return -1.0f; }
  return ms;
}

/* The trajectories on the device are dense in the CURRENT grid: set s starts at s * (nout + 1) * nstore, also after
   ODESBatch_setTimeGrid shortened the grid (the kernels, the launch offsets and these getters use the same pitch). */
static int check_sets(const odesBatch_t *b, int nsets)
{
  if (nsets < 0 || nsets > b->capacity) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "nsets %d exceeds reserved capacity %d", nsets, b->capacity); return 0; }
  return 1;
}
extern "C" int ODESBatch_getResults(odesBatch_t *b, int nsets, double *results)
{
  if (!check_sets(b, nsets)) return 0;
  CUDA_OK(cudaSetDevice(b->device));
  const size_t per = (size_t)(b->nout + 1) * (size_t)b->nstore;
  CUDA_OK(cudaMemcpyAsync(results, b->d_out, sizeof(double) * per * (size_t)nsets, cudaMemcpyDeviceToHost, b->stream));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  return 1;
}
extern "C" int ODESBatch_getTrajectory(odesBatch_t *b, int set, double *traj)
{
  if (set < 0 || set >= b->capacity) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "set %d outside the reserved capacity %d", set, b->capacity); return 0; }
  CUDA_OK(cudaSetDevice(b->device));
  const size_t per = (size_t)(b->nout + 1) * (size_t)b->nstore;
  CUDA_OK(cudaMemcpy(traj, b->d_out + (size_t)set * per, sizeof(double) * per, cudaMemcpyDeviceToHost));
  return 1;
}
extern "C" int ODESBatch_getStatus(odesBatch_t *b, int nsets, int *status)
{
  if (!check_sets(b, nsets)) return 0;
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaMemcpy(status, b->d_status, sizeof(int) * (size_t)nsets, cudaMemcpyDeviceToHost));
  return 1;
}
extern "C" int ODESBatch_getStats(odesBatch_t *b, int nsets, int *stats)
{
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaMemcpy2D(stats, sizeof(int) * (size_t)nsets, b->d_stats, sizeof(int) * (size_t)b->capacity, sizeof(int) * (size_t)nsets, ODES_NUM_STATS, cudaMemcpyDeviceToHost));
  return 1;
}
extern "C" int ODESBatch_getEventLog(odesBatch_t *b, int nsets, unsigned int *fired)
{
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaMemcpy2D(fired, sizeof(unsigned int) * (size_t)nsets, b->d_fired, sizeof(unsigned int) * (size_t)b->capacity,
                       sizeof(unsigned int) * (size_t)nsets, (size_t)(b->nout + 1), cudaMemcpyDeviceToHost));
  return 1;
}

/* End-to-end helper with HOST buffers.  values: [nsets][nvary] rows, results: [nsets][nout+1][nstore].
   The batch is cut into chunks (default: one chunk up to 2^22 sets); per chunk, H2D + transpose, ONE launch, and the
   results stream back while the kernel runs: sets are handed out in index order, the lane that completes a group of
   2^ODES_GROUP_SHIFT consecutive sets raises a flag in mapped host memory, and this thread copies every finished
   group (results are trajectory-major, so a group is one contiguous block) on the copy stream.  Only the last
   groups' copies are left when the kernel ends.  chunk < 0: the older scheme (-chunk sets per launch, copies of
   chunk k-1 overlap the kernel of chunk k). */
static int run_host_chunked(odesBatch_t *b, int nsets, const double *values, double *results, int *status, int chunk)
{
  const int nchunks = (nsets + chunk - 1) / chunk;
  const size_t rows = (size_t)(b->nout + 1) * (size_t)b->nstore;
  for (int k = 0; k < nchunks; k++) {
    const size_t off = (size_t)k * (size_t)chunk;
    const int n = (int)((size_t)nsets - off < (size_t)chunk ? (size_t)nsets - off : (size_t)chunk);
    const int slot = k & 1;
    if (b->nvary > 0) {
      CUDA_OK(cudaMemcpyAsync(b->d_rows + off * (size_t)b->nvary, values + off * (size_t)b->nvary, sizeof(double) * (size_t)n * (size_t)b->nvary, cudaMemcpyHostToDevice, b->copyIn));
      odes_rows_to_soa_kernel<<<(n + 255) / 256, 256, 0, b->copyIn>>>(b->d_rows + off * (size_t)b->nvary, b->d_vary + off, (size_t)b->capacity, n, b->nvary);
    }
    CUDA_OK(cudaEventRecord(b->evIn[slot], b->copyIn));
    CUDA_OK(cudaStreamWaitEvent(b->stream, b->evIn[slot], 0));
    if (!launch(b, n, off, b->stream, k == 0)) return 0;
    CUDA_OK(cudaEventRecord(b->evK[slot], b->stream));
    CUDA_OK(cudaStreamWaitEvent(b->copyOut, b->evK[slot], 0));
    if (results)
      CUDA_OK(cudaMemcpyAsync(results + off * rows, b->d_out + off * rows, sizeof(double) * (size_t)n * rows, cudaMemcpyDeviceToHost, b->copyOut));
    if (status) CUDA_OK(cudaMemcpyAsync(status + off, b->d_status + off, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, b->copyOut));
  }
  CUDA_OK(cudaStreamSynchronize(b->copyOut));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  return 1;
}
/* Results for PAGEABLE caller memory.  A device-to-host copy into pageable memory runs at 17 GB/s on this pool's boxes and
   pinning the caller's array in place costs more than it saves (cudaHostRegister 4.6 GB/s), while copies into pinned memory
   run at 54 GB/s and eight host threads move pinned blocks into pageable memory at 43 GB/s (tools/micro/d2h_ceiling.cu,
   profiles/r02_d2h_ceiling.json).  So finished groups are copied into a small pinned ring the plan owns and a few host
   threads move each block to its place in the caller's array while the next blocks are in flight. */
struct StagePiece { size_t s0, cnt; int slot; };
struct Stager {
  odesBatch_t *b; double *dst; size_t rows, slotSets; int R;
  std::vector<StagePiece> pieces; size_t taken; bool closed;
  std::mutex mu; std::condition_variable cv;
  std::vector<std::atomic<long long>> slotDone;       /* index of the last piece consumed from a slot */
  std::vector<std::thread> workers;
  std::atomic<int> failed;
  Stager(odesBatch_t *b_, double *dst_, size_t rows_, int nthreads)
    : b(b_), dst(dst_), rows(rows_), slotSets(b_->stageSlotBytes / (rows_ * sizeof(double))), R(b_->stageSlots), taken(0), closed(false), slotDone((size_t)b_->stageSlots), failed(0)
  {
    for (int i = 0; i < R; i++) slotDone[(size_t)i].store(-1);
    for (int t = 0; t < nthreads; t++) workers.emplace_back([this]() { run(); });
  }
  void run()
  {
    cudaSetDevice(b->device);
    for (;;) {
      StagePiece pc; long long index;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [this]() { return taken < pieces.size() || closed; });
        if (taken >= pieces.size()) return;
        index = (long long)taken; pc = pieces[taken++];
      }
      if (cudaEventSynchronize(b->stageEv[pc.slot]) != cudaSuccess) { cudaGetLastError(); failed.store(1); }
      memcpy(dst + pc.s0 * rows, b->stage + (size_t)pc.slot * b->stageSlotBytes, pc.cnt * rows * sizeof(double));
      slotDone[(size_t)pc.slot].store(index);
    }
  }
  /* main thread: sets [s0, s0 + cnt) have landed in d_out; split into slot-sized pieces and send them through the ring */
  int emit(size_t s0, size_t cnt)
  {
    while (cnt > 0) {
      const size_t n = cnt < slotSets ? cnt : slotSets;
      size_t p;
      { std::lock_guard<std::mutex> lk(mu); p = pieces.size(); }
      const int slot = (int)(p % (size_t)R);
      while (p >= (size_t)R && slotDone[(size_t)slot].load() < (long long)(p - (size_t)R)) sched_yield();
      CUDA_OK(cudaMemcpyAsync(b->stage + (size_t)slot * b->stageSlotBytes, b->d_out + s0 * rows, n * rows * sizeof(double), cudaMemcpyDeviceToHost, b->copyOut));
      CUDA_OK(cudaEventRecord(b->stageEv[slot], b->copyOut));
      { std::lock_guard<std::mutex> lk(mu); StagePiece pc; pc.s0 = s0; pc.cnt = n; pc.slot = slot; pieces.push_back(pc); }
      cv.notify_one();
      s0 += n; cnt -= n;
    }
    return 1;
  }
  int finish()
  {
    { std::lock_guard<std::mutex> lk(mu); closed = true; }
    cv.notify_all();
    for (auto &w : workers) w.join();
    workers.clear();
    return failed.load() ? 0 : 1;
  }
  ~Stager() { if (!workers.empty()) finish(); }
};
static bool is_pageable(const void *p)
{
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return at.type == cudaMemoryTypeUnregistered;
}
extern "C" int ODESBatch_runHost(odesBatch_t *b, int nsets, const double *values, double *results, int *status, int chunk)
{ return ODESBatch_runHostSens(b, nsets, values, results, NULL, status, chunk); }
extern "C" int ODESBatch_runHostSens(odesBatch_t *b, int nsets, const double *values, double *results, double *sens, int *status, int chunk)
{ return ODESBatch_runHostEx(b, nsets, values, results, sens, NULL, status, chunk); }
extern "C" int ODESBatch_runHostEx(odesBatch_t *b, int nsets, const double *values, double *results, double *sens, double *flux, int *status, int chunk)
{
  if (flux && b->nflux <= 0) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "this plan has no flux function (it must store all model values of a model with reactions)"); return 0; }
  if (flux && chunk < 0) chunk = -chunk;
  if (sens && b->nsens <= 0) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "the plan was created without sensitivity analysis"); return 0; }
  if (sens && chunk < 0) chunk = -chunk;              /* sensitivities travel with the streamed pipeline only */
  if (!ODESBatch_reserve(b, nsets)) return 0;
  const size_t srows = (size_t)(b->nout + 1) * (size_t)(b->nsens > 0 ? b->nsens : 0) * (size_t)b->om->neq;
  const size_t rows = (size_t)(b->nout + 1) * (size_t)b->nstore;
  const size_t need = (size_t)nsets * (size_t)(b->nvary > 0 ? b->nvary : 1);
  if (need > b->rowsCap) { cudaFree(b->d_rows); b->d_rows = NULL; CUDA_OK(cudaMalloc(&b->d_rows, need * sizeof(double))); b->rowsCap = need; }
  if (chunk < 0) return run_host_chunked(b, nsets, values, results, status, ((-chunk) + 31) & ~31);
  if (chunk == 0) chunk = 1 << 22;
  chunk = (chunk + 31) & ~31;
  int shift = env_int("ODES_GROUP_SHIFT", 14);
  /* pageable destination: stage through the pinned ring, in groups of at most ~64 MB */
  const bool staged = results && rows > 0 && env_int("ODES_STAGE", 1) != 0 && is_pageable(results);
  if (staged) {
    const size_t target = (size_t)env_int("ODES_STAGE_MB", 64) << 20;
    while (shift > 5 && ((size_t)1 << shift) * rows * sizeof(double) > target) shift--;
    const size_t slotBytes = ((size_t)1 << shift) * rows * sizeof(double);
    const int slots = env_int("ODES_STAGE_SLOTS", 8) > 8 ? 8 : (env_int("ODES_STAGE_SLOTS", 8) < 2 ? 2 : env_int("ODES_STAGE_SLOTS", 8));
    if (!b->stage || b->stageSlotBytes < slotBytes || b->stageSlots != slots) {
      if (b->stage) { cudaFreeHost(b->stage); b->stage = NULL; }
      CUDA_OK(cudaHostAlloc((void **)&b->stage, slotBytes * (size_t)slots, cudaHostAllocDefault));
      g_counters[CNT_DEVICE_ALLOC]++;
      b->stageSlotBytes = slotBytes; b->stageSlots = slots;
      for (int i = 0; i < 8; i++) if (!b->stageEv[i]) CUDA_OK(cudaEventCreateWithFlags(&b->stageEv[i], cudaEventDisableTiming));
    }
  }
  const int gsize = 1 << shift;
  const int maxGroups = (chunk + gsize - 1) / gsize;
  if (maxGroups > b->groupCap) {
    cudaFree(b->d_groupDone); if (b->h_flags) cudaFreeHost(b->h_flags);
    b->d_groupDone = NULL; b->h_flags = NULL; b->groupCap = 0;
    CUDA_OK(cudaMalloc(&b->d_groupDone, sizeof(unsigned int) * (size_t)maxGroups));
    g_counters[CNT_DEVICE_ALLOC]++;
    CUDA_OK(cudaHostAlloc((void **)&b->h_flags, sizeof(unsigned int) * (size_t)maxGroups, cudaHostAllocMapped));
    CUDA_OK(cudaHostGetDevicePointer((void **)&b->d_flags, b->h_flags, 0));
    b->groupCap = maxGroups;
  }
  for (size_t off = 0; off < (size_t)nsets; off += (size_t)chunk) {
    const int n = (int)((size_t)nsets - off < (size_t)chunk ? (size_t)nsets - off : (size_t)chunk);
    const int ngroups = (n + gsize - 1) / gsize;
    if (b->nvary > 0) {
      CUDA_OK(cudaMemcpyAsync(b->d_rows + off * (size_t)b->nvary, values + off * (size_t)b->nvary, sizeof(double) * (size_t)n * (size_t)b->nvary, cudaMemcpyHostToDevice, b->stream));
      odes_rows_to_soa_kernel<<<(n + 255) / 256, 256, 0, b->stream>>>(b->d_rows + off * (size_t)b->nvary, b->d_vary + off, (size_t)b->capacity, n, b->nvary);
    }
    CUDA_OK(cudaMemsetAsync(b->d_groupDone, 0, sizeof(unsigned int) * (size_t)ngroups, b->stream));
    for (int g = 0; g < ngroups; g++) ((volatile unsigned int *)b->h_flags)[g] = 0u;
    if (!launch(b, n, off, b->stream, off == 0, shift)) return 0;
    CUDA_OK(cudaEventRecord(b->evK[0], b->stream));
    /* copy finished groups while the kernel runs; consecutive finished groups go out as one copy */
    Stager *stager = staged ? new Stager(b, results, rows, env_int("ODES_STAGE_THREADS", 8)) : NULL;
    struct StagerGuard { Stager *s; ~StagerGuard() { delete s; } } stagerGuard = { stager };
    int next = 0;
    while (next < ngroups) {
      const cudaError_t q = cudaEventQuery(b->evK[0]);
      if (q != cudaSuccess && q != cudaErrorNotReady) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_CUDA_CALL_FAILED, "kernel failed: %s", cudaGetErrorString(q)); return 0; }
      if (q == cudaErrorNotReady) (void)cudaGetLastError();
      const bool over = (q == cudaSuccess);
      int upto = next;
      while (upto < ngroups && (over || ((volatile unsigned int *)b->h_flags)[upto])) upto++;
      if (upto > next) {
        const size_t s0 = off + (size_t)next * (size_t)gsize;
        const size_t cnt = (size_t)(upto == ngroups ? n : upto * gsize) - (size_t)next * (size_t)gsize;
        if (stager) { if (!stager->emit(s0, cnt)) return 0; }
        else if (results) CUDA_OK(cudaMemcpyAsync(results + s0 * rows, b->d_out + s0 * rows, sizeof(double) * cnt * rows, cudaMemcpyDeviceToHost, b->copyOut));
        if (status) CUDA_OK(cudaMemcpyAsync(status + s0, b->d_status + s0, sizeof(int) * cnt, cudaMemcpyDeviceToHost, b->copyOut));
        if (sens) CUDA_OK(cudaMemcpyAsync(sens + s0 * srows, b->d_sens + s0 * srows, sizeof(double) * cnt * srows, cudaMemcpyDeviceToHost, b->copyOut));
        next = upto;
      } else sched_yield();
    }
    if (flux) {
      /* the fluxes of this launch's rows: one pass over the results in HBM, then their copy */
      const size_t frows = (size_t)(b->nout + 1) * (size_t)b->nflux;
      if (!launch_fluxes(b, n, off, b->stream)) return 0;
      CUDA_OK(cudaMemcpyAsync(flux + off * frows, b->d_flux + off * frows, sizeof(double) * (size_t)n * frows, cudaMemcpyDeviceToHost, b->stream));
    }
    CUDA_OK(cudaStreamSynchronize(b->copyOut));
    CUDA_OK(cudaStreamSynchronize(b->stream));
    if (stager && !stager->finish()) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_CUDA_CALL_FAILED, "staged device-to-host copy failed"); return 0; }
    if (flux) { float ms = 0; if (cudaEventElapsedTime(&ms, b->f0, b->f1) == cudaSuccess) b->lastFluxMs = ms; else cudaGetLastError(); }
  }
  return 1;
}

/* pinned host buffers for callers that want full-speed host<->device copies */
extern "C" void *ODES_hostAlloc(size_t bytes) { void *p = NULL; if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return NULL; } return p; }
extern "C" void ODES_hostFree(void *p) { if (p) cudaFreeHost(p); }

/* The batch over all GPUs of the box: device k integrates the index range [k n / G, (k+1) n / G) --
   trajectories are independent, there is no exchange step and no collective.  One plan per device (the generated
   source is identical, so devices after the first load the cached cubin), one host thread per device driving
   ODESBatch_runHost on its slice of the caller's arrays.  ndevices <= 0: all devices for batches of at least 2^15
   sets per device (ODES_DEVICES overrides), else one. */
extern "C" int ODESBatch_runHostAllDevices(odeModel_t *om, cvodeSettings_t *opt, int nvary, const int *varyIndex, int nstore, const int *storeIndex,
                                           int nsets, const double *values, double *results, int *status, int ndevices)
{ return ODESBatch_runHostAllDevicesSens(om, opt, nvary, varyIndex, nstore, storeIndex, nsets, values, results, NULL, status, ndevices); }
extern "C" int ODESBatch_runHostAllDevicesSens(odeModel_t *om, cvodeSettings_t *opt, int nvary, const int *varyIndex, int nstore, const int *storeIndex,
                                               int nsets, const double *values, double *results, double *sens, int *status, int ndevices)
{ return ODESBatch_runHostAllDevicesEx(om, opt, nvary, varyIndex, nstore, storeIndex, nsets, values, results, sens, NULL, status, ndevices); }
/* ---- the plan cache of a model ("library owns device buffers cached on ii/om") ----
   The reference's batch loop reuses ONE integrator instance for all design points and all calls (src/odeSolver.c:313-330).
   The one-shot entries here keep what is expensive to make -- the compiled module, the loaded kernel, streams, events and
   the device buffers of every GPU -- in odeModel_t.kernelCache, keyed by the generated kernel text (which folds in the model,
   the varied and stored index sets and every code-relevant setting).  Tolerances, step limit, time grid and the model's
   current base values are refreshed on every call.  Freed by ODEModel_free / ODEModel_freeKernelCache.
   ODES_PLAN_CACHE=0 builds and frees the plans per call as before. */
struct PlanSet { std::string src; cvodeSettings_t *opt; std::vector<odesBatch_t *> plans; std::mutex run; unsigned long long lastUse; };
struct KernelCache { std::mutex mu; std::vector<PlanSet *> sets; unsigned long long clock; };
static std::mutex g_cacheCreate;
static void planset_free(PlanSet *ps)
{
  for (size_t k = 0; k < ps->plans.size(); k++) if (ps->plans[k]) ODESBatch_free(ps->plans[k]);
  if (ps->opt) CvodeSettings_free(ps->opt);
  delete ps;
}
extern "C" void ODEModel_freeKernelCache(odeModel_t *om)
{
  if (!om || !om->kernelCache) return;
  KernelCache *kc = (KernelCache *)om->kernelCache;
  for (size_t i = 0; i < kc->sets.size(); i++) planset_free(kc->sets[i]);
  delete kc;
  om->kernelCache = NULL;
}
/* the caller's settings into the cached clone: everything a launch reads (the code-relevant fields are equal by key) */
static void refresh_settings(cvodeSettings_t *dst, const cvodeSettings_t *src)
{
  dst->Time = src->Time; dst->Error = src->Error; dst->RError = src->RError; dst->Mxstep = src->Mxstep;
  dst->StoreResults = src->StoreResults; dst->DetectNegState = src->DetectNegState;
  if (src->TimePoints && dst->TimePoints) for (int i = 0; i <= src->PrintStep; i++) dst->TimePoints[i] = src->TimePoints[i];
}
/* a cached plan before its next run: the model's current base values and trigger flags, the time grid */
static int refresh_plan(odesBatch_t *b)
{
  if (!compute_base(b->om, b->opt, b->base, b->trigger0)) return 0;
  b->nout = b->opt->PrintStep;
  if (!b->kernel) return 1;                               /* not loaded yet: ensure_loaded uploads everything */
  CUDA_OK(cudaSetDevice(b->device));
  CUdeviceptr sym; size_t bytes;
  DRV_OK(drv.ModuleGetGlobal(&sym, &bytes, (CUmodule)b->code->module, "c_base"));
  if (b->nvalues > 0) DRV_OK(drv.MemcpyHtoD(sym, b->base.data(), sizeof(double) * (size_t)b->nvalues));
  DRV_OK(drv.ModuleGetGlobal(&sym, &bytes, (CUmodule)b->code->module, "c_trigger0"));
  if (b->om->nevents > 0) DRV_OK(drv.MemcpyHtoD(sym, b->trigger0.data(), sizeof(int) * (size_t)b->om->nevents));
  UPLOAD_FENCE();
  return upload_time_grid(b, b->opt->TimePoints, b->nout);
}
static odesBatch_t *batch_prepare(odeModel_t *om, cvodeSettings_t *opt, int nvary, const int *varyIndex, int nstore, const int *storeIndex, int device, std::string &src);

extern "C" int ODESBatch_runHostAllDevicesEx(odeModel_t *om, cvodeSettings_t *opt, int nvary, const int *varyIndex, int nstore, const int *storeIndex,
                                             int nsets, const double *values, double *results, double *sens, double *flux, int *status, int ndevices)
{
  int have = ODES_deviceCount();
  if (ndevices <= 0) {
    ndevices = env_int("ODES_DEVICES", 0);
    if (ndevices <= 0) { ndevices = have; while (ndevices > 1 && nsets / ndevices < (1 << 15)) ndevices--; }
  }
  if (ndevices > have) ndevices = have;
  if (ndevices < 1) ndevices = 1;                              /* no device: the single plan fails loudly below */
  const bool cached = env_int("ODES_PLAN_CACHE", 1) != 0 && have > 0;
  std::vector<odesBatch_t *> plans((size_t)ndevices, (odesBatch_t *)NULL);
  PlanSet *ps = NULL;
  int ok = 1;
  if (cached) {
    /* look the plan set up by the kernel text this call would compile */
    std::string src;
    odesBatch_t *probe = batch_prepare(om, opt, nvary, varyIndex, nstore, storeIndex, 0, src);
    if (!probe) return 0;
    ODESBatch_free(probe);
    std::lock_guard<std::mutex> g(g_cacheCreate);
    if (!om->kernelCache) { KernelCache *n = new KernelCache(); n->clock = 0; om->kernelCache = n; }
    KernelCache *kc = (KernelCache *)om->kernelCache;
    for (size_t i = 0; i < kc->sets.size(); i++) if (kc->sets[i]->src == src) ps = kc->sets[i];
    if (ps) g_counters[CNT_PLAN_HIT]++;
    else {
      /* at most four plan sets per model: the least recently used one makes room (its device buffers are freed) */
      if (kc->sets.size() >= (size_t)env_int("ODES_PLAN_CACHE_SETS", 4)) {
        size_t lru = 0;
        for (size_t i = 1; i < kc->sets.size(); i++) if (kc->sets[i]->lastUse < kc->sets[lru]->lastUse) lru = i;
        planset_free(kc->sets[lru]); kc->sets.erase(kc->sets.begin() + (long)lru);
      }
      ps = new PlanSet(); ps->src = src; ps->opt = CvodeSettings_clone(opt);
      kc->sets.push_back(ps);
    }
    ps->lastUse = ++kc->clock;
  }
  std::unique_lock<std::mutex> running;
  if (ps) {
    running = std::unique_lock<std::mutex>(ps->run);          /* one run at a time on a plan set */
    refresh_settings(ps->opt, opt);
    if (ps->plans.size() < (size_t)ndevices) ps->plans.resize((size_t)ndevices, (odesBatch_t *)NULL);
    for (int k = 0; k < ndevices && ok; k++) {
      if (!ps->plans[(size_t)k]) ps->plans[(size_t)k] = ODESBatch_create(om, ps->opt, nvary, varyIndex, nstore, storeIndex, k);
      else if (!refresh_plan(ps->plans[(size_t)k])) ok = 0;
      plans[(size_t)k] = ps->plans[(size_t)k];
      if (!plans[(size_t)k]) ok = 0;
    }
  } else {
    for (int k = 0; k < ndevices && ok; k++) { plans[(size_t)k] = ODESBatch_create(om, opt, nvary, varyIndex, nstore, storeIndex, k); if (!plans[(size_t)k]) ok = 0; }
  }
  if (ok) {
    const size_t rows = (size_t)(opt->PrintStep + 1) * (size_t)plans[0]->nstore;
    const size_t srows = (size_t)(opt->PrintStep + 1) * (size_t)(plans[0]->nsens > 0 ? plans[0]->nsens : 0) * (size_t)om->neq;
    if (sens && plans[0]->nsens <= 0) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "sensitivities requested but opt->Sensitivity is off"); ok = 0; }
    if (flux && plans[0]->nflux <= 0) { SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_BATCH_ARGUMENT, "fluxes requested but the plan has no flux function"); ok = 0; }
    const size_t frows = (size_t)(opt->PrintStep + 1) * (size_t)(plans[0]->nflux > 0 ? plans[0]->nflux : 0);
    std::vector<int> done((size_t)ndevices, 0);
    std::vector<std::thread> workers;
    for (int k = 0; k < ndevices && ok; k++) {
      size_t lo = ((size_t)nsets * (size_t)k / (size_t)ndevices) & ~(size_t)31, hi = (k + 1 == ndevices) ? (size_t)nsets : (((size_t)nsets * (size_t)(k + 1) / (size_t)ndevices) & ~(size_t)31);
      workers.emplace_back([&, k, lo, hi]() {
        done[(size_t)k] = (hi <= lo) ? 1 : ODESBatch_runHostEx(plans[(size_t)k], (int)(hi - lo), values + lo * (size_t)(nvary > 0 ? nvary : 0),
                                                               results ? results + lo * rows : NULL, sens ? sens + lo * srows : NULL,
                                                               flux ? flux + lo * frows : NULL, status ? status + lo : NULL, 0);
      });
    }
    for (auto &w : workers) w.join();
    for (int k = 0; k < ndevices; k++) if (!done[(size_t)k]) ok = 0;
  }
  if (!ps) for (int k = 0; k < ndevices; k++) if (plans[(size_t)k]) ODESBatch_free(plans[(size_t)k]);
  return ok ? ndevices : 0;
}

/* one-shot drop-in for the reference's batch loop */
extern "C" int IntegratorInstance_integrateBatch(integratorInstance_t *ii, int nsets, int nvary, variableIndex_t *const *vi,
                                                 const double *values, double *out, int *status)
{ return IntegratorInstance_integrateBatchSens(ii, nsets, nvary, vi, values, out, NULL, status); }
extern "C" int IntegratorInstance_integrateBatchSens(integratorInstance_t *ii, int nsets, int nvary, variableIndex_t *const *vi,
                                                     const double *values, double *out, double *sens, int *status)
{
  if (!ii || nsets < 0) return -1;
  std::vector<int> idx;
  for (int j = 0; j < nvary; j++) idx.push_back(vi[j]->index);
  int failed = -1;
  std::vector<int> st((size_t)(nsets > 0 ? nsets : 1));
  if (ODESBatch_runHostAllDevicesSens(ii->om, ii->opt, nvary, idx.data(), 0, NULL, nsets, values, out, sens, st.data(), 0)) {       /* out is already [nsets][nout+1][nvalues] */
    failed = 0;
    for (int s = 0; s < nsets; s++) { if (st[(size_t)s] < 0) failed++; if (status) status[s] = st[(size_t)s]; }
  }
  return failed;
}

/* ---- the callback seam of the reference (src/sbmlsolver/odeModel.h:356-363, src/odeModel.c:3191-3299, :3382-3440) ----
   ODEModel_compileCVODEFunctions assembles the model's right-hand side, Jacobian and event functions, compiles them and
   keeps the code on the model; the getters hand out what CVODE is given as callbacks.  Here the compiled code is the
   model's kernel module for sm_100a (the right-hand side, Jacobian and event functions are device functions inside it), kept
   in the model's plan cache under DEFAULT settings, and the getters return the CUfunction of that module's kernel as an
   opaque handle -- there is no host-callable function to return; NULL with a SolverError when the Jacobian / parametric
   matrix is missing (the reference's error codes) or when no device can load the module. */
static cvodeSettings_t *seam_settings(const odeSense_t *os)
{
  cvodeSettings_t *opt = CvodeSettings_create();
  if (opt && os) {
    std::vector<char *> ids;
    for (int k = 0; k < os->nsens; k++) ids.push_back(os->om->names[os->index_sens[k]]);
    opt->Sensitivity = 1;
    CvodeSettings_setSensParams(opt, ids.data(), (int)ids.size());
  }
  return opt;
}
static int seam_compile(odeModel_t *om, const odeSense_t *os)
{
  if (!om) return 0;
  cvodeSettings_t *opt = seam_settings(os);
  if (!opt) return 0;
  /* a run over zero sets: generates, compiles and caches the plan (without a device: compiles and drops it) */
  const int ok = ODESBatch_runHostAllDevicesEx(om, opt, 0, NULL, 0, NULL, 0, NULL, NULL, NULL, NULL, NULL, 1) > 0;
  CvodeSettings_free(opt);
  return ok;
}
static void *seam_kernel(odeModel_t *om, const odeSense_t *os)
{
  if (!seam_compile(om, os)) return NULL;
  if (!om->kernelCache) { need_device(0); return NULL; }                  /* no device: the error is on the stack */
  cvodeSettings_t *opt = seam_settings(os);
  std::string src;
  odesBatch_t *probe = batch_prepare(om, opt, 0, NULL, 0, NULL, 0, src);
  if (probe) ODESBatch_free(probe);
  CvodeSettings_free(opt);
  KernelCache *kc = (KernelCache *)om->kernelCache;
  std::lock_guard<std::mutex> g(g_cacheCreate);
  for (size_t i = 0; i < kc->sets.size(); i++)
    if (kc->sets[i]->src == src && !kc->sets[i]->plans.empty() && kc->sets[i]->plans[0]) {
      odesBatch_t *b = kc->sets[i]->plans[0];
      if (!ensure_loaded(b)) return NULL;
      return (void *)b->kernel;
    }
  return NULL;
}
extern "C" int ODEModel_compileCVODEFunctions(odeModel_t *om) { return seam_compile(om, NULL); }
extern "C" void *ODEModel_getCompiledCVODERHSFunction(odeModel_t *om) { return om ? seam_kernel(om, NULL) : NULL; }
extern "C" void *ODEModel_getCompiledCVODEJacobianFunction(odeModel_t *om)
{
  if (!om) return NULL;
  if (!om->jacobian) {
    SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_CANNOT_COMPILE_JACOBIAN_NOT_COMPUTED,
                      "Attempting to compile jacobian before the jacobian is computed\nCall ODEModel_constructJacobian before calling\n"
                      "ODEModel_getCompiledCVODEJacobianFunction or ODEModel_getCompiledCVODERHSFunction\n");
    return NULL;
  }
  return seam_kernel(om, NULL);
}
extern "C" int ODESense_compileCVODESenseFunctions(odeSense_t *os) { return (os && os->om && os->sensitivity) ? seam_compile(os->om, os) : 0; }
extern "C" void *ODESense_getCompiledCVODESenseFunction(odeSense_t *os)
{
  if (!os || !os->om) return NULL;
  if (!os->sensitivity) {
    SolverError_error(ERROR_ERROR_TYPE, SOLVER_ERROR_CANNOT_COMPILE_SENSITIVITY_NOT_COMPUTED,
                      "Attempting to compile sensitivity matrix before the matrix is computed\nCall ODESense_constructSensitivity before calling\n"
                      "ODESense_getCompiledCVODESenseFunction\n");
    return NULL;
  }
  return seam_kernel(os->om, os);
}
