// Microbenchmark: tensor memory (TMEM) as a per-lane scratchpad -- latency of dependent tcgen05.st/ld pairs and
// throughput of independent x16 loads, compared with shared memory.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_ld16(uint32_t addr, uint32_t (&r)[16])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(addr));
}
__device__ __forceinline__ void tm_st16(uint32_t addr, const uint32_t (&r)[16])
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
               :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                  "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__global__ void __launch_bounds__(128) probe(long long *out, int iters)
{
  __shared__ uint32_t base_s;
  __shared__ double sm[16 * 128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "l"((uint64_t)__cvta_generic_to_shared(&base_s)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t base = base_s + ((uint32_t)(warp * 32) << 16);       // this warp's 32 lanes, column 0
  uint32_t v[16];
  for (int k = 0; k < 16; k++) v[k] = threadIdx.x * 16 + k;
  // correctness: per-lane private data round trip at several column offsets
  int bad = 0;
  for (int c = 0; c < 256; c += 16) { for (int k = 0; k < 16; k++) v[k] = threadIdx.x * 1000 + c + k; tm_st16(base + c, v); }
  tm_wait_st();
  for (int c = 0; c < 256; c += 16) { uint32_t r[16]; tm_ld16(base + c, r); tm_wait_ld(); for (int k = 0; k < 16; k++) if (r[k] != threadIdx.x * 1000 + c + k) bad++; }
  // dependent chain: st -> wait -> ld -> wait
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    tm_st16(base + ((it * 16) & 255), v); tm_wait_st();
    tm_ld16(base + ((it * 16) & 255), v); tm_wait_ld();
    v[0] += 1;
  }
  long long t1 = clock64();
  // dependent loads only (address depends on the loaded value)
  uint32_t a = 0;
  for (int it = 0; it < iters; it++) { uint32_t r[16]; tm_ld16(base + (a & 240), r); tm_wait_ld(); a = (r[0] & 0) + ((it * 16) & 255); v[1] += r[1]; }
  long long t2 = clock64();
  // independent loads, one wait per 4
  for (int it = 0; it < iters; it += 4) {
    uint32_t r0[16], r1[16], r2[16], r3[16];
    tm_ld16(base + 0, r0); tm_ld16(base + 16, r1); tm_ld16(base + 32, r2); tm_ld16(base + 48, r3); tm_wait_ld();
    v[2] += r0[0] + r1[0] + r2[0] + r3[0];
  }
  long long t3 = clock64();
  // shared memory comparison: 8 dependent double loads (lane-minor layout) per iteration
  for (int k = 0; k < 16; k++) sm[k * 128 + threadIdx.x] = k;
  double acc = 0; int idx = 0;
  for (int it = 0; it < iters; it++) { double s = 0; for (int k = 0; k < 8; k++) s += sm[((idx + k) & 15) * 128 + threadIdx.x]; acc += s; idx = ((int)s) & 1; }
  long long t4 = clock64();
  if (lane == 0) {
    out[warp * 8 + 0] = (t1 - t0) / iters; out[warp * 8 + 1] = (t2 - t1) / iters; out[warp * 8 + 2] = (t3 - t2) / iters;
    out[warp * 8 + 3] = (t4 - t3) / iters; out[warp * 8 + 4] = bad; out[warp * 8 + 5] = v[0] + v[1] + v[2] + (long long)acc;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(base_s), "n"(256));
}

int main()
{
  long long *d, h[32];
  cudaMalloc(&d, sizeof h);
  for (int blocks = 1; blocks <= 296; blocks *= 296) {
    probe<<<blocks, 128>>>(d, 4096);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    for (int w = 0; w < 4; w++)
      printf("blocks %d warp %d: st+ld chain %lld cyc, dependent x16 ld %lld cyc, independent x16 ld %lld cyc/ld, smem 8 dependent-ish LDS.64 %lld cyc, mismatches %lld\n",
             blocks, w, h[w * 8], h[w * 8 + 1], h[w * 8 + 2], h[w * 8 + 3], h[w * 8 + 4]);
  }
  return 0;
}
