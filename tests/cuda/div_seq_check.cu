/* Checks csrc/exact_div.inc's odes_div_seq (the compiler's division sequence with the reciprocal part hoisted) against
   the division itself on the GPU, bit for bit: random mantissas over narrow and wide exponent ranges, divisors next
   to powers of two and with all-ones mantissas, quotients next to rounding boundaries, zeros / infinities / NaNs /
   denormals.  Prints "pairs <n> mismatches <m>"; used by tests/test_exact_div.py (gpu tier). */
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define DEV __device__ __forceinline__
#include "../../sbml_odesolver_b200/csrc/exact_div.inc"

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
  x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull; return x ^ (x >> 31);
}
__device__ __forceinline__ double make(uint64_t r, int span, int kind)
{
  /* sign, exponent in [-span, span), 52-bit mantissa of the chosen kind */
  uint64_t man = r & 0xfffffffffffffull;
  if (kind == 1) man &= 0xfffull;                                  /* just above a power of two */
  if (kind == 2) man |= 0xfffffffffff000ull;                       /* just below the next one */
  if (kind == 3) man &= 0xfffff00000000ull;                        /* short mantissa: exact quotients and ties are frequent */
  const int e = (int)((r >> 52) % (uint64_t)(2 * span)) - span;
  const uint64_t bits = ((r >> 63) << 63) | ((uint64_t)(e + 1023) << 52) | man;
  return __longlong_as_double((long long)bits);
}
__global__ void check(unsigned long long per_thread, unsigned long long *bad, unsigned long long seed)
{
  const unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  unsigned long long nbad = 0;
  const double sp[12] = {0.0, -0.0, __longlong_as_double(0x7ff0000000000000ll), __longlong_as_double(0xfff0000000000000ll),
                         __longlong_as_double(0x7ff8000000000000ll), 4.9406564584124654e-324, 1e-310, 1e308, -1e-200, 3.0, 2.2250738585072014e-308, 1.7976931348623157e308};
  for (unsigned long long i = 0; i < per_thread; i++) {
    const uint64_t r1 = mix(seed + 2 * (tid * per_thread + i)), r2 = mix(seed + 2 * (tid * per_thread + i) + 1);
    const int sel = (int)(i % 16);
    double a, d;
    if (sel < 4)       { a = make(r1, 3, 0);    d = make(r2, 3, 0); }
    else if (sel < 7)  { a = make(r1, 40, 0);   d = make(r2, 40, 0); }
    else if (sel < 9)  { a = make(r1, 1000, 0); d = make(r2, 1000, 0); }         /* through the range tests */
    else if (sel < 11) { a = make(r1, 3, 0);    d = make(r2, 3, 1 + (int)(r1 % 2)); }
    else if (sel < 13) { a = make(r1, 3, 3);    d = make(r2, 3, 3); }
    else if (sel == 13) { d = make(r2, 5, 3); const double q = make(r1, 5, 3); a = q * d; if (r1 & 1) a = __longlong_as_double(__double_as_longlong(a) + ((r1 & 2) ? 1 : -1)); }
    else if (sel == 14) { a = sp[r1 % 12]; d = make(r2, 40, 0); }
    else               { a = (r1 & 1) ? sp[(r1 >> 8) % 12] : make(r1, 1000, 0); d = sp[r2 % 12]; }
    const double got = odes_div_seq(a, d, odes_rcp_seq(d));
    const double want = a / d;
    const bool same = __double_as_longlong(got) == __double_as_longlong(want) || (got != got && want != want);
    if (!same) { if (nbad == 0 && tid < 64) printf("mismatch a %a d %a got %a want %a\n", a, d, got, want); nbad++; }
  }
  if (nbad) atomicAdd(bad, nbad);
}
int main(int argc, char **argv)
{
  const unsigned long long per_thread = argc > 1 ? strtoull(argv[1], 0, 10) : 4096;
  unsigned long long *bad, h = 0;
  cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8);
  const int blocks = 148 * 8, threads = 256;
  check<<<blocks, threads>>>(per_thread, bad, 20261020ull);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 2; }
  cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost);
  printf("pairs %llu mismatches %llu\n", per_thread * blocks * threads, h);
  return h ? 1 : 0;
}
