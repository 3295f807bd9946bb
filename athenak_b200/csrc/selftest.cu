// selftest.cu -- on-device check that the slow-path-free drcp/ddiv/dsqrt (mhd_device.cuh) return
// the same bits as the built-in IEEE operators over random normal-range operands.
#include <cuda_runtime.h>
#include "../../include/athenak_b200.h"
#include "mhd_device.cuh"

namespace {

__device__ __forceinline__ unsigned long long mix(unsigned long long z) {  // splitmix64
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
// random sign/mantissa, exponent uniform in [-span, span]
__device__ __forceinline__ double rnd(unsigned long long &st, int span) {
  st = mix(st);
  const unsigned long long m = st & 0x800fffffffffffffULL;
  st = mix(st);
  const unsigned long long e = 1023ULL - span + (st % (2ULL * span + 1ULL));
  return __longlong_as_double((long long)(m | (e << 52)));
}

__global__ void selftest_kernel(long n, unsigned long long seed, int span, unsigned long long *bad) {
  unsigned long long st = seed ^ (0x632be59bd9b4e019ULL * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1));
  unsigned long long b0 = 0, b1 = 0, b2 = 0;
  for (long q = blockIdx.x * (long)blockDim.x + threadIdx.x; q < n; q += (long)gridDim.x * blockDim.x) {
    double a = rnd(st, span), b = rnd(st, span);
    if ((q & 63) == 0) a = 0.0;                       // exact-zero numerators are common on this path
    volatile double vb = b, va = a;                   // keep the reference operators un-folded
    const double r_ref = 1.0 / vb, q_ref = va / vb, s_ref = sqrt(fabs(vb));
    b0 += (__double_as_longlong(akb::drcp(b)) != __double_as_longlong(r_ref));
    b1 += (__double_as_longlong(akb::ddiv(a, b)) != __double_as_longlong(q_ref));
    b2 += (__double_as_longlong(akb::dsqrt(fabs(b))) != __double_as_longlong(s_ref));
  }
  if (b0) atomicAdd(bad + 0, b0);
  if (b1) atomicAdd(bad + 1, b1);
  if (b2) atomicAdd(bad + 2, b2);
}

}  // namespace

extern "C" int ak_selftest_divsqrt(long n, unsigned long long seed, int exponent_span, long *mismatches) {
  unsigned long long *d = nullptr;
  if (cudaMalloc(&d, 3 * sizeof(unsigned long long)) != cudaSuccess) return 1;
  cudaMemset(d, 0, 3 * sizeof(unsigned long long));
  selftest_kernel<<<148 * 8, 256>>>(n, seed, exponent_span, d);
  unsigned long long h[3];
  cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return 1;
  for (int q = 0; q < 3; ++q) mismatches[q] = (long)h[q];
  return 0;
}
