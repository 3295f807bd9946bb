// selftest.cu -- on-device check that drcp/ddiv/dsqrt (mhd_device.cuh: nvcc's fast paths + a zero-aware
// range guard) return the same bits as the built-in IEEE operators.  exponent_span <= 1022: random
// normal-range operands with exponents in [-span, span] (every 64th numerator an exact zero).
// exponent_span > 1022: the WHOLE double range -- biased exponent field uniform in [0, 0x7ff], so
// subnormals, the denormal floor below 2^-969, values next to overflow, infinities and NaNs are all
// drawn, plus exact +-0 -- in every operand position.  NaN results compare equal to NaN results.
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
  if (span > 1022) {   // whole range: exponent field 0 (zero / subnormal) .. 0x7ff (inf / nan)
    unsigned long long e = st % 0x800ULL, mm = m;
    st = mix(st);
    if ((st & 15) == 0) mm &= 0x8000000000000000ULL;                 // exact powers of two, +-0, +-inf
    if ((st & 0xf0) == 0) mm &= 0x80000000000000ffULL;               // subnormals at the very bottom
    return __longlong_as_double((long long)(mm | (e << 52)));
  }
  const unsigned long long e = 1023ULL - span + (st % (2ULL * span + 1ULL));
  return __longlong_as_double((long long)(m | (e << 52)));
}

__global__ void selftest_kernel(long n, unsigned long long seed, int span, unsigned long long *bad) {
  unsigned long long st = seed ^ (0x632be59bd9b4e019ULL * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1));
  unsigned long long b0 = 0, b1 = 0, b2 = 0, b3 = 0, b4 = 0;
  for (long q = blockIdx.x * (long)blockDim.x + threadIdx.x; q < n; q += (long)gridDim.x * blockDim.x) {
    double a = rnd(st, span), b = rnd(st, span);
    if ((q & 63) == 0) a = 0.0;                       // exact-zero numerators are common on this path
    volatile double vb = b, va = a;                   // keep the reference operators un-folded
    const double sb = (span > 1022) ? b : fabs(b);    // whole range: negative arguments too (NaN on both sides)
    volatile double vsb = sb;
    const double r_ref = 1.0 / vb, q_ref = va / vb, s_ref = sqrt(vsb);
    auto differ = [](double x, double y) { return (x != x || y != y) ? ((x != x) != (y != y)) : (__double_as_longlong(x) != __double_as_longlong(y)); };
    auto differ_v = [&](double x, double y) { return (x == 0.0 && y == 0.0) ? false : differ(x, y); };   // deferred form: the sign of a zero quotient is not kept
    b0 += differ(akb::drcp(b), r_ref);
    b1 += differ(akb::ddiv(a, b), q_ref);
    b2 += differ(akb::dsqrt(sb), s_ref);
    // AKB_COLD: the operators of the repeat side of the deferred guard (scaled exact division first, the compiler's operator for
    // what that form flags) must return the IEEE bits for every operand too, signs of zeros included
    b0 += differ(akb::drcp(b, AKB_COLD), r_ref);
    b1 += differ(akb::ddiv(a, b, AKB_COLD), q_ref);
    b2 += differ(akb::dsqrt(sb, AKB_COLD), s_ref);
    {  // div_by: launch-uniform divisors of the path (12, 6 and awkward cell sizes), numerator over the drawn range
      const double ys[6] = {12.0, 6.0, 0.0050781250000000002, 1.0 / 3.0, 0.0129999999999999, 4.19921875};
      volatile double vy = ys[q % 6];
      const double y = vy, ry = 1.0 / vy, d_ref = va / vy;
      b3 += differ(akb::div_by(a, y, ry), d_ref);
      b3 += differ(akb::div_by(a, y, ry, AKB_COLD), d_ref);
      // deferred form of the guard (flux sweeps, fused stage tail): whenever the flag stays clear the fast result must be the
      // IEEE one (up to the sign of a zero); a set flag sends the caller to the guarded operators checked above
      bool o1 = false, o2 = false, o3 = false, o4 = false;
      const double r1 = akb::drcp(b, &o1), r2 = akb::ddiv(a, b, &o2), r3 = akb::dsqrt(sb, &o3), r4 = akb::div_by(a, y, ry, &o4);
      b4 += (!o1 && differ_v(r1, r_ref)) + (!o2 && differ_v(r2, q_ref)) + (!o3 && differ_v(r3, s_ref)) + (!o4 && differ_v(r4, d_ref));
      // ddiv_x / div_by_x: exact for EVERY numerator (scaled fast path + subnormal tie repair); the flag may only be raised by an
      // odd divisor, an overflow or a NaN -- never by a zero, subnormal or tiny numerator over a normal divisor of moderate size
      bool o5 = false, o6 = false;
      const double r5 = akb::ddiv_x(a, b, &o5), r6 = akb::div_by_x(a, y, ry, &o6);
      b4 += (!o5 && differ_v(r5, q_ref)) + (!o6 && differ_v(r6, d_ref));
      const double ab = fabs(b);
      const bool tame = (a == a) && fabs(a) < 1.0e-200 && ab > 1.0e-60 && ab < 1.0e60;
      b4 += (tame && o5) + (tame && o6);
    }
  }
  if (b0) atomicAdd(bad + 0, b0);
  if (b1) atomicAdd(bad + 1, b1);
  if (b2) atomicAdd(bad + 2, b2);
  if (b3) atomicAdd(bad + 3, b3);
  if (b4) atomicAdd(bad + 4, b4);
}

}  // namespace

extern "C" int ak_selftest_divsqrt(long n, unsigned long long seed, int exponent_span, long *mismatches) {
  unsigned long long *d = nullptr;
  if (cudaMalloc(&d, 5 * sizeof(unsigned long long)) != cudaSuccess) return 1;
  cudaMemset(d, 0, 5 * sizeof(unsigned long long));
  selftest_kernel<<<148 * 8, 256>>>(n, seed, exponent_span, d);
  unsigned long long h[5];
  cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return 1;
  for (int q = 0; q < 5; ++q) mismatches[q] = (long)h[q];
  return 0;
}
