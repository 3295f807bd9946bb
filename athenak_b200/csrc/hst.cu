// hst.cu -- history-output sums as a device reduction (reference
// src/outputs/history.cpp:78-160 LoadHydroHistoryData, :272-372 LoadMHDHistoryData).
// Per active cell: vol*{d, m1, m2, m3, E}, vol*0.5*m_d^2/d (3 kinetic energies) and, for MHD,
// vol*0.25*(b_f(i)^2 + b_f(i+1)^2) (3 magnetic energies).  Two deterministic passes: fixed-shape
// tree per block into a partial array, then one block adds the partials in index order, so the
// result does not depend on scheduling (it differs from the reference's sequential Kokkos::Serial
// sum only by FP64 reassociation: compare with a relative tolerance, not bitwise).
#include "mhd_kernels.h"
#include "mhd_device.cuh"

namespace akb {

#define HST_NV 11
#define HST_THREADS 256

__global__ void __launch_bounds__(HST_THREADS)
hst_partial_kernel(Geom g, bool mhd, const double *__restrict__ u0, const double *__restrict__ b1f,
                   const double *__restrict__ b2f, const double *__restrict__ b3f,
                   const double *__restrict__ mb_size, double *__restrict__ partial) {
  __shared__ double sm[HST_NV][HST_THREADS / 32];
  const long total = (long)g.nmb * g.nx3 * g.nx2 * g.nx1;
  double acc[HST_NV];
#pragma unroll
  for (int n = 0; n < HST_NV; ++n) acc[n] = 0.0;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    long r = t;
    const int i = g.is + (int)(r % g.nx1); r /= g.nx1;
    const int j = g.js + (int)(r % g.nx2); r /= g.nx2;
    const int k = g.ks + (int)(r % g.nx3);
    const int m = (int)(r / g.nx3);
    const double vol = mb_size[m * 9 + 6] * mb_size[m * 9 + 7] * mb_size[m * 9 + 8];
    const long o = ((long)k * g.n2 + j) * g.n1 + i;
    const double *u = u0 + (long)m * g.nvar * g.cc + o;
    const double d = u[0], m1 = u[g.cc], m2 = u[2 * g.cc], m3 = u[3 * g.cc];
    const double e = (g.nfl == 5) ? u[4 * g.cc] : 0.0;   // isothermal: no energy variable (history.cpp:140-142)
    acc[0] += vol * d; acc[1] += vol * m1; acc[2] += vol * m2; acc[3] += vol * m3; acc[4] += vol * e;
    acc[5] += vol * 0.5 * sqr(m1) / d;
    acc[6] += vol * 0.5 * sqr(m2) / d;
    acc[7] += vol * 0.5 * sqr(m3) / d;
    if (mhd) {
      const double *p1 = b1f + (long)m * g.f1 + ((long)k * g.n2 + j) * (g.n1 + 1) + i;
      const double *p2 = b2f + (long)m * g.f2 + ((long)k * (g.n2 + 1) + j) * g.n1 + i;
      const double *p3 = b3f + (long)m * g.f3 + o;
      acc[8] += vol * 0.25 * (sqr(p1[1]) + sqr(p1[0]));
      acc[9] += vol * 0.25 * (sqr(p2[g.n1]) + sqr(p2[0]));
      acc[10] += vol * 0.25 * (sqr(p3[(long)g.n1 * g.n2]) + sqr(p3[0]));
    }
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int n = 0; n < HST_NV; ++n) {
    double v = acc[n];
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == 0) sm[n][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < HST_NV) {
    double v = 0.0;
    for (int q = 0; q < HST_THREADS / 32; ++q) v += sm[threadIdx.x][q];
    partial[(long)blockIdx.x * HST_NV + threadIdx.x] = v;
  }
}

// skip_e: isothermal runs have no tot-E column, the later columns move up by one (history.cpp:84-88)
__global__ void hst_final_kernel(const double *__restrict__ partial, int nblocks, double *__restrict__ out, int skip_e) {
  if (threadIdx.x < HST_NV) {
    double v = 0.0;
    for (int b = 0; b < nblocks; ++b) v += partial[(long)b * HST_NV + threadIdx.x];
    if (!skip_e) out[threadIdx.x] = v;
    else if (threadIdx.x != 4) out[threadIdx.x < 4 ? threadIdx.x : threadIdx.x - 1] = v;
  }
}

// scratch: at least (nblocks+1)*HST_NV doubles; result in scratch[nblocks*HST_NV ..]
cudaError_t launch_hst_sums(const Geom &g, const Arrays &a, bool mhd, double *scratch, int nblocks, cudaStream_t st) {
  hst_partial_kernel<<<nblocks, HST_THREADS, 0, st>>>(g, mhd, a.u0, a.b0x1f, a.b0x2f, a.b0x3f, a.mb_size, scratch);
  hst_final_kernel<<<1, 32, 0, st>>>(scratch, nblocks, scratch + (long)nblocks * HST_NV, g.nfl == 4 ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace akb
