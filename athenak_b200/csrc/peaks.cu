// peaks.cu -- measured FP64 issue peak of the device the engine runs on.
//
// The flux sweeps are bounded by FP64 instruction issue, not by HBM, so their
// roofline denominator has to be measured like the HBM copy peak is: a register-only kernel that
// issues nothing but dependent chains of one FP64 instruction (DFMA, DADD, DMUL, or the DADD/DMUL
// mix an --fmad=false build consists of), with ILP independent chains per thread and a chosen number
// of resident warps per SM sub-partition.  The result is lane-instructions per second
// (= thread-level FP64 instructions retired per second); bench.py reports the flux kernel's own
// FP64 lane-instruction rate against the best figure found here.  Sweeping warps and ILP also
// gives the dependent-issue latency that explains the sweeps' pipe utilisation at 4 warps/SMSP.
#include <cuda_runtime.h>
#include "../../include/athenak_b200.h"

namespace {

template <int OP, int ILP>
__global__ void __launch_bounds__(128) fp64_peak_kernel(double *out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int q = 0; q < ILP; ++q) x[q] = a + (double)(threadIdx.x + q);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
#pragma unroll
      for (int q = 0; q < ILP; ++q) {
        if (OP == 0) x[q] = fma(x[q], b, a);
        else if (OP == 1) x[q] = x[q] + a;
        else if (OP == 2) x[q] = x[q] * b;
        else x[q] = (u & 1) ? x[q] + a : x[q] * b;
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < ILP; ++q) s += x[q];
  if (s == 123.456) out[0] = s;   // never true for the arguments used; keeps the chains alive
}

template <int OP>
cudaError_t launch(int ilp, int grid, double *out, int iters, cudaStream_t st) {
  const double a = 1.0e-9, b = 1.0000001;
  switch (ilp) {
    case 1: fp64_peak_kernel<OP, 1><<<grid, 128, 0, st>>>(out, iters, a, b); break;
    case 2: fp64_peak_kernel<OP, 2><<<grid, 128, 0, st>>>(out, iters, a, b); break;
    case 4: fp64_peak_kernel<OP, 4><<<grid, 128, 0, st>>>(out, iters, a, b); break;
    case 8: fp64_peak_kernel<OP, 8><<<grid, 128, 0, st>>>(out, iters, a, b); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace

// op: 0 DFMA, 1 DADD, 2 DMUL, 3 alternating DADD/DMUL.  warps_per_smsp: resident warps per SM sub-partition
// (1..16; one 128-thread CTA = one warp on each of the four sub-partitions).  Returns lane-instructions/s.
extern "C" int ak_peak_fp64(int op, int ilp, int warps_per_smsp, int iters, double *lane_instr_per_s, double *ms_out) {
  int dev = 0, nsm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1;
  if (warps_per_smsp < 1 || warps_per_smsp > 16 || iters < 1) return 1;
  double *out = nullptr;
  if (cudaMalloc(&out, 8) != cudaSuccess) return 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = nsm * warps_per_smsp;
  cudaError_t err = cudaSuccess;
  float best = 1e30f;
  for (int rep = 0; rep < 4 && err == cudaSuccess; ++rep) {   // first repetition is the warm-up
    cudaEventRecord(e0, 0);
    switch (op) {
      case 0: err = launch<0>(ilp, grid, out, iters, 0); break;
      case 1: err = launch<1>(ilp, grid, out, iters, 0); break;
      case 2: err = launch<2>(ilp, grid, out, iters, 0); break;
      case 3: err = launch<3>(ilp, grid, out, iters, 0); break;
      default: err = cudaErrorInvalidValue;
    }
    cudaEventRecord(e1, 0);
    if (err == cudaSuccess) err = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (err == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (err != cudaSuccess) return 1;
  const double n = (double)grid * 128.0 * (double)iters * 16.0 * (double)ilp;
  if (lane_instr_per_s) *lane_instr_per_s = n / (best * 1e-3);
  if (ms_out) *ms_out = best;
  return 0;
}
