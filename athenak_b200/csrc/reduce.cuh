// reduce.cuh -- block-level min reduction used by the new-time-step kernels (MHD and hydro).
#pragma once
#include "mhd_kernels.h"
#include "mhd_device.cuh"

namespace akb {

// identity of the min-reduction = (double)FLT_MAX, the reference's initial value (mhd_newdt.cpp:40-42)
#define DT_IDENT 3.4028234663852886e+38
__device__ __forceinline__ void atomic_min_pos(unsigned long long *addr, double v) {
  atomicMin(addr, (unsigned long long)__double_as_longlong(v));  // v > 0: bit order == value order
}

// Block-wide min of three positive doubles, then ONE atomic per block and value into one of
// DT_SLOTS slots (chosen by block index) to keep same-address atomic traffic low; the host
// takes the min over the slots.  All threads of the block must call it.
__device__ __forceinline__ void block_min3(double d1, double d2, double d3, unsigned long long *__restrict__ dtmin) {
  __shared__ double sm[3][32];    // up to 1024 threads per block
  for (int s = 16; s > 0; s >>= 1) {
    d1 = dmin(d1, __shfl_xor_sync(0xffffffffu, d1, s));
    d2 = dmin(d2, __shfl_xor_sync(0xffffffffu, d2, s));
    d3 = dmin(d3, __shfl_xor_sync(0xffffffffu, d3, s));
  }
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if ((threadIdx.x & 31) == 0) { sm[0][w] = d1; sm[1][w] = d2; sm[2][w] = d3; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = sm[threadIdx.x][0];
    for (int q = 1; q < nw; ++q) v = dmin(v, sm[threadIdx.x][q]);
    atomic_min_pos(dtmin + 3 * (blockIdx.x % DT_SLOTS) + threadIdx.x, v);
  }
}

// floor counters (dfloor, efloor, tfloor): warp-level sum, one atomic per warp and only when needed
__device__ __forceinline__ void warp_count3(int nd, int ne, int nt, int *__restrict__ counters) {
  const unsigned any = __ballot_sync(0xffffffffu, nd | ne | nt);
  if (any) {
    for (int s = 16; s > 0; s >>= 1) {
      nd += __shfl_xor_sync(0xffffffffu, nd, s);
      ne += __shfl_xor_sync(0xffffffffu, ne, s);
      nt += __shfl_xor_sync(0xffffffffu, nt, s);
    }
    if ((threadIdx.x & 31) == 0) {
      if (nd) atomicAdd(counters + 0, nd);
      if (ne) atomicAdd(counters + 1, ne);
      if (nt) atomicAdd(counters + 2, nt);
    }
  }
}

}  // namespace akb
