#include "flux_kernels.cuh"
namespace akb {
template __global__ void flux_x1_kernel<RECON_PPM4, RS_HLLD, false>(Geom, EosParams, const double*, const double*, const double*, double*, double*, double*, unsigned, AccArgs);
template __global__ void flux_march_kernel<RECON_PPM4, RS_HLLD, 1, 0>(Geom, EosParams, const double*, const double*, const double*, double*, double*, double*, int, int, AccArgs);
}
