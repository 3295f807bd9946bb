// amr.cu -- adaptive mesh refinement on the device (dynamic refinement): refinement criteria as
// per-block extrema, and the data movement of one AMR step as gathers from the OLD arrays into the NEW arrays.
//
// Reference: src/mesh/refinement_criteria.cpp:177-336, src/mesh/mesh_refinement.cpp:427-656 (RedistAndRefineMeshBlocks),
// :696-1184 (Derefine / Copy / CopyForRefinement / Refine), :1189-1221 (RepairAMRFC).  Bodies and design: amr_body.h.
// The mesh bookkeeping (which block is copied / derefined / created) is the host's (ak_mesh_adapt, host_mesh.cpp).
#include <cuda_runtime.h>
#include <algorithm>
#include <cfloat>
#include <string>
#include <vector>

#include "../../include/athenak_b200.h"
#include "amr_body.h"

namespace {
using namespace akb_smr;

thread_local std::string g_amr_err;
int amr_fail(const std::string &m) { g_amr_err = m; return 1; }

FcDims dims_from_cfg(const ak_mhd_cfg &c) {
  FcDims D;
  const int nx[3] = {c.nx1, c.nx2, c.nx3};
  for (int d = 0; d < 3; ++d) {
    const bool act = (d == 0) || nx[d] > 1;
    const int cnx = std::max(1, nx[d] / 2);
    D.s[d] = act ? c.ng : 0; D.e[d] = D.s[d] + nx[d] - 1;
    D.cs[d] = act ? c.ng : 0; D.ce[d] = D.cs[d] + cnx - 1;
    D.n[d] = act ? nx[d] + 2 * c.ng : 1;
    D.cn[d] = act ? cnx + 2 * c.ng : 1;
  }
  D.multi_d = c.nx2 > 1; D.three_d = c.nx3 > 1; D.ng = c.ng; D.nmb = c.nmb;
  for (int v = 0; v < 3; ++v) {
    D.f[v] = (long)(D.n[0] + (v == 0)) * (D.n[1] + (v == 1)) * (D.n[2] + (v == 2));
    D.cf[v] = (long)(D.cn[0] + (v == 0)) * (D.cn[1] + (v == 1)) * (D.cn[2] + (v == 2));
    D.ef[v] = (long)(D.n[0] + (v != 0)) * (D.n[1] + (v != 1)) * (D.n[2] + (v != 2));
  }
  return D;
}

// blockIdx.y = new block, blockIdx.z = variable; threads over the cells of the block (decode + body: amr_body.h)
__global__ void __launch_bounds__(256)
amr_move_cc_kernel(FcDims D, const AmrBlock *__restrict__ blocks, int nvar, const double *__restrict__ uo, double *__restrict__ un) {
  const int m = blockIdx.y, v = blockIdx.z;
  const AmrBlock B = blocks[m];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < amr_move_cc_count(D); t += gridDim.x * blockDim.x) amr_move_cc_flat(D, B, m, nvar, v, t, uo, un);
}

struct Fc3 { const double *o[3]; double *n[3]; };

// blockIdx.y = new block, blockIdx.z = component
__global__ void __launch_bounds__(256)
amr_move_fc_kernel(FcDims D, const AmrBlock *__restrict__ blocks, Fc3 F) {
  const int m = blockIdx.y, v = blockIdx.z;
  const AmrBlock B = blocks[m];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < (int)D.f[v]; t += gridDim.x * blockDim.x) amr_move_fc_flat(D, B, m, v, t, F.o[v], F.n[v]);
}

// blockIdx.y = new block (only blocks created by refinement do work), blockIdx.z = component; threads over the coarse faces
__global__ void __launch_bounds__(128)
amr_refine_fc_shared_kernel(FcDims D, const AmrBlock *__restrict__ blocks, Fc3 F) {
  const int m = blockIdx.y, v = blockIdx.z;
  const AmrBlock B = blocks[m];
  if (B.kind != AMR_REFINE) return;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < amr_refine_fc_shared_count(D, v); t += gridDim.x * blockDim.x)
    amr_refine_fc_shared_flat(D, B, m, v, t, F.o[v], F.n[v]);
}

// interior faces of every refined coarse cell of the flagged blocks (RefineFC second half, and RepairAMRFC)
__global__ void __launch_bounds__(128)
amr_fc_interior_kernel(FcDims D, const int *__restrict__ blockflag, double *b1, double *b2, double *b3) {
  const int m = blockIdx.y;
  if (!blockflag[m]) return;
  const FcCell d = amr_whole_block(D, m);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < d.n[0] * d.n[1] * d.n[2]; t += gridDim.x * blockDim.x) fc_prolong_interior_elem(D, d, t, b1, b2, b3, true);
}

// per-block max and min of the criterion value over the active cells: out[m][0] = max, out[m][1] = min.  One CTA per block;
// max / min are exact in any order, so the tree reduction equals the reference's serial loop bit for bit
__global__ void __launch_bounds__(256)
amr_extrema_kernel(FcDims D, int method, int nvar, int var, const double *__restrict__ q, double *__restrict__ out) {
  const int m = blockIdx.x;
  const double *a = q + ((long)m * nvar + var) * ((long)D.n[0] * D.n[1] * D.n[2]);
  double vmax = -DBL_MAX, vmin = DBL_MAX;
  for (int t = threadIdx.x; t < amr_active_count(D); t += blockDim.x) {
    const double x = amr_criterion_flat(D, method, a, t);
    vmax = fmax(x, vmax); vmin = fmin(x, vmin);
  }
  __shared__ double smax[256], smin[256];
  smax[threadIdx.x] = vmax; smin[threadIdx.x] = vmin;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) { smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + s]); smin[threadIdx.x] = fmin(smin[threadIdx.x], smin[threadIdx.x + s]); }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[2 * m] = smax[0]; out[2 * m + 1] = smin[0]; }
}

struct Scratch { void *p = nullptr; size_t n = 0; };
thread_local Scratch g_scr;
void *scratch(size_t bytes) {
  if (g_scr.n < bytes) { cudaFree(g_scr.p); g_scr.p = nullptr; g_scr.n = 0; if (cudaMalloc(&g_scr.p, bytes) != cudaSuccess) return nullptr; g_scr.n = bytes; }
  return g_scr.p;
}

}  // namespace

extern "C" {

const char *ak_amr_last_error(void) { return g_amr_err.c_str(); }

/* per NEW block: what it is made of (host logic shared by the device path and the emulation harness of the tests) */
int ak_amr_plan(const ak_mhd_cfg *cfg_new, int old_nmb, const int *newtoold, const int *old_flag, const int *lloc_new, int *plan /*[new_nmb][5]*/) {
  if (!cfg_new || !newtoold || !old_flag || !lloc_new || !plan) return amr_fail("ak_amr_plan: null argument");
  const int nx[3] = {cfg_new->nx1, cfg_new->nx2, cfg_new->nx3};
  for (int m = 0; m < cfg_new->nmb; ++m) {
    const int o = newtoold[m];
    if (o < 0 || o >= old_nmb) return amr_fail("ak_amr_plan: new -> old map out of range");
    int *P = plan + 5 * m;
    P[0] = old_flag[o] > 0 ? AMR_REFINE : old_flag[o] < -1 ? AMR_DEREF : AMR_COPY;
    P[1] = o;
    for (int d = 0; d < 3; ++d) P[2 + d] = (P[0] == AMR_REFINE && (d == 0 || nx[d] > 1)) ? (lloc_new[4 * m + d] & 1) * (nx[d] / 2) : 0;
  }
  return 0;
}

/* One AMR step's data movement: every element of the NEW arrays gathered from the OLD ones (unchanged blocks copied whole,
 * derefined blocks restricted from their old siblings, blocks created by refinement prolongated from their old parent,
 * face fields with the divergence-preserving operators).  plan = ak_amr_plan's table on the HOST.  b_* may be NULL (hydro).
 * Ghost zones of changed blocks are left to the ghost exchange the host runs next (InitBoundaryValuesAndPrimitives). */
int ak_amr_move(const ak_mhd_cfg *cfg_new, const int *plan, const double *u_old, double *u_new, int nvar,
                const double *b1_old, const double *b2_old, const double *b3_old, double *b1_new, double *b2_new, double *b3_new, ak_stream_t s) {
  if (!cfg_new || !plan || !u_old || !u_new) return amr_fail("ak_amr_move: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(s);
  const FcDims D = dims_from_cfg(*cfg_new);
  const int nmb = cfg_new->nmb;
  if (nmb < 1 || nmb > 65535 || nvar < 1 || nvar > 65535) return amr_fail("ak_amr_move: nmb and nvar must be in 1..65535 (grid dimensions y, z)");
  static_assert(sizeof(AmrBlock) == 5 * sizeof(int), "plan layout");
  AmrBlock *d_plan = static_cast<AmrBlock *>(scratch(sizeof(AmrBlock) * nmb + sizeof(int) * nmb));
  if (!d_plan) return amr_fail("ak_amr_move: out of device memory");
  int *d_flag = reinterpret_cast<int *>(d_plan + nmb);
  std::vector<int> refined(nmb);
  for (int m = 0; m < nmb; ++m) refined[m] = plan[5 * m] == AMR_REFINE;
  cudaMemcpyAsync(d_plan, plan, sizeof(AmrBlock) * nmb, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_flag, refined.data(), sizeof(int) * nmb, cudaMemcpyHostToDevice, st);
  cudaStreamSynchronize(st);          // `refined` is a host temporary
  const int ncell = D.n[0] * D.n[1] * D.n[2];
  amr_move_cc_kernel<<<dim3((unsigned)std::min((ncell + 255) / 256, 64), nmb, nvar), 256, 0, st>>>(D, d_plan, nvar, u_old, u_new);
  if (b1_old) {
    Fc3 F;
    F.o[0] = b1_old; F.o[1] = b2_old; F.o[2] = b3_old; F.n[0] = b1_new; F.n[1] = b2_new; F.n[2] = b3_new;
    const int nface = (int)std::max(D.f[0], std::max(D.f[1], D.f[2]));
    amr_move_fc_kernel<<<dim3((unsigned)std::min((nface + 255) / 256, 64), nmb, 3), 256, 0, st>>>(D, d_plan, F);
    const int ncf = (int)((long)(D.ce[0] - D.cs[0] + 2) * (D.ce[1] - D.cs[1] + 2) * (D.ce[2] - D.cs[2] + 2));
    amr_refine_fc_shared_kernel<<<dim3((unsigned)std::min((ncf + 127) / 128, 64), nmb, 3), 128, 0, st>>>(D, d_plan, F);
    amr_fc_interior_kernel<<<dim3((unsigned)std::min((ncf + 127) / 128, 64), nmb), 128, 0, st>>>(D, d_flag, b1_new, b2_new, b3_new);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return amr_fail(std::string("ak_amr_move: ") + cudaGetErrorString(e));
  return 0;
}

/* RepairAMRFC (mesh_refinement.cpp:1189-1221): recompute the interior faces of the blocks that were refined or derefined
 * (blockflag on the HOST, [nmb]) after the first ghost exchange has finalised their exterior faces */
int ak_amr_repair_fc(const ak_mhd_cfg *cfg, const int *blockflag, double *b1, double *b2, double *b3, ak_stream_t s) {
  if (!cfg || !blockflag || !b1 || !b2 || !b3) return amr_fail("ak_amr_repair_fc: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(s);
  const FcDims D = dims_from_cfg(*cfg);
  if (cfg->nmb < 1 || cfg->nmb > 65535) return amr_fail("ak_amr_repair_fc: nmb must be in 1..65535 (grid dimension y)");
  int *d_flag = static_cast<int *>(scratch(sizeof(int) * cfg->nmb));
  if (!d_flag) return amr_fail("ak_amr_repair_fc: out of device memory");
  cudaMemcpyAsync(d_flag, blockflag, sizeof(int) * cfg->nmb, cudaMemcpyHostToDevice, st);
  cudaStreamSynchronize(st);
  const int ncf = (int)((long)(D.ce[0] - D.cs[0] + 1) * (D.ce[1] - D.cs[1] + 1) * (D.ce[2] - D.cs[2] + 1));
  amr_fc_interior_kernel<<<dim3((unsigned)std::min((ncf + 127) / 128, 64), cfg->nmb), 128, 0, st>>>(D, d_flag, b1, b2, b3);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return amr_fail(std::string("ak_amr_repair_fc: ") + cudaGetErrorString(e));
  return 0;
}

/* One <amr_criterion> block (refinement_criteria.cpp:177-386): method 0 min_max, 1 slope, 2 second_deriv, 3 location.
 * q = cell-centred DEVICE array [nmb][nvar][n3][n2][n1] (e.g. w0), var = the variable in it (IDN for *_d); flag = HOST
 * [nmb], updated as the reference updates refine_flag (a block flagged for refinement by an earlier criterion is never
 * turned into a derefinement).  loc4 = location_x1/x2/x3/rad, mb_size = HOST [nmb][9] (both only for method 3).
 * Synchronises the stream (the verdict goes to the host's tree update). */
int ak_amr_criterion(const ak_mhd_cfg *cfg, int method, const double *q, int nvar, int var, double value_min, double value_max,
                     const double *loc4, const double *mb_size, int *flag, ak_stream_t s) {
  if (!cfg || !flag) return amr_fail("ak_amr_criterion: null argument");
  const int nmb = cfg->nmb;
  const bool multi_d = cfg->nx2 > 1, three_d = cfg->nx3 > 1;
  if (method == 3) {
    if (!loc4 || !mb_size) return amr_fail("ak_amr_criterion: location needs loc4 and mb_size");
    for (int m = 0; m < nmb; ++m) {
      const double *z = mb_size + 9 * m;
      bool in[3];
      for (int d = 0; d < 3; ++d) {
        const double lo = z[d], hi = z[3 + d], x = loc4[d], r = loc4[3];
        in[d] = ((lo < (x + r)) && (lo > (x - r))) || ((hi < (x + r)) && (hi > (x - r))) || ((hi > (x + r)) && (lo < (x - r)));
      }
      if (in[0] && (!multi_d || in[1]) && (!three_d || in[2])) flag[m] = 1;
    }
    return 0;
  }
  if (method < 0 || method > 2 || !q) return amr_fail("ak_amr_criterion: unknown method / null data");
  cudaStream_t st = static_cast<cudaStream_t>(s);
  const FcDims D = dims_from_cfg(*cfg);
  double *d_out = static_cast<double *>(scratch(sizeof(double) * 2 * nmb));
  if (!d_out) return amr_fail("ak_amr_criterion: out of device memory");
  amr_extrema_kernel<<<nmb, 256, 0, st>>>(D, method, nvar, var, q, d_out);
  std::vector<double> out(2 * (size_t)nmb);
  cudaError_t e = cudaMemcpyAsync(out.data(), d_out, sizeof(double) * 2 * nmb, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return amr_fail(std::string("ak_amr_criterion: ") + cudaGetErrorString(e));
  for (int m = 0; m < nmb; ++m) {
    const double qmax = out[2 * m], qmin = out[2 * m + 1];
    if (value_max < (double)FLT_MAX) {
      if (qmax > value_max) flag[m] = 1;
      if ((qmax < value_max) && (flag[m] == 0)) flag[m] = -1;
    }
    if (method == 0 && value_min > -(double)FLT_MAX) {
      if (qmin < value_min) flag[m] = 1;
      if ((qmin > value_min) && (flag[m] == 0)) flag[m] = -1;
    }
  }
  return 0;
}

}  // extern "C"
