// bcs.cu -- task 26 MHD::ApplyPhysicalBCs (src/mhd/mhd_tasks.cpp:504-520): physical boundary
// conditions on the conserved variables (src/bvals/physics/hydro_bcs.cpp:68-335) and on the
// face-centred field (src/bvals/physics/bfield_bcs.cpp:67-358) of the blocks that touch the edge
// of the mesh.
//
// One launch per boundary DIRECTION covers both sides, all four arrays (u0, x1f, x2f, x3f) and the
// full transverse extent, one thread per ghost element; the three directions stay ordered on the
// stream (x1, x2, x3) because the later passes read the transverse ghosts the earlier ones wrote
// -- that ordering is what fills the edge and corner ghosts in the reference.  Within a pass
// every source is an active index along the pass direction, so elements are independent.
// Pure HBM-bound gather/scatter; the ghost shell is ~2*ng/nx of a block per direction.
#include <cuda_runtime.h>
#include <string>
#include <vector>

#include "../../include/athenak_b200.h"

namespace {

struct BcPlan {
  ak_mhd_cfg cfg;
  int mesh_bcs[6];
  int *d_mb_bcs = nullptr;   // [nmb][6]
  bool any = false;          // false: strictly periodic (mesh.cpp:47,92,131,158)
  std::string err;
};

thread_local std::string g_bc_err;

struct BcArray {
  double *p;
  int ext[3];      // x1, x2, x3 extents of one variable of one block
  int nvar;
  int kind;        // -1 cell-centred conserved variables, 0/1/2 = x1f/x2f/x3f
};
struct BcArgs {
  BcArray a[4];
  long start[5];   // prefix sums of ghost elements per array for this direction
};

// thread -> (array, block, variable, transverse pair, side, ghost layer)
template <int DIR>
__global__ void __launch_bounds__(256) bc_kernel(const BcArgs A, const int *__restrict__ mb_bcs, int nmb, int ng,
                                                 int s, int e) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.start[4]) return;
  int q = 0;
  while (t >= A.start[q + 1]) ++q;
  const BcArray a = A.a[q];
  long r = t - A.start[q];
  // fastest-varying coordinate first: x1 for the x2/x3 passes (coalesced rows), the ghost layer
  // for the x1 pass
  int c[3], g, side;
  constexpr int TA = (DIR + 1) % 3, TB = (DIR + 2) % 3;
  if (DIR == 0) {
    g = (int)(r % ng); r /= ng;
    side = (int)(r & 1); r >>= 1;
    c[1] = (int)(r % a.ext[1]); r /= a.ext[1];
    c[2] = (int)(r % a.ext[2]); r /= a.ext[2];
  } else {
    c[0] = (int)(r % a.ext[0]); r /= a.ext[0];
    g = (int)(r % ng); r /= ng;
    side = (int)(r & 1); r >>= 1;
    constexpr int T2 = (DIR == 1) ? 2 : 1;
    c[T2] = (int)(r % a.ext[T2]); r /= a.ext[T2];
  }
  (void)TA; (void)TB;
  const int n = (int)(r % a.nvar);
  const int m = (int)(r / a.nvar);
  const int flag = mb_bcs[m * 6 + 2 * DIR + side];
  if (flag != AK_BC_REFLECT && flag != AK_BC_OUTFLOW && flag != AK_BC_DIODE && flag != AK_BC_VACUUM) return;

  const int face_like = (a.kind == DIR);       // the field component normal to this boundary
  int dst, mir, edge;
  if (side == 0) { dst = s - g - 1; mir = s + g + face_like; edge = s; }
  else { dst = e + g + 1 + face_like; mir = e - g; edge = e + face_like; }

  const long st[3] = {1, a.ext[0], (long)a.ext[0] * a.ext[1]};
  c[DIR] = 0;
  double *line = a.p + ((long)m * a.nvar + n) * st[2] * a.ext[2] + c[0] * st[0] + c[1] * st[1] + c[2] * st[2];
  const bool normal = (a.kind < 0) ? (n == 1 + DIR) : (face_like != 0);   // IVX+dir / normal B
  double v;
  if (flag == AK_BC_REFLECT) {
    v = line[mir * st[DIR]];
    if (normal) v = -v;
  } else if (flag == AK_BC_VACUUM && a.kind < 0) {
    v = 0.0;
  } else {
    v = line[edge * st[DIR]];
    if (flag == AK_BC_DIODE && a.kind < 0 && normal) v = (side == 0) ? fmin(0.0, v) : fmax(0.0, v);
  }
  line[dst * st[DIR]] = v;
}

int launch_dir(BcPlan *P, int dir, const BcArray *arr, int narr, cudaStream_t st) {
  const ak_mhd_cfg &c = P->cfg;
  BcArgs A;
  long tot = 0;
  for (int q = 0; q < 4; ++q) {
    A.start[q] = tot;
    if (q < narr) {
      A.a[q] = arr[q];
      const int ta = (dir + 1) % 3, tb = (dir + 2) % 3;
      tot += (long)c.nmb * arr[q].nvar * arr[q].ext[ta] * arr[q].ext[tb] * 2 * c.ng;
    } else {
      A.a[q] = arr[0];
    }
  }
  A.start[4] = tot;
  for (int q = narr; q < 4; ++q) A.start[q] = tot;
  const int nx[3] = {c.nx1, c.nx2, c.nx3};
  const int s = c.ng, e = c.ng + nx[dir] - 1;
  const unsigned nb = (unsigned)((tot + 255) / 256);
  if (dir == 0) bc_kernel<0><<<nb, 256, 0, st>>>(A, P->d_mb_bcs, c.nmb, c.ng, s, e);
  else if (dir == 1) bc_kernel<1><<<nb, 256, 0, st>>>(A, P->d_mb_bcs, c.nmb, c.ng, s, e);
  else bc_kernel<2><<<nb, 256, 0, st>>>(A, P->d_mb_bcs, c.nmb, c.ng, s, e);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { P->err = std::string("bc_kernel: ") + cudaGetErrorString(err); return 1; }
  return 0;
}

BcArray make_array(const ak_mhd_cfg &c, double *p, int nvar, int kind) {
  BcArray a;
  a.p = p; a.nvar = nvar; a.kind = kind;
  a.ext[0] = c.nx1 + 2 * c.ng + (kind == 0);
  a.ext[1] = (c.nx2 > 1 ? c.nx2 + 2 * c.ng : 1) + (kind == 1);   // inactive dimensions: extent 1, no ghosts
  a.ext[2] = (c.nx3 > 1 ? c.nx3 + 2 * c.ng : 1) + (kind == 2);
  return a;
}

// the reference gates each pass on the MESH flag of the inner face (hydro_bcs.cpp:80-81,167,252;
// bfield_bcs.cpp:78,169,262)
bool pass_active(const BcPlan *P, int dir, bool hydro) {
  const int f = P->mesh_bcs[2 * dir];
  if ((dir == 1 && P->cfg.nx2 <= 1) || (dir == 2 && P->cfg.nx3 <= 1)) return false;   // hydro_bcs.cpp:166,251 (multi_d / three_d)
  if (f == AK_BC_PERIODIC) return false;
  if (hydro && dir == 0 && f == AK_BC_SHEAR_PERIODIC) return false;
  return true;
}

}  // namespace

extern "C" {

const char *ak_bcs_last_error(void *bcs) {
  return bcs ? static_cast<BcPlan *>(bcs)->err.c_str() : g_bc_err.c_str();
}

int ak_bcs_create(const ak_mhd_cfg *cfg, const int *mesh_bcs, const int *mb_bcs, void **bcs) {
  if (!cfg || !mesh_bcs || !mb_bcs || !bcs) { g_bc_err = "ak_bcs_create: null argument"; return 1; }
  if (cfg->nmb < 1 || cfg->ng < 1) { g_bc_err = "ak_bcs_create: bad cfg"; return 1; }
  BcPlan *P = new BcPlan;
  P->cfg = *cfg;
  for (int q = 0; q < 6; ++q) {
    P->mesh_bcs[q] = mesh_bcs[q];
    if (mesh_bcs[q] != AK_BC_PERIODIC) P->any = true;
  }
  for (int q = 0; q < 6 * cfg->nmb; ++q) {
    const int f = mb_bcs[q];
    // inflow needs the problem generator's u_in/b_in tables, user BCs its callback, shear_periodic
    // the shearing-box remap (bvals.hpp:231-232, pgen.hpp): none is on this path
    if (f == AK_BC_INFLOW || f == AK_BC_USER || f == AK_BC_SHEAR_PERIODIC || f < AK_BC_UNDEF || f > AK_BC_VACUUM) {
      g_bc_err = "ak_bcs_create: boundary flag " + std::to_string(f) +
                 " is not supported (reflect, outflow, diode, vacuum, periodic, block are)";
      delete P;
      return 1;
    }
  }
  if (P->any) {
    cudaError_t e = cudaMalloc(&P->d_mb_bcs, sizeof(int) * 6 * cfg->nmb);
    if (e == cudaSuccess) e = cudaMemcpy(P->d_mb_bcs, mb_bcs, sizeof(int) * 6 * cfg->nmb, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      g_bc_err = std::string("ak_bcs_create: ") + cudaGetErrorString(e);
      if (P->d_mb_bcs) cudaFree(P->d_mb_bcs);
      delete P;
      return 1;
    }
  }
  *bcs = P;
  return 0;
}

int ak_bcs_destroy(void *bcs) {
  BcPlan *P = static_cast<BcPlan *>(bcs);
  if (!P) return 0;
  if (P->d_mb_bcs) cudaFree(P->d_mb_bcs);
  delete P;
  return 0;
}

int ak_bcs_strictly_periodic(void *bcs) { return static_cast<BcPlan *>(bcs)->any ? 0 : 1; }

int ak_bcs_hydro(void *bcs, double *u0, int nvar, ak_stream_t s) {
  BcPlan *P = static_cast<BcPlan *>(bcs);
  if (!P) { g_bc_err = "ak_bcs_hydro: null plan"; return 1; }
  if (!P->any) return 0;
  if (!u0 || nvar < 4) { P->err = "ak_bcs_hydro: null array or nvar < 4"; return 1; }
  BcArray a = make_array(P->cfg, u0, nvar, -1);
  for (int d = 0; d < 3; ++d)
    if (pass_active(P, d, true) && launch_dir(P, d, &a, 1, static_cast<cudaStream_t>(s))) return 1;
  return 0;
}

int ak_bcs_bfield(void *bcs, double *x1f, double *x2f, double *x3f, ak_stream_t s) {
  BcPlan *P = static_cast<BcPlan *>(bcs);
  if (!P) { g_bc_err = "ak_bcs_bfield: null plan"; return 1; }
  if (!P->any) return 0;
  if (!x1f || !x2f || !x3f) { P->err = "ak_bcs_bfield: null array"; return 1; }
  BcArray a[3] = {make_array(P->cfg, x1f, 1, 0), make_array(P->cfg, x2f, 1, 1), make_array(P->cfg, x3f, 1, 2)};
  for (int d = 0; d < 3; ++d)
    if (pass_active(P, d, false) && launch_dir(P, d, a, 3, static_cast<cudaStream_t>(s))) return 1;
  return 0;
}

// HydroBCs + BFieldBCs in one launch per direction (the two never read each other's arrays)
int ak_bcs_apply(void *bcs, double *u0, int nvar, double *x1f, double *x2f, double *x3f, ak_stream_t s) {
  BcPlan *P = static_cast<BcPlan *>(bcs);
  if (!P) { g_bc_err = "ak_bcs_apply: null plan"; return 1; }
  if (!P->any) return 0;
  if (!u0 || !x1f || !x2f || !x3f || nvar < 4) { P->err = "ak_bcs_apply: null array or nvar < 4"; return 1; }
  BcArray a[4] = {make_array(P->cfg, u0, nvar, -1), make_array(P->cfg, x1f, 1, 0), make_array(P->cfg, x2f, 1, 1),
                  make_array(P->cfg, x3f, 1, 2)};
  for (int d = 0; d < 3; ++d) {
    const bool hy = pass_active(P, d, true), bf = pass_active(P, d, false);
    if (hy && bf) { if (launch_dir(P, d, a, 4, static_cast<cudaStream_t>(s))) return 1; }
    else if (hy) { if (launch_dir(P, d, a, 1, static_cast<cudaStream_t>(s))) return 1; }
    else if (bf) { if (launch_dir(P, d, a + 1, 3, static_cast<cudaStream_t>(s))) return 1; }
  }
  return 0;
}

}  // extern "C"
