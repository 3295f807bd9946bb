// smr.cu -- fine/coarse boundaries of CELL-CENTRED data on a statically refined mesh (SURVEY.md 8f item 2,
// hydro path): restriction into the coarse arrays, ghost exchange across refinement levels, coarse-buffer
// fill + slope-limited prolongation of the fine ghost zones, and flux correction at fine/coarse faces.
//
// Reference: src/mesh/mesh_refinement.cpp:1226-1284 (RestrictCC), src/bvals/buffs_cc.cpp:28-484 (index
// tables), bvals_cc.cpp:42-167,273-402 (PackAndSendCC / RecvAndUnpackCC with coarser and finer neighbours),
// prolongation.cpp:366-547 (FillCoarseInBndryCC, ProlongateCC), mesh/prolongation.hpp:21-60 (ProlongCC),
// flux_correct_cc.cpp:28-290; task order src/hydro/hydro_tasks.cpp:55-71.
//
// Design: the host resolves the neighbour table once into flat descriptor lists (box copies between the
// fine and coarse arrays of two blocks, coarse boxes to fill / prolongate, flux boxes); each step is ONE
// launch over all descriptors of all blocks -- no pack buffers, no per-neighbour kernels.  Where two
// descriptors write the same cell they carry the same bits (a physical cell has one owner; both
// prolongations read the same coarse data), so launch-internal order does not matter; the order BETWEEN the
// steps (exchange, fill, prolongate) is the reference's and is kept by the stream.  Single rank: every
// neighbour must live in this pack (the multi-rank plan is the next step).
#include <cuda_runtime.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/athenak_b200.h"
#include "halo.h"

namespace {

struct Idx {   // RegionIndcs incl. the coarse-array members (mesh.cpp:285-333)
  int ng, nx[3], cnx[3], s[3], e[3], cs[3], ce[3];
  bool act[3];
  int n[3], cn[3];         // array extents, fine and coarse
};

Idx make_idx(const ak_mhd_cfg &c) {
  Idx I;
  I.ng = c.ng;
  const int nx[3] = {c.nx1, c.nx2, c.nx3};
  for (int d = 0; d < 3; ++d) {
    I.nx[d] = nx[d];
    I.act[d] = (d == 0) || nx[d] > 1;
    I.cnx[d] = std::max(1, nx[d] / 2);
    I.s[d] = I.act[d] ? c.ng : 0;  I.e[d] = I.s[d] + nx[d] - 1;
    I.cs[d] = I.act[d] ? c.ng : 0; I.ce[d] = I.cs[d] + I.cnx[d] - 1;
    I.n[d] = I.act[d] ? nx[d] + 2 * c.ng : 1;
    I.cn[d] = I.act[d] ? I.cnx[d] + 2 * c.ng : 1;
  }
  return I;
}

struct Box { int lo[3], hi[3]; };
enum Kind { SAME, COAR, FINE, FLUX, PROL };

// which sub-block index (f1 or f2) governs direction d for offset o (buffs_cc.cpp: x1 uses f1; x2 uses f1 when
// ox1 != 0, else f2; x3 uses f1 when ox1 and ox2 are both non-zero, else f2)
inline int sub_of(int d, const int o[3], int f1, int f2) {
  if (d == 0) return f1;
  if (d == 1) return (o[0] != 0) ? f1 : f2;
  return (o[0] != 0 && o[1] != 0) ? f1 : f2;
}

Box send_box(const Idx &I, const int o[3], int f1, int f2, Kind kind) {   // buffs_cc.cpp:28-152
  Box b;
  const int ng = I.ng, ng1 = ng - 1;
  for (int d = 0; d < 3; ++d) {
    if (kind == FLUX) {
      if (o[d] == 0) { b.lo[d] = I.cs[d]; b.hi[d] = I.ce[d]; }
      else if (o[d] > 0) { b.lo[d] = b.hi[d] = I.ce[d] + 1; }
      else { b.lo[d] = b.hi[d] = I.cs[d]; }
      continue;
    }
    const int s = (kind == COAR) ? I.cs[d] : I.s[d], e = (kind == COAR) ? I.ce[d] : I.e[d];
    b.lo[d] = (o[d] > 0) ? (e - ng1) : s;
    b.hi[d] = (o[d] < 0) ? (s + ng1) : e;
    if (kind == FINE && o[d] == 0 && I.act[d]) {    // internal edges / corners of the face or edge: half + nghost
      if (sub_of(d, o, f1, f2) == 1) b.lo[d] += I.cnx[d] - ng; else b.hi[d] -= I.cnx[d] - ng;
    }
  }
  return b;
}

Box recv_box(const Idx &I, const int o[3], int f1, int f2, Kind kind) {   // buffs_cc.cpp:164-484
  Box b;
  const int ng = I.ng;
  for (int d = 0; d < 3; ++d) {
    if (kind == COAR || kind == PROL) {
      const int cn = (kind == COAR) ? ng : ng / 2;
      if (o[d] == 0) {
        b.lo[d] = I.cs[d]; b.hi[d] = I.ce[d];
        if (I.act[d]) { if (sub_of(d, o, f1, f2) == 0) b.hi[d] += cn; else b.lo[d] -= cn; }
      } else if (o[d] > 0) { b.lo[d] = I.ce[d] + 1; b.hi[d] = I.ce[d] + cn; }
      else { b.lo[d] = I.cs[d] - cn; b.hi[d] = I.cs[d] - 1; }
    } else {
      const int w = (kind == FLUX) ? 1 : ng;
      if (o[d] == 0) {
        b.lo[d] = I.s[d]; b.hi[d] = I.e[d];
        if ((kind == FINE || kind == FLUX) && I.act[d]) { if (sub_of(d, o, f1, f2) == 1) b.lo[d] += I.cnx[d]; else b.hi[d] -= I.cnx[d]; }
      } else if (o[d] > 0) { b.lo[d] = I.e[d] + 1; b.hi[d] = I.e[d] + w; }
      else if (kind == FLUX) { b.lo[d] = b.hi[d] = I.s[d]; }
      else { b.lo[d] = I.s[d] - ng; b.hi[d] = I.s[d] - 1; }
    }
  }
  return b;
}

// slot -> offset and sub-block indices (inverse of nghbr_index.hpp:28-54; faces carry f1,f2, edges f1)
bool slot_decode(int n, int o[3], int &f1, int &f2) {
  for (int iz = -1; iz <= 1; ++iz) for (int iy = -1; iy <= 1; ++iy) for (int ix = -1; ix <= 1; ++ix) {
    if (!ix && !iy && !iz) continue;
    const int nzero = (ix == 0) + (iy == 0) + (iz == 0);
    for (int a = 0; a <= (nzero >= 1 ? 1 : 0); ++a) for (int b = 0; b <= (nzero == 2 ? 1 : 0); ++b)
      if (akb::neighbor_index(ix, iy, iz, a, b) == n) { o[0] = ix; o[1] = iy; o[2] = iz; f1 = a; f2 = b; return true; }
  }
  return false;
}

struct CopyDesc { int src_m, dst_m, src_coarse, dst_coarse; int slo[3], dlo[3], n[3]; };
struct BoxDesc { int m; int lo[3], n[3]; };                                 // coarse-index box of one block
struct FluxDesc { int src_m, dst_m, dir; int slo[3], dlo[3], n[3]; };      // coarse indices on the fine sender, fine indices on the receiver

struct Plan {
  ak_mhd_cfg cfg;
  Idx I;
  std::vector<CopyDesc> copy;
  std::vector<BoxDesc> fill, prol;
  std::vector<FluxDesc> flux;
  CopyDesc *d_copy = nullptr; BoxDesc *d_fill = nullptr, *d_prol = nullptr; FluxDesc *d_flux = nullptr;
  long max_copy = 0, max_fill = 0, max_prol = 0, max_flux = 0;
  bool uploaded = false;
  std::string err;
};

thread_local std::string g_smr_err;

struct Dims { int n1, n2, n3, c1, c2, c3; long cc, ccc; int s[3], cs[3]; int multi_d, three_d; long f[3]; };

Dims dims_of(const Idx &I) {
  Dims D;
  D.n1 = I.n[0]; D.n2 = I.n[1]; D.n3 = I.n[2]; D.c1 = I.cn[0]; D.c2 = I.cn[1]; D.c3 = I.cn[2];
  D.cc = (long)D.n1 * D.n2 * D.n3; D.ccc = (long)D.c1 * D.c2 * D.c3;
  for (int d = 0; d < 3; ++d) { D.s[d] = I.s[d]; D.cs[d] = I.cs[d]; }
  D.multi_d = I.act[1]; D.three_d = I.act[2];
  D.f[0] = (long)D.n3 * D.n2 * (D.n1 + 1); D.f[1] = (long)D.n3 * (D.n2 + 1) * D.n1; D.f[2] = (long)(D.n3 + 1) * D.n2 * D.n1;
  return D;
}

// ---------------------------------------------------------------------------------------------------
// kernels

// mesh_refinement.cpp:1226-1284: thread per (block, variable, coarse active cell)
__global__ void __launch_bounds__(256)
smr_restrict_kernel(Dims D, int nmb, int nvar, int cnx1, int cnx2, int cnx3, const double *__restrict__ u, double *__restrict__ cu) {
  const long total = (long)nmb * nvar * cnx3 * cnx2 * cnx1;
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int i = D.cs[0] + (int)(t % cnx1); t /= cnx1;
  const int j = D.cs[1] + (int)(t % cnx2); t /= cnx2;
  const int k = D.cs[2] + (int)(t % cnx3); t /= cnx3;      // t = m*nvar + v
  const int fi = 2 * i - D.cs[0], fj = 2 * j - D.cs[1], fk = 2 * k - D.cs[2];
  const double *a = u + t * D.cc;
  auto A = [&](int kk, int jj, int ii) { return a[((long)kk * D.n2 + jj) * D.n1 + ii]; };
  double r;
  if (!D.multi_d) r = 0.5 * (A(fk, fj, fi) + A(fk, fj, fi + 1));
  else if (!D.three_d) r = 0.25 * (A(fk, fj, fi) + A(fk, fj, fi + 1) + A(fk, fj + 1, fi) + A(fk, fj + 1, fi + 1));
  else r = 0.125 * (A(fk, fj, fi) + A(fk, fj, fi + 1) + A(fk, fj + 1, fi) + A(fk, fj + 1, fi + 1)
                  + A(fk + 1, fj, fi) + A(fk + 1, fj, fi + 1) + A(fk + 1, fj + 1, fi) + A(fk + 1, fj + 1, fi + 1));
  cu[t * D.ccc + ((long)k * D.c2 + j) * D.c1 + i] = r;
}

// bvals_cc.cpp:42-167 + 273-402 without the buffer: blockIdx.x = descriptor, blockIdx.y x threads over (variable, box)
__global__ void __launch_bounds__(256)
smr_copy_kernel(Dims D, int nvar, const CopyDesc *__restrict__ descs, double *__restrict__ u, double *__restrict__ cu) {
  const CopyDesc d = descs[blockIdx.x];
  const long cnt = (long)d.n[0] * d.n[1] * d.n[2] * nvar;
  for (long t = (long)blockIdx.y * blockDim.x + threadIdx.x; t < cnt; t += (long)gridDim.y * blockDim.x) {
    long r = t;
    const int i = (int)(r % d.n[0]); r /= d.n[0];
    const int j = (int)(r % d.n[1]); r /= d.n[1];
    const int k = (int)(r % d.n[2]);
    const int v = (int)(r / d.n[2]);
    const double *src = d.src_coarse ? cu + ((long)d.src_m * nvar + v) * D.ccc + ((long)(d.slo[2] + k) * D.c2 + d.slo[1] + j) * D.c1 + d.slo[0] + i
                                     : u + ((long)d.src_m * nvar + v) * D.cc + ((long)(d.slo[2] + k) * D.n2 + d.slo[1] + j) * D.n1 + d.slo[0] + i;
    double *dst = d.dst_coarse ? cu + ((long)d.dst_m * nvar + v) * D.ccc + ((long)(d.dlo[2] + k) * D.c2 + d.dlo[1] + j) * D.c1 + d.dlo[0] + i
                               : u + ((long)d.dst_m * nvar + v) * D.cc + ((long)(d.dlo[2] + k) * D.n2 + d.dlo[1] + j) * D.n1 + d.dlo[0] + i;
    *dst = *src;
  }
}

// prolongation.cpp:366-464 (FillCoarseInBndryCC): restrict same-level ghost data into the coarse array
__global__ void __launch_bounds__(256)
smr_fill_kernel(Dims D, int nvar, const BoxDesc *__restrict__ descs, const double *__restrict__ u, double *__restrict__ cu) {
  const BoxDesc d = descs[blockIdx.x];
  const long cnt = (long)d.n[0] * d.n[1] * d.n[2] * nvar;
  for (long t = (long)blockIdx.y * blockDim.x + threadIdx.x; t < cnt; t += (long)gridDim.y * blockDim.x) {
    long r = t;
    const int i = d.lo[0] + (int)(r % d.n[0]); r /= d.n[0];
    const int j = d.lo[1] + (int)(r % d.n[1]); r /= d.n[1];
    const int k = d.lo[2] + (int)(r % d.n[2]);
    const int v = (int)(r / d.n[2]);
    const int fi = (i - D.cs[0]) * 2 + D.s[0], fj = (j - D.cs[1]) * 2 + D.s[1], fk = (k - D.cs[2]) * 2 + D.s[2];
    const double *a = u + ((long)d.m * nvar + v) * D.cc;
    auto A = [&](int kk, int jj, int ii) { return a[((long)kk * D.n2 + jj) * D.n1 + ii]; };
    double q;
    if (!D.three_d) q = 0.25 * (A(fk, fj, fi) + A(fk, fj, fi + 1) + A(fk, fj + 1, fi) + A(fk, fj + 1, fi + 1));
    else q = 0.125 * (A(fk, fj, fi) + A(fk, fj, fi + 1) + A(fk, fj + 1, fi) + A(fk, fj + 1, fi + 1)
                    + A(fk + 1, fj, fi) + A(fk + 1, fj, fi + 1) + A(fk + 1, fj + 1, fi) + A(fk + 1, fj + 1, fi + 1));
    cu[((long)d.m * nvar + v) * D.ccc + ((long)k * D.c2 + j) * D.c1 + i] = q;
  }
}

__device__ __forceinline__ double sgn(double x) { return (x < 0.0) ? -1.0 : 1.0; }      // athena.hpp:52
__device__ __forceinline__ double minmod_slope(double dl, double dr) {                  // prolongation.hpp:27-29
  const double al = fabs(dl), ar = fabs(dr);
  return 0.125 * (sgn(dl) + sgn(dr)) * ((al < ar) ? al : ar);
}

// prolongation.cpp:471-547 + mesh/prolongation.hpp:21-60: thread per coarse cell of a prolongation box
__global__ void __launch_bounds__(256)
smr_prolong_kernel(Dims D, int nvar, const BoxDesc *__restrict__ descs, double *__restrict__ u, const double *__restrict__ cu) {
  const BoxDesc d = descs[blockIdx.x];
  const long cnt = (long)d.n[0] * d.n[1] * d.n[2] * nvar;
  for (long t = (long)blockIdx.y * blockDim.x + threadIdx.x; t < cnt; t += (long)gridDim.y * blockDim.x) {
    long r = t;
    const int i = d.lo[0] + (int)(r % d.n[0]); r /= d.n[0];
    const int j = d.lo[1] + (int)(r % d.n[1]); r /= d.n[1];
    const int k = d.lo[2] + (int)(r % d.n[2]);
    const int v = (int)(r / d.n[2]);
    const int fi = (i - D.cs[0]) * 2 + D.s[0], fj = (j - D.cs[1]) * 2 + D.s[1], fk = (k - D.cs[2]) * 2 + D.s[2];
    const double *c = cu + ((long)d.m * nvar + v) * D.ccc + ((long)k * D.c2 + j) * D.c1 + i;
    const long sj = D.c1, sk = (long)D.c1 * D.c2;
    const double q = c[0];
    const double dv1 = minmod_slope(q - c[-1], c[1] - q);
    const double dv2 = D.multi_d ? minmod_slope(q - c[-sj], c[sj] - q) : 0.0;
    const double dv3 = D.three_d ? minmod_slope(q - c[-sk], c[sk] - q) : 0.0;
    double *a = u + ((long)d.m * nvar + v) * D.cc + ((long)fk * D.n2 + fj) * D.n1 + fi;
    const long fsj = D.n1, fsk = (long)D.n1 * D.n2;
    a[0] = q - dv1 - dv2 - dv3;
    a[1] = q + dv1 - dv2 - dv3;
    if (D.multi_d) {
      a[fsj] = q - dv1 + dv2 - dv3;
      a[fsj + 1] = q + dv1 + dv2 - dv3;
    }
    if (D.three_d) {
      a[fsk] = q - dv1 - dv2 + dv3;
      a[fsk + 1] = q + dv1 - dv2 + dv3;
      a[fsk + fsj] = q - dv1 + dv2 + dv3;
      a[fsk + fsj + 1] = q + dv1 + dv2 + dv3;
    }
  }
}

// flux_correct_cc.cpp:28-290 without the buffer: the fine block's boundary-face fluxes, averaged over the
// 1 / 2 / 4 fine faces of a coarse face, replace the coarse neighbour's flux there.
__global__ void __launch_bounds__(128)
smr_flux_kernel(Dims D, int nvar, const FluxDesc *__restrict__ descs, double *__restrict__ flx1, double *__restrict__ flx2,
                double *__restrict__ flx3) {
  const FluxDesc d = descs[blockIdx.x];
  const long cnt = (long)d.n[0] * d.n[1] * d.n[2] * nvar;
  double *flx = d.dir == 0 ? flx1 : d.dir == 1 ? flx2 : flx3;
  const long e1 = D.n1 + (d.dir == 0), e2 = D.n2 + (d.dir == 1);      // face-array extents along x1, x2
  const long fsz = D.f[d.dir];
  for (long t = (long)blockIdx.y * blockDim.x + threadIdx.x; t < cnt; t += (long)gridDim.y * blockDim.x) {
    long r = t;
    const int a = (int)(r % d.n[0]); r /= d.n[0];
    const int b = (int)(r % d.n[1]); r /= d.n[1];
    const int c = (int)(r % d.n[2]);
    const int v = (int)(r / d.n[2]);
    const int i = d.slo[0] + a, j = d.slo[1] + b, k = d.slo[2] + c;          // coarse indices on the sender
    const int fi = 2 * i - D.cs[0], fj = 2 * j - D.cs[1], fk = 2 * k - D.cs[2];
    const double *s = flx + ((long)d.src_m * nvar + v) * fsz;
    auto F = [&](int kk, int jj, int ii) { return s[((long)kk * e2 + jj) * e1 + ii]; };
    double q;
    if (d.dir == 0) {
      if (!D.multi_d) q = F(0, 0, fi);
      else if (!D.three_d) q = 0.5 * (F(0, fj, fi) + F(0, fj + 1, fi));
      else q = 0.25 * (F(fk, fj, fi) + F(fk, fj + 1, fi) + F(fk + 1, fj, fi) + F(fk + 1, fj + 1, fi));
    } else if (d.dir == 1) {
      if (!D.three_d) q = 0.5 * (F(0, fj, fi) + F(0, fj, fi + 1));
      else q = 0.25 * (F(fk, fj, fi) + F(fk, fj, fi + 1) + F(fk + 1, fj, fi) + F(fk + 1, fj, fi + 1));
    } else {
      q = 0.25 * (F(fk, fj, fi) + F(fk, fj, fi + 1) + F(fk, fj + 1, fi) + F(fk, fj + 1, fi + 1));
    }
    flx[((long)d.dst_m * nvar + v) * fsz + ((long)(d.dlo[2] + c) * e2 + d.dlo[1] + b) * e1 + d.dlo[0] + a] = q;
  }
}

template <typename T>
cudaError_t upload(const std::vector<T> &h, T **d) {
  if (h.empty()) return cudaSuccess;
  cudaError_t e = cudaMalloc(d, sizeof(T) * h.size());
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice);
}

int smr_fail(Plan *P, const std::string &m) { if (P) P->err = m; g_smr_err = m; return 1; }

int ensure_uploaded(Plan *P) {
  if (P->uploaded) return 0;
  cudaError_t e;
  if ((e = upload(P->copy, &P->d_copy)) != cudaSuccess || (e = upload(P->fill, &P->d_fill)) != cudaSuccess ||
      (e = upload(P->prol, &P->d_prol)) != cudaSuccess || (e = upload(P->flux, &P->d_flux)) != cudaSuccess)
    return smr_fail(P, std::string("ak_smr: ") + cudaGetErrorString(e));
  P->uploaded = true;
  return 0;
}

inline unsigned gx(long maxcount, int bs) { return (unsigned)std::max(1L, std::min((maxcount + bs - 1) / bs, 64L)); }

}  // namespace

extern "C" {

const char *ak_smr_last_error(void *plan) { return plan ? static_cast<Plan *>(plan)->err.c_str() : g_smr_err.c_str(); }

int ak_smr_create(const ak_mhd_cfg *cfg, const int *nghbr, const int *mb_lev, int my_rank, int gids_start, void **plan) {
  if (!cfg || !nghbr || !mb_lev || !plan) return smr_fail(nullptr, "ak_smr_create: null argument");
  *plan = nullptr;
  if (cfg->ng % 2) return smr_fail(nullptr, "ak_smr_create: nghost must be even with SMR/AMR (mesh.cpp:234-239)");
  const int nxs[3] = {cfg->nx1, cfg->nx2, cfg->nx3};
  for (int d = 0; d < 3; ++d)
    if ((d == 0 || nxs[d] > 1) && (nxs[d] % 2 || nxs[d] / 2 < cfg->ng))
      return smr_fail(nullptr, "ak_smr_create: MeshBlock sizes must be even and at least 2*nghost in every active direction");
  Plan *P = new Plan();
  P->cfg = *cfg;
  P->I = make_idx(*cfg);
  const Idx &I = P->I;
  const int nmb = cfg->nmb;
  auto cnt = [](const int n[3]) { return (long)n[0] * n[1] * n[2]; };
  for (int m = 0; m < nmb; ++m)
    for (int n = 0; n < 56; ++n) {
      const int *e = nghbr + ((size_t)m * 56 + n) * 4;
      if (e[0] < 0) continue;
      if (e[2] != my_rank || e[0] < gids_start || e[0] >= gids_start + nmb) {
        delete P;
        return smr_fail(nullptr, "ak_smr_create: neighbour on another rank (the multi-level plan is single-rank for now)");
      }
      int o[3], f1, f2;
      slot_decode(n, o, f1, f2);
      const int dm = e[0] - gids_start, dn = e[3];
      const int mylev = mb_lev[m], nlev = e[1];
      // ---- exchange: my send box -> the receive box of slot dn on block dm (bvals_cc.cpp:76-140, 322-384)
      {
        const Kind sk = (nlev < mylev) ? COAR : (nlev == mylev) ? SAME : FINE;
        const Box sb = send_box(I, o, f1, f2, sk);
        int ro[3], rf1, rf2;
        slot_decode(dn, ro, rf1, rf2);
        // the receiver sees ME at level mylev from level nlev
        const Kind rk = (mylev < nlev) ? COAR : (mylev == nlev) ? SAME : FINE;
        const Box rb = recv_box(I, ro, rf1, rf2, rk);
        CopyDesc d;
        d.src_m = m; d.dst_m = dm; d.src_coarse = (sk == COAR); d.dst_coarse = (rk == COAR);
        for (int q = 0; q < 3; ++q) {
          d.slo[q] = sb.lo[q]; d.dlo[q] = rb.lo[q]; d.n[q] = sb.hi[q] - sb.lo[q] + 1;
          if (d.n[q] != rb.hi[q] - rb.lo[q] + 1) { delete P; return smr_fail(nullptr, "ak_smr_create: send/receive boxes differ in size"); }
        }
        P->copy.push_back(d);
        P->max_copy = std::max(P->max_copy, cnt(d.n));
      }
      // ---- FillCoarseInBndryCC: same-level neighbours, multi-D only (prolongation.cpp:396-414)
      if (nlev == mylev && I.act[1]) {
        const Box rb = recv_box(I, o, 0, 0, SAME);
        BoxDesc d; d.m = m;
        for (int q = 0; q < 3; ++q) { d.lo[q] = (rb.lo[q] + I.cs[q]) / 2; d.n[q] = (rb.hi[q] + I.cs[q]) / 2 - d.lo[q] + 1; }
        P->fill.push_back(d);
        P->max_fill = std::max(P->max_fill, cnt(d.n));
      }
      // ---- ProlongateCC: coarser neighbours (prolongation.cpp:498-507)
      if (nlev < mylev) {
        const Box pb = recv_box(I, o, f1, f2, PROL);
        BoxDesc d; d.m = m;
        for (int q = 0; q < 3; ++q) { d.lo[q] = pb.lo[q]; d.n[q] = pb.hi[q] - pb.lo[q] + 1; }
        P->prol.push_back(d);
        P->max_prol = std::max(P->max_prol, cnt(d.n));
      }
      // ---- flux correction: faces toward a coarser neighbour (flux_correct_cc.cpp:76-78, 246-248)
      if (nlev < mylev && ((n < 16) || (n >= 24 && n < 32))) {
        const Box sb = send_box(I, o, f1, f2, FLUX);
        int ro[3], rf1, rf2;
        slot_decode(dn, ro, rf1, rf2);
        const Box rb = recv_box(I, ro, rf1, rf2, FLUX);
        FluxDesc d;
        d.src_m = m; d.dst_m = dm; d.dir = (n < 8) ? 0 : (n < 16) ? 1 : 2;
        for (int q = 0; q < 3; ++q) {
          d.slo[q] = sb.lo[q]; d.dlo[q] = rb.lo[q]; d.n[q] = sb.hi[q] - sb.lo[q] + 1;
          if (d.n[q] != rb.hi[q] - rb.lo[q] + 1) { delete P; return smr_fail(nullptr, "ak_smr_create: flux send/receive boxes differ in size"); }
        }
        P->flux.push_back(d);
        P->max_flux = std::max(P->max_flux, cnt(d.n));
      }
    }
  *plan = P;
  return 0;
}

int ak_smr_destroy(void *plan) {
  Plan *P = static_cast<Plan *>(plan);
  if (!P) return 0;
  cudaFree(P->d_copy); cudaFree(P->d_fill); cudaFree(P->d_prol); cudaFree(P->d_flux);
  delete P;
  return 0;
}

long ak_smr_coarse_cells(void *plan) {
  Plan *P = static_cast<Plan *>(plan);
  return P ? (long)P->I.cn[0] * P->I.cn[1] * P->I.cn[2] : 0;
}

/* counts[4] = exchange, fill, prolongation and flux descriptors (host-side plan, no device needed) */
int ak_smr_counts(void *plan, long *counts) {
  Plan *P = static_cast<Plan *>(plan);
  if (!P || !counts) return 1;
  counts[0] = (long)P->copy.size(); counts[1] = (long)P->fill.size(); counts[2] = (long)P->prol.size(); counts[3] = (long)P->flux.size();
  return 0;
}

#define SMR_PROLOGUE(name)                                               \
  Plan *P = static_cast<Plan *>(plan);                                   \
  if (!P) return smr_fail(nullptr, name ": null plan");                  \
  if (ensure_uploaded(P)) return 1;                                      \
  cudaStream_t st = static_cast<cudaStream_t>(s);                        \
  const Dims D = dims_of(P->I);

#define SMR_EPILOGUE(name)                                               \
  { cudaError_t e = cudaGetLastError();                                  \
    if (e != cudaSuccess) return smr_fail(P, std::string(name ": ") + cudaGetErrorString(e)); }  \
  return 0;

int ak_smr_restrict_cc(void *plan, const double *u, double *cu, int nvar, ak_stream_t s) {
  SMR_PROLOGUE("ak_smr_restrict_cc")
  const Idx &I = P->I;
  const long n = (long)P->cfg.nmb * nvar * I.cnx[0] * (I.act[1] ? I.cnx[1] : 1) * (I.act[2] ? I.cnx[2] : 1);
  smr_restrict_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(D, P->cfg.nmb, nvar, I.cnx[0], I.act[1] ? I.cnx[1] : 1, I.act[2] ? I.cnx[2] : 1, u, cu);
  SMR_EPILOGUE("ak_smr_restrict_cc")
}

int ak_smr_exchange_cc(void *plan, double *u, double *cu, int nvar, ak_stream_t s) {
  SMR_PROLOGUE("ak_smr_exchange_cc")
  if (!P->copy.empty())
    smr_copy_kernel<<<dim3((unsigned)P->copy.size(), gx(P->max_copy * nvar, 256)), 256, 0, st>>>(D, nvar, P->d_copy, u, cu);
  SMR_EPILOGUE("ak_smr_exchange_cc")
}

int ak_smr_prolongate_cc(void *plan, double *u, double *cu, int nvar, ak_stream_t s) {
  return ak_smr_prolongate_cc_steps(plan, u, cu, nvar, 3, s);
}

int ak_smr_prolongate_cc_steps(void *plan, double *u, double *cu, int nvar, int steps, ak_stream_t s) {
  SMR_PROLOGUE("ak_smr_prolongate_cc")
  if ((steps & 1) && !P->fill.empty())
    smr_fill_kernel<<<dim3((unsigned)P->fill.size(), gx(P->max_fill * nvar, 256)), 256, 0, st>>>(D, nvar, P->d_fill, u, cu);
  if ((steps & 2) && !P->prol.empty())
    smr_prolong_kernel<<<dim3((unsigned)P->prol.size(), gx(P->max_prol * nvar, 256)), 256, 0, st>>>(D, nvar, P->d_prol, u, cu);
  SMR_EPILOGUE("ak_smr_prolongate_cc")
}

int ak_smr_flux_correct_cc(void *plan, double *flx1, double *flx2, double *flx3, int nvar, ak_stream_t s) {
  SMR_PROLOGUE("ak_smr_flux_correct_cc")
  if (!P->flux.empty())
    smr_flux_kernel<<<dim3((unsigned)P->flux.size(), gx(P->max_flux * nvar, 128)), 128, 0, st>>>(D, nvar, P->d_flux, flx1, flx2, flx3);
  SMR_EPILOGUE("ak_smr_flux_correct_cc")
}

}  // extern "C"
