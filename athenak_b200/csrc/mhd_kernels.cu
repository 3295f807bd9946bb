// mhd_kernels.cu -- hand-written sm_100a kernels for the per-MeshBlock ideal-MHD stage update
// (reference path: src/reconstruct -> src/mhd/rsolvers -> mhd_corner_e -> mhd_update/mhd_ct ->
// src/eos ConsToPrim).  No tensor cores: the path has no contraction.
// FP64 SIMT, --fmad=false, bit-faithful to the reference's expression order.
//
// Kernel inventory (algorithmic traffic in FP64 words per active cell per stage):
//   flux_x1        warp = 32 consecutive cells of one i-pencil; 5-point stencils come from L1,
//                  the left face state from lane-1 by warp shuffle; recon + Riemann in registers.
//   flux_march<D>  x2/x3: one thread per (transverse, i) column marching along the sweep
//                  direction, carrying the left face state in registers (coalesced along i).
//   corner_emf     GS07 upwinded corner EMFs, cell-centred E recomputed on the fly.
//   update_ct      RK update of u0 and CT update of b0 in one pass (+ CopyCons at stage 1).
//   cons2prim      c2p over a cell box, floors, floor counters; optional fused new-dt reduce.
#include "mhd_kernels.h"
#include <cstdlib>
#include "flux_kernels.cuh"
#include "reduce.cuh"
#include "c2p_device.cuh"

#ifndef AKB_UPD_FORM
#define AKB_UPD_FORM 0        // conserved update: 0 = variable by variable (guard check per variable), 1 = all variables, one check per cell
#endif
#ifndef AKB_UPD_SCALED
#define AKB_UPD_SCALED 1      // 1: the update's quotients through div_by_x (exact for every numerator, practically no repeat): 0.92 -> 0.84 ms
#endif
#if AKB_UPD_SCALED
#define UPD_DIV div_by_x
#else
#define UPD_DIV div_by
#endif

namespace akb {

// ---------------------------------------------------------------------------------------
// GS07 corner EMFs (mhd_corner_e.cpp:309-318,338-414), one thread per corner
// k in [ks,ke+1], j in [js,je+1], i in [is,ie+1].
// Thread -> (kk, jj, ii) of the corner / face index space [0, nk) x [0, nj) x [0, ni) of block m.  FLAT = false: 32 x 4 x 2 tiles
// (neighbours in j and k mostly come from L1).  FLAT = true: the (jj, ii) plane flattened, 128 consecutive positions x 2
// k-layers per CTA -- for extents like 64 + 1 the tiles would leave a third of the warps with one active lane.
template <bool FLAT>
__device__ __forceinline__ bool stencil_index(int ni, int nj, int nk, int &ii, int &jj, int &kk, int &m) {
  const int nkb = (nk + TILE_Z - 1) / TILE_Z;
  if (FLAT) {
    const int p = blockIdx.x * 128 + (threadIdx.x & 127);
    jj = p / ni; ii = p - jj * ni;
    kk = (blockIdx.y % nkb) * TILE_Z + (threadIdx.x >> 7);
    m = blockIdx.y / nkb;
  } else {
    ii = blockIdx.x * TILE_X + (threadIdx.x & 31); jj = blockIdx.y * TILE_Y + ((threadIdx.x >> 5) & 3);
    kk = (blockIdx.z % nkb) * TILE_Z + (threadIdx.x >> 7);
    m = blockIdx.z / nkb;
  }
  return ii < ni && jj < nj && kk < nk;
}
// tile efficiency below 90 %: flatten.  257^3 positions (the headline block's corners / faces) fill 87.8 % of the 3-D tiles, and every ninth
// tile along x1 holds one position per warp: flattened, corner EMF 0.563 -> 0.520 ms and CT 0.418 -> 0.399 ms (profiles/r02_s44.sh).
// AKB_FLAT_THRESHOLD overrides the limit (measurement only).
static inline bool stencil_flat(int ni, int nj, int nk) {
  static const double thr = [] { const char *e = getenv("AKB_FLAT_THRESHOLD"); return e ? atof(e) : 0.90; }();
  const double used = (double)ni * nj * nk;
  const double launched = (double)((ni + TILE_X - 1) / TILE_X * TILE_X) * ((nj + TILE_Y - 1) / TILE_Y * TILE_Y) * ((nk + TILE_Z - 1) / TILE_Z * TILE_Z);
  return used < thr * launched;
}
static inline dim3 stencil_grid(bool flat, int ni, int nj, int nk, int nmb) {
  const int nkb = (nk + TILE_Z - 1) / TILE_Z;
  return flat ? dim3((ni * nj + 127) / 128, nkb * nmb, 1) : dim3((ni + TILE_X - 1) / TILE_X, (nj + TILE_Y - 1) / TILE_Y, nkb * nmb);
}

template <bool FLAT>
__global__ void __launch_bounds__(256, AKB_CORNER_MINB)
corner_emf_kernel(Geom g, const double *__restrict__ w0, const double *__restrict__ bcc,
                  const double *__restrict__ flx1, const double *__restrict__ flx2,
                  const double *__restrict__ flx3, const double *__restrict__ e3x1,
                  const double *__restrict__ e2x1, const double *__restrict__ e1x2,
                  const double *__restrict__ e3x2, const double *__restrict__ e2x3,
                  const double *__restrict__ e1x3, double *__restrict__ e1, double *__restrict__ e2,
                  double *__restrict__ e3) {
  // 3-D thread tile (32 x 4 x 2): the j-1 / k-1 neighbours of most threads are loaded by the
  // same block, so they come from L1
  int ii, jj, kk, m;
  if (!stencil_index<FLAT>(g.nx1 + 1, g.nx2 + 1, g.nx3 + 1, ii, jj, kk, m)) return;
  const int i = g.is + ii, j = g.js + jj, k = g.ks + kk;
  const long cc = g.cc, n1 = g.n1, n12 = (long)g.n1 * g.n2;
  const double *w = w0 + (long)m * g.nvar * cc, *b = bcc + (long)m * 3 * cc;
  auto C = [&](int kk, int jj, int ii) -> long { return ((long)kk * g.n2 + jj) * n1 + ii; };
  // cell-centred E = -(v x B)  (mhd_corner_e.cpp:310-317)
  auto e1cc = [&](long o) { return __ldg(w + IVZ * cc + o) * __ldg(b + 1 * cc + o) - __ldg(w + IVY * cc + o) * __ldg(b + 2 * cc + o); };
  auto e2cc = [&](long o) { return __ldg(w + IVX * cc + o) * __ldg(b + 2 * cc + o) - __ldg(w + IVZ * cc + o) * __ldg(b + 0 * cc + o); };
  auto e3cc = [&](long o) { return __ldg(w + IVY * cc + o) * __ldg(b + 0 * cc + o) - __ldg(w + IVX * cc + o) * __ldg(b + 1 * cc + o); };
  // mass fluxes (IDN component) on the three face families
  const double *f1 = flx1 + (long)m * g.nvar * g.f1;
  const double *f2 = flx2 + (long)m * g.nvar * g.f2;
  const double *f3 = flx3 + (long)m * g.nvar * g.f3;
  auto F1 = [&](int kk, int jj, int ii) { return __ldg(f1 + ((long)kk * g.n2 + jj) * (n1 + 1) + ii); };
  auto F2 = [&](int kk, int jj, int ii) { return __ldg(f2 + ((long)kk * (g.n2 + 1) + jj) * n1 + ii); };
  auto F3 = [&](int kk, int jj, int ii) { return __ldg(f3 + ((long)kk * g.n2 + jj) * n1 + ii); };
  const long mo = (long)m * cc;
  const long o = C(k, j, i);
  // Branch-free form of mhd_corner_e.cpp:342-413: all four cell-centred values of a component
  // are evaluated (loads are independent, shared between components by CSE), both candidate
  // differences are formed and the mass-flux sign selects one.  Same values, same sum order.
  {
    const double x3_jm = __ldg(e1x3 + mo + o - n1), x3_j = __ldg(e1x3 + mo + o);
    const double x2_km = __ldg(e1x2 + mo + o - n12), x2_k = __ldg(e1x2 + mo + o);
    const double c_mm = e1cc(o - n12 - n1), c_m0 = e1cc(o - n12), c_0m = e1cc(o - n1), c_00 = e1cc(o);
    const double l3 = (F2(k - 1, j, i) >= 0.0) ? (x3_jm - c_mm) : (x3_j - c_m0);
    const double r3 = (F2(k, j, i) >= 0.0) ? (x3_jm - c_0m) : (x3_j - c_00);
    const double l2 = (F3(k, j - 1, i) >= 0.0) ? (x2_km - c_mm) : (x2_k - c_0m);
    const double r2 = (F3(k, j, i) >= 0.0) ? (x2_km - c_m0) : (x2_k - c_00);
    e1[(long)m * g.e1 + ((long)k * (g.n2 + 1) + j) * n1 + i] = 0.25 * (l3 + r3 + l2 + r2 + x2_km + x2_k + x3_jm + x3_j);
  }
  {
    const double x3_im = __ldg(e2x3 + mo + o - 1), x3_i = __ldg(e2x3 + mo + o);
    const double x1_km = __ldg(e2x1 + mo + o - n12), x1_k = __ldg(e2x1 + mo + o);
    const double c_mm = e2cc(o - n12 - 1), c_m0 = e2cc(o - n12), c_0m = e2cc(o - 1), c_00 = e2cc(o);
    const double l3 = (F1(k - 1, j, i) >= 0.0) ? (x3_im - c_mm) : (x3_i - c_m0);
    const double r3 = (F1(k, j, i) >= 0.0) ? (x3_im - c_0m) : (x3_i - c_00);
    const double l1 = (F3(k, j, i - 1) >= 0.0) ? (x1_km - c_mm) : (x1_k - c_0m);
    const double r1 = (F3(k, j, i) >= 0.0) ? (x1_km - c_m0) : (x1_k - c_00);
    e2[(long)m * g.e2 + ((long)k * g.n2 + j) * (n1 + 1) + i] = 0.25 * (l3 + r3 + l1 + r1 + x3_im + x3_i + x1_km + x1_k);
  }
  {
    const double x2_im = __ldg(e3x2 + mo + o - 1), x2_i = __ldg(e3x2 + mo + o);
    const double x1_jm = __ldg(e3x1 + mo + o - n1), x1_j = __ldg(e3x1 + mo + o);
    const double c_mm = e3cc(o - n1 - 1), c_m0 = e3cc(o - n1), c_0m = e3cc(o - 1), c_00 = e3cc(o);
    const double l2 = (F1(k, j - 1, i) >= 0.0) ? (x2_im - c_mm) : (x2_i - c_m0);
    const double r2 = (F1(k, j, i) >= 0.0) ? (x2_im - c_0m) : (x2_i - c_00);
    const double l1 = (F2(k, j, i - 1) >= 0.0) ? (x1_jm - c_mm) : (x1_j - c_0m);
    const double r1 = (F2(k, j, i) >= 0.0) ? (x1_jm - c_m0) : (x1_j - c_00);
    e3[(long)m * g.e3 + ((long)k * (g.n2 + 1) + j) * (n1 + 1) + i] = 0.25 * (l1 + r1 + l2 + r2 + x2_im + x2_i + x1_jm + x1_j);
  }
}

// 1-D: the face-centred EMFs are the edge EMFs (mhd_corner_e.cpp:41-54).  One thread per x1-face.
__global__ void __launch_bounds__(128)
corner_emf_1d_kernel(Geom g, const double *__restrict__ e3x1, const double *__restrict__ e2x1,
                     double *__restrict__ e2, double *__restrict__ e3) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ni = g.nx1 + 1;
  if (t >= (long)g.nmb * ni) return;
  const int m = (int)(t / ni), i = g.is + (int)(t % ni);
  const double v2 = __ldg(e2x1 + (long)m * g.cc + i), v3 = __ldg(e3x1 + (long)m * g.cc + i);
  double *E2 = e2 + (long)m * g.e2, *E3 = e3 + (long)m * g.e3;   // x2e (2,1,n1+1), x3e (1,2,n1+1)
  E2[i] = v2;
  E2[(long)(g.n1 + 1) + i] = v2;          // k = ke+1
  E3[i] = v3;
  E3[(long)(g.n1 + 1) + i] = v3;          // j = je+1
}

// 2-D: E1, E2 copied from the face EMFs, E3 by the GS07 upwinding (mhd_corner_e.cpp:59-201,
// Newtonian branch).  One thread per corner j in [js,je+1], i in [is,ie+1]; branch-free form of
// the selection as in corner_emf_kernel.
__global__ void __launch_bounds__(128)
corner_emf_2d_kernel(Geom g, const double *__restrict__ w0, const double *__restrict__ bcc,
                     const double *__restrict__ flx1, const double *__restrict__ flx2,
                     const double *__restrict__ e3x1, const double *__restrict__ e2x1,
                     const double *__restrict__ e1x2, const double *__restrict__ e3x2,
                     double *__restrict__ e1, double *__restrict__ e2, double *__restrict__ e3) {
  const int ii = blockIdx.x * 32 + threadIdx.x, jj = blockIdx.y * 4 + threadIdx.y;
  const int m = blockIdx.z;
  if (ii > g.nx1 || jj > g.nx2) return;
  const int i = g.is + ii, j = g.js + jj;
  const long cc = g.cc, n1 = g.n1;
  const long o = (long)j * n1 + i, mo = (long)m * cc;
  const double *w = w0 + (long)m * g.nvar * cc, *b = bcc + (long)m * 3 * cc;
  auto e3cc = [&](long q) { return __ldg(w + IVY * cc + q) * __ldg(b + 0 * cc + q) - __ldg(w + IVX * cc + q) * __ldg(b + 1 * cc + q); };
  const double *f1 = flx1 + (long)m * g.nvar * g.f1, *f2 = flx2 + (long)m * g.nvar * g.f2;
  auto F1 = [&](int jq, int iq) { return __ldg(f1 + (long)jq * (n1 + 1) + iq); };
  auto F2 = [&](int jq, int iq) { return __ldg(f2 + (long)jq * n1 + iq); };
  const double v2 = __ldg(e2x1 + mo + o), v1 = __ldg(e1x2 + mo + o);
  double *E1 = e1 + (long)m * g.e1, *E2 = e2 + (long)m * g.e2, *E3 = e3 + (long)m * g.e3;
  // x1e (2, n2+1, n1); x2e (2, n2, n1+1); x3e (1, n2+1, n1+1)
  E2[(long)j * (n1 + 1) + i] = v2;
  E2[(long)g.n2 * (n1 + 1) + (long)j * (n1 + 1) + i] = v2;
  E1[(long)j * n1 + i] = v1;
  E1[(long)(g.n2 + 1) * n1 + (long)j * n1 + i] = v1;
  const double x2_im = __ldg(e3x2 + mo + o - 1), x2_i = __ldg(e3x2 + mo + o);
  const double x1_jm = __ldg(e3x1 + mo + o - n1), x1_j = __ldg(e3x1 + mo + o);
  const double c_mm = e3cc(o - n1 - 1), c_m0 = e3cc(o - n1), c_0m = e3cc(o - 1), c_00 = e3cc(o);
  const double l2 = (F1(j - 1, i) >= 0.0) ? (x2_im - c_mm) : (x2_i - c_m0);
  const double r2 = (F1(j, i) >= 0.0) ? (x2_im - c_0m) : (x2_i - c_00);
  const double l1 = (F2(j, i - 1) >= 0.0) ? (x1_jm - c_mm) : (x1_j - c_0m);
  const double r1 = (F2(j, i) >= 0.0) ? (x1_jm - c_m0) : (x1_j - c_00);
  E3[(long)j * (n1 + 1) + i] = 0.25 * (l1 + r1 + l2 + r2 + x2_im + x2_i + x1_jm + x1_j);
}

// ---------------------------------------------------------------------------------------
// RK update (mhd_update.cpp:51-82) and CT (mhd_ct.cpp:45-77) in one pass.  Thread per
// (k in [ks,ke+1], j in [js,je+1], i in [is,ie+1]); the cell update runs where all three
// indices are active, each face update where its own range allows.  SAVE=true also performs
// CopyCons (u1<-u0, b1<-b0, mhd_tasks.cpp:162-170) for the entries it touches; the caller
// copies the ghost entries separately when the host contract needs them.
template <bool DO_U, bool DO_B, bool SAVE, bool FLAT>
__global__ void __launch_bounds__(256)
update_ct_kernel(Geom g, double *__restrict__ u0, double *__restrict__ u1,
                 const double *__restrict__ flx1, const double *__restrict__ flx2,
                 const double *__restrict__ flx3, double *__restrict__ b1f, double *__restrict__ b2f,
                 double *__restrict__ b3f, double *__restrict__ o1f, double *__restrict__ o2f,
                 double *__restrict__ o3f, const double *__restrict__ e1, const double *__restrict__ e2,
                 const double *__restrict__ e3, const double *__restrict__ mb_size, double gam0, double gam1,
                 double beta_dt) {
  // the conserved update alone covers cells only: no "+1" positions (exact tiles for power-of-two blocks)
  const int xb = DO_B ? 1 : 0;
  int ii, jj, kk, m;
  if (!stencil_index<FLAT>(g.nx1 + xb, g.nx2 + xb, g.nx3 + xb, ii, jj, kk, m)) return;
  const int i = g.is + ii, j = g.js + jj, k = g.ks + kk;
  const double dx1 = __ldg(mb_size + m * 9 + 6), dx2 = __ldg(mb_size + m * 9 + 7), dx3 = __ldg(mb_size + m * 9 + 8);
  const double rdx1 = drcp(dx1), rdx2 = drcp(dx2), rdx3 = drcp(dx3);
  const long n1 = g.n1, n2 = g.n2;
  const bool ia = (i <= g.ie), ja = (j <= g.je), ka = (k <= g.ke);
  if (DO_U && ia && ja && ka) {
    const long o = ((long)k * n2 + j) * n1 + i;
    const long o1 = ((long)k * n2 + j) * (n1 + 1) + i;
    const long o2 = ((long)k * (n2 + 1) + j) * n1 + i;
    const long o3 = o;
#if AKB_UPD_FORM == 0
    // Deferred range guard (mhd_device.cuh): flux differences are exact zeros wherever the flow is uniform, so the guard must
    // not branch on them; the rare out-of-range operand repeats the three quotients with the guarded operator.  (Measured
    // against a form that gathers all five variables before one check per cell: 0.86 vs 0.98 ms, this one wins.)
    auto update_var = [&](int n) {
      const double *p1 = flx1 + ((long)m * g.nvar + n) * g.f1 + o1;
      const double *p2 = flx2 + ((long)m * g.nvar + n) * g.f2 + o2;
      const double *p3 = flx3 + ((long)m * g.nvar + n) * g.f3 + o3;
      const double df1 = __ldg(p1 + 1) - __ldg(p1), df2 = g.multi_d ? __ldg(p2 + n1) - __ldg(p2) : 0.0;
      const double df3 = g.three_d ? __ldg(p3 + n1 * n2) - __ldg(p3) : 0.0;
      bool oor = false;
      double divf = UPD_DIV(df1, dx1, rdx1, &oor);
      if (g.multi_d) divf += UPD_DIV(df2, dx2, rdx2, &oor);   // hydro_update.cpp:57-75
      if (g.three_d) divf += UPD_DIV(df3, dx3, rdx3, &oor);
      if (oor) {
        divf = div_by(df1, dx1, rdx1);
        if (g.multi_d) divf += div_by(df2, dx2, rdx2);
        if (g.three_d) divf += div_by(df3, dx3, rdx3);
      }
      const long uo = ((long)m * g.nvar + n) * g.cc + o;
      const double uold = u0[uo];
      double ureg;
      if (SAVE) { ureg = uold; u1[uo] = uold; } else { ureg = u1[uo]; }   // SAVE: CopyCons fused in
      u0[uo] = gam0 * uold + gam1 * ureg - beta_dt * divf;
    };
#pragma unroll
    for (int n = 0; n < 4; ++n) update_var(n);
    if (g.nfl == 5) update_var(IEN);                        // no energy equation with the isothermal EOS
    for (int n = g.nfl; n < g.nvar; ++n) update_var(n);   // passive scalars (mhd_update.cpp:51: n over nmhd+nscalars)
#else
    // Deferred range guard (mhd_device.cuh): flux differences are exact zeros wherever the flow is uniform, so the guard must not
    // branch on them.  All variables first (no branch between their loads, so every load of the cell is in flight together),
    // ONE check for the cell, and the rare out-of-range operand repeats the quotients with the guarded operator.
    auto divergence = [&](int n, bool *oor) -> double {
      const double *p1 = flx1 + ((long)m * g.nvar + n) * g.f1 + o1;
      const double *p2 = flx2 + ((long)m * g.nvar + n) * g.f2 + o2;
      const double *p3 = flx3 + ((long)m * g.nvar + n) * g.f3 + o3;
      double divf = (oor ? UPD_DIV(__ldg(p1 + 1) - __ldg(p1), dx1, rdx1, oor) : div_by(__ldg(p1 + 1) - __ldg(p1), dx1, rdx1));
      if (g.multi_d) divf += (oor ? UPD_DIV(__ldg(p2 + n1) - __ldg(p2), dx2, rdx2, oor) : div_by(__ldg(p2 + n1) - __ldg(p2), dx2, rdx2));   // hydro_update.cpp:57-75
      if (g.three_d) divf += (oor ? UPD_DIV(__ldg(p3 + n1 * n2) - __ldg(p3), dx3, rdx3, oor) : div_by(__ldg(p3 + n1 * n2) - __ldg(p3), dx3, rdx3));
      return divf;
    };
    // the state loads do not wait for the guard's branch either: gam0*u0 + gam1*u1 is formed with the divergences
    auto combine_var = [&](int n) -> double {
      const long uo = ((long)m * g.nvar + n) * g.cc + o;
      const double uold = u0[uo];
      double ureg;
      if (SAVE) { ureg = uold; u1[uo] = uold; } else { ureg = u1[uo]; }   // SAVE: CopyCons fused in
      return gam0 * uold + gam1 * ureg;
    };
    const bool energy = (g.nfl == 5);                       // no energy equation with the isothermal EOS
    bool oor = false;
    double dv[5], uc[5];
#pragma unroll
    for (int n = 0; n < 4; ++n) { dv[n] = divergence(n, &oor); uc[n] = combine_var(n); }
    dv[4] = 0.0; uc[4] = 0.0;
    if (energy) { dv[4] = divergence(IEN, &oor); uc[4] = combine_var(IEN); }
    if (oor) {
#pragma unroll
      for (int n = 0; n < 4; ++n) dv[n] = divergence(n, nullptr);
      if (energy) dv[4] = divergence(IEN, nullptr);
    }
    double *ub = u0 + (long)m * g.nvar * g.cc + o;
#pragma unroll
    for (int n = 0; n < 4; ++n) ub[n * g.cc] = uc[n] - beta_dt * dv[n];
    if (energy) ub[IEN * g.cc] = uc[4] - beta_dt * dv[4];
    for (int n = g.nfl; n < g.nvar; ++n) {                  // passive scalars (mhd_update.cpp:51: n over nmhd+nscalars)
      const double c = combine_var(n);
      ub[n * g.cc] = c - beta_dt * divergence(n, nullptr);
    }
  #endif
}
  if (DO_B) {
    const double *E1 = e1 + (long)m * g.e1, *E2 = e2 + (long)m * g.e2, *E3 = e3 + (long)m * g.e3;
    const long a1 = ((long)k * (n2 + 1) + j) * n1 + i;        // x1e(k,j,i)
    const long a2 = ((long)k * n2 + j) * (n1 + 1) + i;        // x2e(k,j,i)
    const long a3 = ((long)k * (n2 + 1) + j) * (n1 + 1) + i;  // x3e(k,j,i)
    if (ja && ka && g.multi_d) {  // x1f(k,j,i), i in [is,ie+1]; B1 is not touched in 1-D (mhd_ct.cpp:41-53)
      const long o = (long)m * g.f1 + ((long)k * n2 + j) * (n1 + 1) + i;
      const double bold = b1f[o];
      double breg;
      if (SAVE) { breg = bold; o1f[o] = bold; } else { breg = o1f[o]; }
      double b = gam0 * bold + gam1 * breg;
      const double de3 = beta_dt * (__ldg(E3 + a3 + (n1 + 1)) - __ldg(E3 + a3));
      const double de2 = g.three_d ? beta_dt * (__ldg(E2 + a2 + n2 * (n1 + 1)) - __ldg(E2 + a2)) : 0.0;
      bool oor = false;
      double q3 = div_by(de3, dx2, rdx2, &oor), q2 = g.three_d ? div_by(de2, dx3, rdx3, &oor) : 0.0;
      if (oor) { q3 = div_by(de3, dx2, rdx2); if (g.three_d) q2 = div_by(de2, dx3, rdx3); }
      b -= q3;
      if (g.three_d) b += q2;
      b1f[o] = b;
    }
    if (ia && ka) {  // x2f(k,j,i), j in [js,je+1]
      const long o = (long)m * g.f2 + ((long)k * (n2 + 1) + j) * n1 + i;
      const double bold = b2f[o];
      double breg;
      if (SAVE) { breg = bold; o2f[o] = bold; } else { breg = o2f[o]; }
      double b = gam0 * bold + gam1 * breg;
      const double de3 = beta_dt * (__ldg(E3 + a3 + 1) - __ldg(E3 + a3));
      const double de1 = g.three_d ? beta_dt * (__ldg(E1 + a1 + (n2 + 1) * n1) - __ldg(E1 + a1)) : 0.0;
      bool oor = false;
      double q3 = div_by(de3, dx1, rdx1, &oor), q1 = g.three_d ? div_by(de1, dx3, rdx3, &oor) : 0.0;
      if (oor) { q3 = div_by(de3, dx1, rdx1); if (g.three_d) q1 = div_by(de1, dx3, rdx3); }
      b += q3;
      if (g.three_d) b -= q1;
      b2f[o] = b;
    }
    if (ia && ja) {  // x3f(k,j,i), k in [ks,ke+1]
      const long o = (long)m * g.f3 + ((long)k * n2 + j) * n1 + i;
      const double bold = b3f[o];
      double breg;
      if (SAVE) { breg = bold; o3f[o] = bold; } else { breg = o3f[o]; }
      double b = gam0 * bold + gam1 * breg;
      const double de2 = beta_dt * (__ldg(E2 + a2 + 1) - __ldg(E2 + a2));
      const double de1 = g.multi_d ? beta_dt * (__ldg(E1 + a1 + n1) - __ldg(E1 + a1)) : 0.0;
      bool oor = false;
      double q2 = div_by(de2, dx1, rdx1, &oor), q1 = g.multi_d ? div_by(de1, dx2, rdx2, &oor) : 0.0;
      if (oor) { q2 = div_by(de2, dx1, rdx1); if (g.multi_d) q1 = div_by(de1, dx2, rdx2); }
      b -= q2;
      if (g.multi_d) b += q1;
      b3f[o] = b;
    }
  }
}

// ---------------------------------------------------------------------------------------
// ConsToPrim (ideal_mhd.cpp:47-122 + ideal_c2p_mhd.hpp:20-67) over a cell box, optionally
// fused with the new-dt reduction over the active cells (mhd_newdt.cpp:76-150).
// counters: int[3] (dfloor,efloor,tfloor); dtmin: 3 doubles stored as ordered uint64 bits.
template <bool NEWDT, bool IDEAL>
__global__ void __launch_bounds__(256)
cons2prim_kernel(Geom g, EosParams eos, double *__restrict__ u0, const double *__restrict__ b1f,
                 const double *__restrict__ b2f, const double *__restrict__ b3f, double *__restrict__ w0,
                 double *__restrict__ bcc, int il, int iu, int jl, int ju, int kl, int ku,
                 int *__restrict__ counters, const double *__restrict__ mb_size,
                 unsigned long long *__restrict__ dtmin) {
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ni = iu - il + 1, nj = ju - jl + 1, nk = ku - kl + 1;
  const long total = (long)g.nmb * nk * nj * ni;
  int nd = 0, ne = 0, nt = 0;
  double d1 = DT_IDENT, d2 = DT_IDENT, d3 = DT_IDENT;
  bool act = false;
  if (gt < total) {
    long t = gt;
    const int i = il + (int)(t % ni); t /= ni;
    const int j = jl + (int)(t % nj); t /= nj;
    const int k = kl + (int)(t % nk);
    const int m = (int)(t / nk);
    const long n1 = g.n1, n2 = g.n2, cc = g.cc;
    const long o = ((long)k * n2 + j) * n1 + i;
    double *u = u0 + (long)m * g.nvar * cc + o;
    double ud = u[IDN * cc], umx = u[IM1 * cc], umy = u[IM2 * cc], umz = u[IM3 * cc], ue = 0.0;
    if constexpr (IDEAL) ue = u[IEN * cc];
    const double *p1 = b1f + (long)m * g.f1 + ((long)k * n2 + j) * (n1 + 1) + i;
    const double *p2 = b2f + (long)m * g.f2 + ((long)k * (n2 + 1) + j) * n1 + i;
    const double *p3 = b3f + (long)m * g.f3 + o;
    const double ubx = 0.5 * (__ldg(p1) + __ldg(p1 + 1));
    const double uby = 0.5 * (__ldg(p2) + __ldg(p2 + n1));
    const double ubz = 0.5 * (__ldg(p3) + __ldg(p3 + n1 * n2));
    // one cell of c2p (c2p_device.cuh), deferred range guard: the rare out-of-range operand repeats the cell with the guarded operators
    bool dfl = false, efl = false, tfl = false, oor = false;
    const double ud0 = ud, ue0 = ue;
    double wd, wvx, wvy, wvz, we;
    c2p_cell<IDEAL>(eos, ud, umx, umy, umz, ue, ubx, uby, ubz, wd, wvx, wvy, wvz, we, dfl, efl, tfl, &oor);
    if (oor) {
      ud = ud0; ue = ue0; dfl = efl = tfl = false;
      c2p_cell<IDEAL>(eos, ud, umx, umy, umz, ue, ubx, uby, ubz, wd, wvx, wvy, wvz, we, dfl, efl, tfl, nullptr);
    }
    double *w = w0 + (long)m * g.nvar * cc + o;
    if constexpr (IDEAL) {
      if (efl) { u[IEN * cc] = ue; ne = 1; }
      if (tfl) { u[IEN * cc] = ue; nt = 1; }
      w[IEN * cc] = we;
    }
    if (dfl) { u[IDN * cc] = ud; nd = 1; }
    w[IDN * cc] = wd; w[IVX * cc] = wvx; w[IVY * cc] = wvy; w[IVZ * cc] = wvz;
    double *bc = bcc + (long)m * 3 * cc + o;
    bc[0] = ubx; bc[cc] = uby; bc[2 * cc] = ubz;
    for (int n = g.nfl; n < g.nvar; ++n) {   // ideal_mhd.cpp:113-120: scalar floor, then s / d (floored density)
      double sc = u[n * cc];
      if (IDEAL && sc < 0.0) { sc = 0.0; u[n * cc] = 0.0; }   // isothermal_mhd.cpp:130-132: no scalar floor
      w[n * cc] = sc / ud;
    }
    if (NEWDT) {
      act = (i >= g.is && i <= g.ie && j >= g.js && j <= g.je && k >= g.ks && k <= g.ke);
      if (act) {  // mhd_newdt.cpp:128-138,146-148
        bool o2 = false;
        const double dx1 = __ldg(mb_size + m * 9 + 6), dx2 = __ldg(mb_size + m * 9 + 7), dx3 = __ldg(mb_size + m * 9 + 8);
        newdt_cell<IDEAL>(eos, dx1, dx2, dx3, wd, wvx, wvy, wvz, we, ubx, uby, ubz, d1, d2, d3, &o2);
        if (o2) newdt_cell<IDEAL>(eos, dx1, dx2, dx3, wd, wvx, wvy, wvz, we, ubx, uby, ubz, d1, d2, d3, nullptr);
      }
    }
  }
  // warp-level reduction, one atomic per warp and only when needed
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
  if (NEWDT) block_min3(d1, d2, d3, dtmin);
}

// MHD::NewTimeStep stand-alone (task 28): reads w0, bcc0 over the active cells.
template <bool IDEAL>
__global__ void __launch_bounds__(256)
newdt_kernel(Geom g, EosParams eos, const double *__restrict__ w0, const double *__restrict__ bcc,
             const double *__restrict__ mb_size, unsigned long long *__restrict__ dtmin) {
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)g.nmb * g.nx3 * g.nx2 * g.nx1;
  double d1 = DT_IDENT, d2 = DT_IDENT, d3 = DT_IDENT;
  if (gt < total) {
    long t = gt;
    const int i = g.is + (int)(t % g.nx1); t /= g.nx1;
    const int j = g.js + (int)(t % g.nx2); t /= g.nx2;
    const int k = g.ks + (int)(t % g.nx3);
    const int m = (int)(t / g.nx3);
    const long cc = g.cc, o = ((long)k * g.n2 + j) * g.n1 + i;
    const double *w = w0 + (long)m * g.nvar * cc + o, *b = bcc + (long)m * 3 * cc + o;
    const double wd = __ldg(w), bx = __ldg(b), by = __ldg(b + cc), bz = __ldg(b + 2 * cc);
    double p = 0.0;
    if constexpr (IDEAL) p = (eos.gamma - 1.0) * __ldg(w + IEN * cc);
    auto cf = [&](double b1, double b2, double b3) {
      if constexpr (IDEAL) return fast_speed(eos.gamma, wd, p, b1, b2, b3);
      else return fast_speed_iso(eos.iso_cs, wd, b1, b2, b3);
    };
    d1 = ddiv(__ldg(mb_size + m * 9 + 6), fabs(__ldg(w + IVX * cc)) + cf(bx, by, bz));
    d2 = ddiv(__ldg(mb_size + m * 9 + 7), fabs(__ldg(w + IVY * cc)) + cf(by, bz, bx));
    d3 = ddiv(__ldg(mb_size + m * 9 + 8), fabs(__ldg(w + IVZ * cc)) + cf(bz, bx, by));
  }
  block_min3(d1, d2, d3, dtmin);
}

// u1 <- u0 / b1 <- b0 fused into the stage-1 update needs the ghost entries too when the host
// contract reads them; a plain vectorised copy covers that case (CopyCons, task 1).
__global__ void __launch_bounds__(256) copy_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst, long n2,
                                                    const double *__restrict__ tail_src, double *__restrict__ tail_dst, int ntail) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < n2; q += stride) dst[q] = src[q];
  if (blockIdx.x == 0 && threadIdx.x < ntail) tail_dst[threadIdx.x] = tail_src[threadIdx.x];
}

// ---------------------------------------------------------------------------------------
// launchers
template <int RECON>
static cudaError_t launch_fluxes_r(int rs, const Geom &g, const EosParams &eos, const Arrays &a, cudaStream_t st) {
  switch (rs) {
    case RS_LLF: return launch_fluxes_t<RECON, RS_LLF>(g, eos, a, st);
    case RS_HLLE: return launch_fluxes_t<RECON, RS_HLLE>(g, eos, a, st);
    case RS_HLLD: return launch_fluxes_t<RECON, RS_HLLD>(g, eos, a, st);
    case RS_LLF_ISO: return launch_fluxes_t<RECON, RS_LLF_ISO>(g, eos, a, st);
    case RS_HLLE_ISO: return launch_fluxes_t<RECON, RS_HLLE_ISO>(g, eos, a, st);
    case RS_HLLD_ISO: return launch_fluxes_t<RECON, RS_HLLD_ISO>(g, eos, a, st);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_fluxes(int recon, int rs, const Geom &g, const EosParams &eos, const Arrays &a, cudaStream_t st) {
  switch (recon) {
    case RECON_DC: return launch_fluxes_r<RECON_DC>(rs, g, eos, a, st);
    case RECON_PLM: return launch_fluxes_r<RECON_PLM>(rs, g, eos, a, st);
    case RECON_PPM4: return launch_fluxes_r<RECON_PPM4>(rs, g, eos, a, st);
    case RECON_PPMX: return launch_fluxes_r<RECON_PPMX>(rs, g, eos, a, st);
    case RECON_WENOZ: return launch_fluxes_r<RECON_WENOZ>(rs, g, eos, a, st);
    case RECON_TENO: return launch_fluxes_r<RECON_TENO>(rs, g, eos, a, st);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_corner_e(const Geom &g, const Arrays &a, cudaStream_t st) {
  if (!g.multi_d) {
    corner_emf_1d_kernel<<<nblocks((long)g.nmb * (g.nx1 + 1), 128), 128, 0, st>>>(g, a.e3x1, a.e2x1, a.e2, a.e3);
    return cudaGetLastError();
  }
  if (!g.three_d) {
    const dim3 gr2((g.nx1 + 1 + 31) / 32, (g.nx2 + 1 + 3) / 4, g.nmb);
    corner_emf_2d_kernel<<<gr2, dim3(32, 4, 1), 0, st>>>(g, a.w0, a.bcc0, a.flx1, a.flx2, a.e3x1, a.e2x1, a.e1x2, a.e3x2, a.e1, a.e2, a.e3);
    return cudaGetLastError();
  }
  const bool flat = stencil_flat(g.nx1 + 1, g.nx2 + 1, g.nx3 + 1);
  const dim3 gr = stencil_grid(flat, g.nx1 + 1, g.nx2 + 1, g.nx3 + 1, g.nmb);
  if (flat) corner_emf_kernel<true><<<gr, 256, 0, st>>>(g, a.w0, a.bcc0, a.flx1, a.flx2, a.flx3, a.e3x1, a.e2x1, a.e1x2, a.e3x2, a.e2x3, a.e1x3, a.e1, a.e2, a.e3);
  else corner_emf_kernel<false><<<gr, 256, 0, st>>>(g, a.w0, a.bcc0, a.flx1, a.flx2, a.flx3, a.e3x1, a.e2x1, a.e1x2, a.e3x2, a.e2x3, a.e1x3, a.e1, a.e2, a.e3);
  return cudaGetLastError();
}

cudaError_t launch_update_ct(const Geom &g, const Arrays &a, bool do_u, bool do_b, bool save, double gam0, double gam1,
                             double beta_dt, cudaStream_t st) {
  const int xb = do_b ? 1 : 0;
  const int ni = g.nx1 + xb, nj = g.nx2 + xb, nk = g.nx3 + xb;    // CT reaches the +1 faces in every dimension (mhd_ct.cpp:41-77)
  const bool flat = stencil_flat(ni, nj, nk);
  const dim3 gr = stencil_grid(flat, ni, nj, nk, g.nmb);
  const int tb = 256;
#define AKB_UCT(U, B, S) do { if (flat) update_ct_kernel<U, B, S, true><<<gr, tb, 0, st>>>(AKB_UCT_ARGS); else update_ct_kernel<U, B, S, false><<<gr, tb, 0, st>>>(AKB_UCT_ARGS); } while (0)
#define AKB_UCT_ARGS g, a.u0, a.u1, a.flx1, a.flx2, a.flx3, a.b0x1f, a.b0x2f, a.b0x3f, a.b1x1f, a.b1x2f, a.b1x3f, a.e1, a.e2, a.e3, a.mb_size, gam0, gam1, beta_dt
  if (do_u && do_b && save) AKB_UCT(true, true, true);
  else if (do_u && do_b) AKB_UCT(true, true, false);
  else if (do_u && save) AKB_UCT(true, false, true);   // hydro twin, stage 1
  else if (do_u) AKB_UCT(true, false, false);
  else if (do_b && save) AKB_UCT(false, true, true);   // CT after an overlapped U update, stage 1
  else if (do_b) AKB_UCT(false, true, false);
#undef AKB_UCT
#undef AKB_UCT_ARGS
  return cudaGetLastError();
}

cudaError_t launch_cons2prim(const Geom &g, const EosParams &eos, const Arrays &a, int il, int iu, int jl, int ju,
                             int kl, int ku, int *counters, bool newdt, unsigned long long *dtmin, cudaStream_t st) {
  const long n = (long)g.nmb * (ku - kl + 1) * (ju - jl + 1) * (iu - il + 1);
  if (n <= 0) return cudaSuccess;
#define AKB_C2P_ARGS g, eos, a.u0, a.b0x1f, a.b0x2f, a.b0x3f, a.w0, a.bcc0, il, iu, jl, ju, kl, ku, counters, a.mb_size, dtmin
  const bool ideal = (g.nfl == 5);
  if (newdt && ideal) cons2prim_kernel<true, true><<<nblocks(n, 256), 256, 0, st>>>(AKB_C2P_ARGS);
  else if (ideal) cons2prim_kernel<false, true><<<nblocks(n, 256), 256, 0, st>>>(AKB_C2P_ARGS);
  else if (newdt) cons2prim_kernel<true, false><<<nblocks(n, 256), 256, 0, st>>>(AKB_C2P_ARGS);
  else cons2prim_kernel<false, false><<<nblocks(n, 256), 256, 0, st>>>(AKB_C2P_ARGS);
#undef AKB_C2P_ARGS
  return cudaGetLastError();
}

cudaError_t launch_newdt(const Geom &g, const EosParams &eos, const Arrays &a, unsigned long long *dtmin, cudaStream_t st) {
  const long n = (long)g.nmb * g.nx3 * g.nx2 * g.nx1;
  if (g.nfl == 5) newdt_kernel<true><<<nblocks(n, 256), 256, 0, st>>>(g, eos, a.w0, a.bcc0, a.mb_size, dtmin);
  else newdt_kernel<false><<<nblocks(n, 256), 256, 0, st>>>(g, eos, a.w0, a.bcc0, a.mb_size, dtmin);
  return cudaGetLastError();
}

int set_march_tma(int on) { int &f = march_tma_flag(); const int old = f; f = on ? 1 : 0; return old; }

cudaError_t launch_copy(const double *src, double *dst, long n, cudaStream_t st) {
  // 16-byte vector copy when both pointers allow it, scalar tail otherwise
  if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
    const long n2 = n / 2;
    const int ntail = (int)(n - 2 * n2);
    long want = (n2 + 255) / 256;
    unsigned nb = (unsigned)(want < 148 * 16 ? (want > 0 ? want : 1) : 148 * 16);
    copy_kernel<<<nb, 256, 0, st>>>((const double2 *)src, (double2 *)dst, n2, src + 2 * n2, dst + 2 * n2, ntail);
    return cudaGetLastError();
  }
  return cudaMemcpyAsync(dst, src, (size_t)n * 8, cudaMemcpyDeviceToDevice, st);
}

}  // namespace akb
