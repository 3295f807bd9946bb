// fofc.cu -- first-order flux correction for MHD (reference
// src/mhd/mhd_fofc.cpp:31-489, Newtonian branch without excision; called from MHD::Fluxes,
// mhd_tasks.cpp:208-211, after CalculateFluxes).
//
//   fofc_flag_kernel   one thread per cell of [is-1,ie+1] x [js-1,je+1] x [ks-1,ke+1] (active
//                      dimensions): estimate of the updated conserved state and cell-centred field
//                      (mhd_fofc.cpp:70-108) and the only_testfloors branch of ConsToPrim
//                      (ideal_mhd.cpp:57-88, ideal_c2p_mhd.hpp:20-67; isothermal_mhd.cpp:95-118)
//                      folded into one pass: no utest / bcctest arrays, one flag byte per cell.
//   fofc_face_kernel   one thread per face of that box and direction: a face with a flagged cell
//                      on either side is recomputed with the single-state LLF solver from the two
//                      cell-centred states (gather form of the reference's two scatter passes
//                      mhd_fofc.cpp:131-474: both passes store the same value on a shared face).
// The sweep kernels (flux_kernels.cuh) compute fluxes and face EMFs one cell further along the
// face normal when FOFC is on (Geom::next; mhd_fluxes.cpp:96-106).
#include "iso_device.cuh"
#include "mhd_device.cuh"
#include "mhd_kernels.h"

namespace akb {

template <bool IDEAL>
__global__ void __launch_bounds__(256)
fofc_flag_kernel(Geom g, EosParams eos, const double *__restrict__ u0, const double *__restrict__ u1,
                 const double *__restrict__ bcc0, const double *__restrict__ b0x1f, const double *__restrict__ b0x2f,
                 const double *__restrict__ b0x3f, const double *__restrict__ b1x1f, const double *__restrict__ b1x2f,
                 const double *__restrict__ b1x3f, const double *__restrict__ flx1, const double *__restrict__ flx2,
                 const double *__restrict__ flx3, const double *__restrict__ e3x1, const double *__restrict__ e2x1,
                 const double *__restrict__ e1x2, const double *__restrict__ e3x2, const double *__restrict__ e2x3,
                 const double *__restrict__ e1x3, const double *__restrict__ mb_size, double gam0, double gam1,
                 double beta_dt, unsigned char *__restrict__ flag, int *__restrict__ counter) {
  const int ni = g.nx1 + 2, nj = g.nx2 + 2 * g.multi_d, nk = g.nx3 + 2 * g.three_d;
  const long total = (long)g.nmb * nk * nj * ni;
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  int used = 0;
  if (gt < total) {
    long t = gt;
    const int i = g.is - 1 + (int)(t % ni); t /= ni;
    const int j = g.js - g.multi_d + (int)(t % nj); t /= nj;
    const int k = g.ks - g.three_d + (int)(t % nk);
    const int m = (int)(t / nk);
    const long n1 = g.n1, n2 = g.n2, cc = g.cc;
    const long o = ((long)k * n2 + j) * n1 + i;
    const long o1 = ((long)k * n2 + j) * (n1 + 1) + i;        // x1-face (k,j,i)
    const long o2 = ((long)k * (n2 + 1) + j) * n1 + i;        // x2-face (k,j,i)
    const double dtodx1 = ddiv(beta_dt, __ldg(mb_size + m * 9 + 6));
    const double dtodx2 = ddiv(beta_dt, __ldg(mb_size + m * 9 + 7));
    const double dtodx3 = ddiv(beta_dt, __ldg(mb_size + m * 9 + 8));
    double ut[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int n = 0; n < 5; ++n) {
      if (n < g.nfl) {   // mhd_fofc.cpp:78-88
        const double *p1 = flx1 + ((long)m * g.nvar + n) * g.f1 + o1;
        double divf = dtodx1 * (__ldg(p1 + 1) - __ldg(p1));
        if (g.multi_d) {
          const double *p2 = flx2 + ((long)m * g.nvar + n) * g.f2 + o2;
          divf += dtodx2 * (__ldg(p2 + n1) - __ldg(p2));
        }
        if (g.three_d) {
          const double *p3 = flx3 + ((long)m * g.nvar + n) * g.f3 + o;
          divf += dtodx3 * (__ldg(p3 + n1 * n2) - __ldg(p3));
        }
        const long uo = ((long)m * g.nvar + n) * cc + o;
        ut[n] = gam0 * __ldg(u0 + uo) + gam1 * __ldg(u1 + uo) - divf;
      }
    }
    if constexpr (IDEAL) {
      // estimate of the cell-centred field (mhd_fofc.cpp:91-107); x2f/x3f keep a j+1 / k+1 row in
      // inactive dimensions too (extent n2+1 / n3+1)
      const double *q1 = b1x1f + (long)m * g.f1 + o1, *q2 = b1x2f + (long)m * g.f2 + o2, *q3 = b1x3f + (long)m * g.f3 + o;
      const double b1old = 0.5 * (__ldg(q1) + __ldg(q1 + 1));
      const double b2old = 0.5 * (__ldg(q2) + __ldg(q2 + n1));
      const double b3old = 0.5 * (__ldg(q3) + __ldg(q3 + n1 * n2));
      const double *bc = bcc0 + (long)m * 3 * cc + o;
      double bx = gam0 * __ldg(bc) + gam1 * b1old;
      double by = gam0 * __ldg(bc + cc) + gam1 * b2old;
      double bz = gam0 * __ldg(bc + 2 * cc) + gam1 * b3old;
      const long mo = (long)m * cc + o;
      by += dtodx1 * (__ldg(e3x1 + mo + 1) - __ldg(e3x1 + mo));
      bz -= dtodx1 * (__ldg(e2x1 + mo + 1) - __ldg(e2x1 + mo));
      if (g.multi_d) {
        bx -= dtodx2 * (__ldg(e3x2 + mo + n1) - __ldg(e3x2 + mo));
        bz += dtodx2 * (__ldg(e1x2 + mo + n1) - __ldg(e1x2 + mo));
      }
      if (g.three_d) {
        bx += dtodx3 * (__ldg(e2x3 + mo + n1 * n2) - __ldg(e2x3 + mo));
        by -= dtodx3 * (__ldg(e1x3 + mo + n1 * n2) - __ldg(e1x3 + mo));
      }
      // SingleC2P_IdealMHD on (utest, bcctest): which floors would fire
      const double gm1 = eos.gamma - 1.0;
      const double b2 = sqr(bx) + sqr(by) + sqr(bz);
      const double dfloor_ = dmax(eos.dfloor, ddiv(b2, eos.sigma_max));
      const double efloor = ddiv(eos.pfloor, eos.gamma - 1.0);
      double ud = ut[IDN];
      if (ud < dfloor_) { ud = dfloor_; used = 1; }
      const double wd = ud;
      const double di = drcp(ud);
      const double e_k = 0.5 * di * (sqr(ut[IM1]) + sqr(ut[IM2]) + sqr(ut[IM3]));
      const double e_m = 0.5 * (sqr(bx) + sqr(by) + sqr(bz));
      double we = (ut[IEN] - e_k - e_m);
      if (we < efloor) { we = efloor; used = 1; }
      if (gm1 * we * di < eos.tfloor) { we = wd * eos.tfloor / gm1; used = 1; }
      // entropy floor: pow() only when it can fire (same bound as cons2prim_kernel)
      const double pow_ub = dmax(1.0, wd * wd);
      if (!(gm1 > 0.0 && gm1 <= 2.0 && (gm1 * we * di) > 4.0 * eos.sfloor * pow_ub)) {
        const double spe_over_eps = gm1 / pow(wd, gm1);
        const double spe = spe_over_eps * we * di;
        if (spe <= eos.sfloor) used = 1;
      }
    } else {
      // isothermal_mhd.cpp:95-118: the floor test averages the FACE fields it is handed (b0), not bcctest
      const double *q1 = b0x1f + (long)m * g.f1 + o1, *q2 = b0x2f + (long)m * g.f2 + o2, *q3 = b0x3f + (long)m * g.f3 + o;
      const double bx = 0.5 * (__ldg(q1) + __ldg(q1 + 1));
      const double by = 0.5 * (__ldg(q2) + __ldg(q2 + n1));
      const double bz = 0.5 * (__ldg(q3) + __ldg(q3 + n1 * n2));
      const double b2 = sqr(bx) + sqr(by) + sqr(bz);
      if (ut[IDN] < dmax(eos.dfloor, ddiv(b2, eos.sigma_max))) used = 1;
    }
    flag[(long)m * cc + o] = (unsigned char)used;
  }
  const unsigned mask = __ballot_sync(0xffffffffu, used);
  if (mask && (threadIdx.x & 31) == 0) atomicAdd(counter, __popc(mask));
}

// DIR: 0 x1-faces i in [is-1,ie+2], 1 x2-faces j in [js-1,je+2], 2 x3-faces k in [ks-1,ke+2]; the
// other two indices run over the flagged-cell box.
template <int DIR, bool IDEAL>
__global__ void __launch_bounds__(256)
fofc_face_kernel(Geom g, EosParams eos, const unsigned char *__restrict__ flag, const double *__restrict__ w0,
                 const double *__restrict__ bcc0, const double *__restrict__ bnf, double *__restrict__ flx,
                 double *__restrict__ eyo, double *__restrict__ ezo) {
  const int il = g.is - 1, jl = g.js - g.multi_d, kl = g.ks - g.three_d;
  const int ci = g.nx1 + 2, cj = g.nx2 + 2 * g.multi_d, ck = g.nx3 + 2 * g.three_d;   // flagged-cell box
  const int ni = ci + (DIR == 0), nj = cj + (DIR == 1), nk = ck + (DIR == 2);
  const long total = (long)g.nmb * nk * nj * ni;
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int i = il + (int)(t % ni); t /= ni;
  const int j = jl + (int)(t % nj); t /= nj;
  const int k = kl + (int)(t % nk);
  const int m = (int)(t / nk);
  const long n1 = g.n1, n2 = g.n2, cc = g.cc;
  const long s = (DIR == 0) ? 1 : (DIR == 1) ? n1 : n1 * n2;
  const long o = ((long)k * n2 + j) * n1 + i;         // upper cell of the face
  const int a = (DIR == 0) ? i - il : (DIR == 1) ? j - jl : k - kl;       // position along the normal
  const int na = (DIR == 0) ? ci : (DIR == 1) ? cj : ck;
  const unsigned char *fl = flag + (long)m * cc;
  const bool hit = (a >= 1 && fl[o - s]) || (a < na && fl[o]);
  if (!hit) return;
  constexpr int ivx = IVX + DIR, ivy = IVX + (DIR + 1) % 3, ivz = IVX + (DIR + 2) % 3;
  constexpr int iby = (DIR + 1) % 3, ibz = (DIR + 2) % 3;
  const double *w = w0 + (long)m * g.nvar * cc + o, *b = bcc0 + (long)m * 3 * cc + o;
  State7 l, r;
  l.d = __ldg(w + IDN * cc - s); l.vx = __ldg(w + ivx * cc - s); l.vy = __ldg(w + ivy * cc - s); l.vz = __ldg(w + ivz * cc - s);
  r.d = __ldg(w + IDN * cc); r.vx = __ldg(w + ivx * cc); r.vy = __ldg(w + ivy * cc); r.vz = __ldg(w + ivz * cc);
  l.e = r.e = 0.0;
  if constexpr (IDEAL) { l.e = __ldg(w + IEN * cc - s); r.e = __ldg(w + IEN * cc); }
  l.by = __ldg(b + iby * cc - s); l.bz = __ldg(b + ibz * cc - s);
  r.by = __ldg(b + iby * cc); r.bz = __ldg(b + ibz * cc);
  const long fo = (DIR == 0) ? ((long)k * n2 + j) * (n1 + 1) + i : (DIR == 1) ? ((long)k * (n2 + 1) + j) * n1 + i : o;
  const long fsz = (DIR == 0) ? g.f1 : (DIR == 1) ? g.f2 : g.f3;
  const double bxi = __ldg(bnf + (long)m * fsz + fo);
  Flux7 f;
  if constexpr (IDEAL) llf(l, r, bxi, eos, f); else llf_iso(l, r, bxi, eos, f);
  double *fp = flx + (long)m * g.nvar * fsz + fo;
  fp[IDN * fsz] = f.d; fp[ivx * fsz] = f.mx; fp[ivy * fsz] = f.my; fp[ivz * fsz] = f.mz;
  if constexpr (IDEAL) fp[IEN * fsz] = f.e;
  eyo[(long)m * cc + o] = f.ey;     // x1: (e3x1,e2x1)  x2: (e1x2,e3x2)  x3: (e2x3,e1x3); mhd_fofc.cpp:196-197
  ezo[(long)m * cc + o] = f.ez;
}

static inline unsigned nblk(long n) { return (unsigned)((n + 255) / 256); }

template <bool IDEAL>
static cudaError_t launch_fofc_t(const Geom &g, const EosParams &eos, const Arrays &a, double gam0, double gam1,
                                 double beta_dt, unsigned char *flag, int *counter, cudaStream_t st) {
  const long ci = g.nx1 + 2, cj = g.nx2 + 2 * g.multi_d, ck = g.nx3 + 2 * g.three_d;
  fofc_flag_kernel<IDEAL><<<nblk(g.nmb * ck * cj * ci), 256, 0, st>>>(
      g, eos, a.u0, a.u1, a.bcc0, a.b0x1f, a.b0x2f, a.b0x3f, a.b1x1f, a.b1x2f, a.b1x3f, a.flx1, a.flx2, a.flx3, a.e3x1,
      a.e2x1, a.e1x2, a.e3x2, a.e2x3, a.e1x3, a.mb_size, gam0, gam1, beta_dt, flag, counter);
  fofc_face_kernel<0, IDEAL><<<nblk(g.nmb * ck * cj * (ci + 1)), 256, 0, st>>>(g, eos, flag, a.w0, a.bcc0, a.b0x1f, a.flx1, a.e3x1, a.e2x1);
  if (g.multi_d)
    fofc_face_kernel<1, IDEAL><<<nblk(g.nmb * ck * (cj + 1) * ci), 256, 0, st>>>(g, eos, flag, a.w0, a.bcc0, a.b0x2f, a.flx2, a.e1x2, a.e3x2);
  if (g.three_d)
    fofc_face_kernel<2, IDEAL><<<nblk(g.nmb * (ck + 1) * cj * ci), 256, 0, st>>>(g, eos, flag, a.w0, a.bcc0, a.b0x3f, a.flx3, a.e2x3, a.e1x3);
  return cudaGetLastError();
}

cudaError_t launch_fofc(const Geom &g, const EosParams &eos, const Arrays &a, double gam0, double gam1, double beta_dt,
                        unsigned char *flag, int *counter, cudaStream_t st) {
  return (g.nfl == 5) ? launch_fofc_t<true>(g, eos, a, gam0, gam1, beta_dt, flag, counter, st)
                      : launch_fofc_t<false>(g, eos, a, gam0, gam1, beta_dt, flag, counter, st);
}

}  // namespace akb
