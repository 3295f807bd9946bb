// hydro_device.cuh -- per-face Riemann solvers of the Newtonian ideal-gas HYDRO twin of the path
// (reference src/hydro/rsolvers).  Same conventions as mhd_device.cuh: FP64, --fmad=false,
// the reference's expression order, the slow-path-free exact division / sqrt of mhd_device.cuh.
// States travel in State7 / Flux7 with the magnetic members unused (dead code for the compiler).
#pragma once
#include "mhd_device.cuh"

namespace akb {

// hydro solver selectors continue the RSolver enum (include/athenak_b200.h, AK_RSOLVER_HYD_*)
enum RSolverHyd { RS_HYD_ADVECT = 10, RS_HYD_LLF = 11, RS_HYD_HLLE = 12, RS_HYD_HLLC = 13, RS_HYD_ROE = 14 };

// eos.hpp:41-45
__device__ __forceinline__ double sound_speed(double gamma, double d, double p, bool *oor = nullptr) { return dsqrt(ddiv(gamma * p, d, oor), oor); }

// advect_hyd.hpp:18-52 (ideal gas)
__device__ __forceinline__ void hyd_advect(const State7 &l, const State7 &r, const EosParams &eos, Flux7 &f, bool *oor = nullptr) {
  if (l.vx >= 0.0) {
    f.d = l.d * l.vx;
    f.mx = l.d * l.vx * l.vx;
    f.my = l.vy * l.vx;
    f.mz = l.vz * l.vx;
    f.e = l.e * l.vx;
  } else {
    f.d = r.d * r.vx;
    f.mx = r.d * r.vx * r.vx;
    f.my = r.vy * r.vx;
    f.mz = r.vz * r.vx;
    f.e = r.e * r.vx;
  }
}

// llf_hyd_singlestate.hpp:27-82 (ideal gas)
__device__ __forceinline__ void hyd_llf(const State7 &l, const State7 &r, const EosParams &eos, Flux7 &f, bool *oor = nullptr) {
  const double gm1 = eos.gamma - 1.0;
  double qa = l.d * l.vx;
  double qb = r.d * r.vx;
  double s_d = qa + qb;
  double s_mx = qa * l.vx + qb * r.vx;
  double s_my = qa * l.vy + qb * r.vy;
  double s_mz = qa * l.vz + qb * r.vz;
  double pl = gm1 * l.e;
  double pr = gm1 * r.e;
  double el = l.e + 0.5 * l.d * (sqr(l.vx) + sqr(l.vy) + sqr(l.vz));
  double er = r.e + 0.5 * r.d * (sqr(r.vx) + sqr(r.vy) + sqr(r.vz));
  s_mx += (pl + pr);
  double s_e = (el + pl) * l.vx + (er + pr) * r.vx;
  qa = sound_speed(eos.gamma, l.d, pl, oor);
  qb = sound_speed(eos.gamma, r.d, pr, oor);
  double a = dmax((fabs(l.vx) + qa), (fabs(r.vx) + qb));
  double du_d = a * (r.d - l.d);
  double du_mx = a * (r.d * r.vx - l.d * l.vx);
  double du_my = a * (r.d * r.vy - l.d * l.vy);
  double du_mz = a * (r.d * r.vz - l.d * l.vz);
  double du_e = a * (er - el);
  f.d = 0.5 * (s_d - du_d);
  f.mx = 0.5 * (s_mx - du_mx);
  f.my = 0.5 * (s_my - du_my);
  f.mz = 0.5 * (s_mz - du_mz);
  f.e = 0.5 * (s_e - du_e);
}

// hlle_hyd.hpp:26-129 (ideal gas)
__device__ __forceinline__ void hyd_hlle(const State7 &l, const State7 &r, const EosParams &eos, Flux7 &f, bool *oor = nullptr) {
  const double gm1 = eos.gamma - 1.0;
  const double igm1 = eos.igm1;
  const double pl = gm1 * l.e, pr = gm1 * r.e;
  double sqrtdl = dsqrt(l.d, oor);
  double sqrtdr = dsqrt(r.d, oor);
  double isdlpdr = drcp(sqrtdl + sqrtdr, oor);
  double roe_vx = (sqrtdl * l.vx + sqrtdr * r.vx) * isdlpdr;
  double roe_vy = (sqrtdl * l.vy + sqrtdr * r.vy) * isdlpdr;
  double roe_vz = (sqrtdl * l.vz + sqrtdr * r.vz) * isdlpdr;
  double el = pl * igm1 + 0.5 * l.d * (sqr(l.vx) + sqr(l.vy) + sqr(l.vz));
  double er = pr * igm1 + 0.5 * r.d * (sqr(r.vx) + sqr(r.vy) + sqr(r.vz));
  double hroe = (ddiv(el + pl, sqrtdl, oor) + ddiv(er + pr, sqrtdr, oor)) * isdlpdr;
  double qa = sound_speed(eos.gamma, l.d, pl, oor);
  double qb = sound_speed(eos.gamma, r.d, pr, oor);
  double a = hroe - 0.5 * (sqr(roe_vx) + sqr(roe_vy) + sqr(roe_vz));
  a = (a < 0.0) ? 0.0 : dsqrt(gm1 * a, oor);
  double al = dmin((roe_vx - a), (l.vx - qa));
  double ar = dmax((roe_vx + a), (r.vx + qb));
  double bp = (ar > 0.0) ? ar : 1.0e-20;
  double bm = (al < 0.0) ? al : -1.0e-20;
  qa = l.vx - bm;
  qb = r.vx - bp;
  double fl_d = l.d * qa, fr_d = r.d * qb;
  double fl_mx = l.d * l.vx * qa, fr_mx = r.d * r.vx * qb;
  double fl_my = l.d * l.vy * qa, fr_my = r.d * r.vy * qb;
  double fl_mz = l.d * l.vz * qa, fr_mz = r.d * r.vz * qb;
  fl_mx += pl;
  fr_mx += pr;
  double fl_e = el * qa + pl * l.vx;
  double fr_e = er * qb + pr * r.vx;
  qa = 0.0;
  if (bp != bm) qa = ddiv(0.5 * (bp + bm), bp - bm, oor);
  f.d = 0.5 * (fl_d + fr_d) + qa * (fl_d - fr_d);
  f.mx = 0.5 * (fl_mx + fr_mx) + qa * (fl_mx - fr_mx);
  f.my = 0.5 * (fl_my + fr_my) + qa * (fl_my - fr_my);
  f.mz = 0.5 * (fl_mz + fr_mz) + qa * (fl_mz - fr_mz);
  f.e = 0.5 * (fl_e + fr_e) + qa * (fl_e - fr_e);
}

// hllc_hyd.hpp:20-115 (ideal gas only, as in the reference)
__device__ __forceinline__ void hyd_hllc(const State7 &l, const State7 &r, const EosParams &eos, Flux7 &f, bool *oor = nullptr) {
  const double gm1 = eos.gamma - 1.0;
  const double igm1 = eos.igm1;
  const double alpha = ddiv((eos.gamma) + 1.0, 2.0 * (eos.gamma), oor);
  const double pl = gm1 * l.e, pr = gm1 * r.e;
  double qa, qb, qc, qd, qe, qf;
  qa = sound_speed(eos.gamma, l.d, pl, oor);
  qb = sound_speed(eos.gamma, r.d, pr, oor);
  double el = pl * igm1 + 0.5 * l.d * (sqr(l.vx) + sqr(l.vy) + sqr(l.vz));
  double er = pr * igm1 + 0.5 * r.d * (sqr(r.vx) + sqr(r.vy) + sqr(r.vz));
  qc = 0.25 * (l.d + r.d) * (qa + qb);
  qd = 0.5 * (pl + pr + (l.vx - r.vx) * qc);   // P_mid
  qe = (qd <= pl) ? 1.0 : dsqrt(1.0 + alpha * (ddiv(qd, pl, oor) - 1.0), oor);
  qf = (qd <= pr) ? 1.0 : dsqrt(1.0 + alpha * (ddiv(qd, pr, oor) - 1.0), oor);
  qc = l.vx - qa * qe;   // al
  qd = r.vx + qb * qf;   // ar
  qa = qd > 0.0 ? qd : 1.0e-20;    // bp
  qb = qc < 0.0 ? qc : -1.0e-20;   // bm
  qe = l.vx - qc;   // vxl
  qf = r.vx - qd;   // vxr
  qc = pl + qe * l.d * l.vx;   // tl
  qd = pr + qf * r.d * r.vx;   // tr
  double ml = l.d * qe;
  double mr = -(r.d * qf);
  double am = ddiv(qc - qd, ml + mr, oor);
  double cp = ddiv(ml * qd + mr * qc, ml + mr, oor);
  cp = cp > 0.0 ? cp : 0.0;
  qe = l.d * (l.vx - qb);
  qf = r.d * (r.vx - qa);
  double fl_d = qe, fr_d = qf;
  double fl_mx = qe * l.vx + pl, fr_mx = qf * r.vx + pr;
  double fl_my = qe * l.vy, fr_my = qf * r.vy;
  double fl_mz = qe * l.vz, fr_mz = qf * r.vz;
  double fl_e = el * (l.vx - qb) + pl * l.vx;
  double fr_e = er * (r.vx - qa) + pr * r.vx;
  if (am >= 0.0) {
    qc = ddiv(am, am - qb, oor);
    qd = 0.0;
    qe = ddiv(-qb, am - qb, oor);
  } else {
    qc = 0.0;
    qd = ddiv(-am, qa - am, oor);
    qe = ddiv(qa, qa - am, oor);
  }
  f.d = qc * fl_d + qd * fr_d;
  f.mx = qc * fl_mx + qd * fr_mx + qe * cp;
  f.my = qc * fl_my + qd * fr_my;
  f.mz = qc * fl_mz + qd * fr_mz;
  f.e = qc * fl_e + qd * fr_e + qe * cp * am;
}

// roe_hyd.hpp:38-176 + RoeFluxAdb :187-283 (ideal gas), with the LLF fallback of Step 6
__device__ __forceinline__ void hyd_roe(const State7 &l, const State7 &r, const EosParams &eos, Flux7 &f, bool *oor = nullptr) {
  const double gm1 = eos.gamma - 1.0;
  const double pl = gm1 * l.e, pr = gm1 * r.e;
  double sqrtdl = dsqrt(l.d, oor);
  double sqrtdr = dsqrt(r.d, oor);
  double isdlpdr = drcp(sqrtdl + sqrtdr, oor);
  const double v1 = (sqrtdl * l.vx + sqrtdr * r.vx) * isdlpdr;
  const double v2 = (sqrtdl * l.vy + sqrtdr * r.vy) * isdlpdr;
  const double v3 = (sqrtdl * l.vz + sqrtdr * r.vz) * isdlpdr;
  double el = ddiv(pl, gm1, oor) + 0.5 * l.d * (sqr(l.vx) + sqr(l.vy) + sqr(l.vz));
  double er = ddiv(pr, gm1, oor) + 0.5 * r.d * (sqr(r.vx) + sqr(r.vy) + sqr(r.vz));
  const double h = (ddiv(el + pl, sqrtdl, oor) + ddiv(er + pr, sqrtdr, oor)) * isdlpdr;
  double mxl = l.d * l.vx;
  double mxr = r.d * r.vx;
  double fl[5], fr[5], flx[5], du[5];
  fl[0] = mxl;              fr[0] = mxr;
  fl[1] = mxl * l.vx;       fr[1] = mxr * r.vx;
  fl[2] = mxl * l.vy;       fr[2] = mxr * r.vy;
  fl[3] = mxl * l.vz;       fr[3] = mxr * r.vz;
  fl[1] += pl;              fr[1] += pr;
  fl[4] = (el + pl) * l.vx; fr[4] = (er + pr) * r.vx;
  du[0] = r.d - l.d;
  du[1] = r.d * r.vx - l.d * l.vx;
  du[2] = r.d * r.vy - l.d * l.vy;
  du[3] = r.d * r.vz - l.d * l.vz;
  du[4] = er - el;
#pragma unroll
  for (int n = 0; n < 5; ++n) flx[n] = 0.5 * (fl[n] + fr[n]);
  // RoeFluxAdb
  int llf_flag = 0;
  double ev0, ev4;
  {
    double vsq = v1 * v1 + v2 * v2 + v3 * v3;
    double q = h - 0.5 * vsq;
    double cs_sq = (q < 0.0) ? (double)1.17549435e-38f : gm1 * q;   // FLT_MIN
    double cs = dsqrt(cs_sq, oor);
    ev0 = v1 - cs;
    ev4 = v1 + cs;
    double a[5];
    double na = ddiv(0.5, cs_sq, oor);
    a[0] = du[0] * (0.5 * gm1 * vsq + v1 * cs);
    a[0] -= du[1] * (gm1 * v1 + cs);
    a[0] -= du[2] * gm1 * v2;
    a[0] -= du[3] * gm1 * v3;
    a[0] += du[4] * gm1;
    a[0] *= na;
    a[1] = du[0] * (-v2);
    a[1] += du[2];
    a[2] = du[0] * (-v3);
    a[2] += du[3];
    double qa = ddiv(gm1, cs_sq, oor);
    a[3] = du[0] * (1.0 - na * gm1 * vsq);
    a[3] += du[1] * qa * v1;
    a[3] += du[2] * qa * v2;
    a[3] += du[3] * qa * v3;
    a[3] -= du[4] * qa;
    a[4] = du[0] * (0.5 * gm1 * vsq - v1 * cs);
    a[4] -= du[1] * (gm1 * v1 - cs);
    a[4] -= du[2] * gm1 * v2;
    a[4] -= du[3] * gm1 * v3;
    a[4] += du[4] * gm1;
    a[4] *= na;
    double c0 = -0.5 * fabs(ev0) * a[0];
    double c1 = -0.5 * fabs(v1) * a[1];
    double c2 = -0.5 * fabs(v1) * a[2];
    double c3 = -0.5 * fabs(v1) * a[3];
    double c4 = -0.5 * fabs(ev4) * a[4];
    double dens = l.d + a[0];
    if (dens < 0.0) llf_flag = 1;
    dens += a[3];
    if (dens < 0.0) llf_flag = 1;
    flx[0] += c0;
    flx[0] += c3;
    flx[0] += c4;
    flx[1] += c0 * (v1 - cs);
    flx[1] += c3 * v1;
    flx[1] += c4 * (v1 + cs);
    flx[2] += c0 * v2;
    flx[2] += c1;
    flx[2] += c3 * v2;
    flx[2] += c4 * v2;
    flx[3] += c0 * v3;
    flx[3] += c2;
    flx[3] += c3 * v3;
    flx[3] += c4 * v3;
    flx[4] += c0 * (h - v1 * cs);
    flx[4] += c1 * v2;
    flx[4] += c2 * v3;
    flx[4] += c3 * 0.5 * vsq;
    flx[4] += c4 * (h + v1 * cs);
  }
  if (ev0 >= 0.0) {
#pragma unroll
    for (int n = 0; n < 5; ++n) flx[n] = fl[n];
  }
  if (ev4 <= 0.0) {
#pragma unroll
    for (int n = 0; n < 5; ++n) flx[n] = fr[n];
  }
  if (llf_flag != 0) {
    double cl = sound_speed(eos.gamma, l.d, pl, oor);
    double cr = sound_speed(eos.gamma, r.d, pr, oor);
    double a = 0.5 * dmax((fabs(l.vx) + cl), (fabs(r.vx) + cr));
#pragma unroll
    for (int n = 0; n < 5; ++n) flx[n] = 0.5 * (fl[n] + fr[n]) - a * du[n];
  }
  f.d = flx[0]; f.mx = flx[1]; f.my = flx[2]; f.mz = flx[3]; f.e = flx[4];
}

template <int RS>
__device__ __forceinline__ void riemann_hyd(const State7 &l, const State7 &r, const EosParams &eos, Flux7 &f, bool *oor = nullptr) {
  if constexpr (RS == RS_HYD_ADVECT) hyd_advect(l, r, eos, f, oor);
  else if constexpr (RS == RS_HYD_LLF) hyd_llf(l, r, eos, f, oor);
  else if constexpr (RS == RS_HYD_HLLE) hyd_hlle(l, r, eos, f, oor);
  else if constexpr (RS == RS_HYD_HLLC) hyd_hllc(l, r, eos, f, oor);
  else hyd_roe(l, r, eos, f, oor);
}

}  // namespace akb
