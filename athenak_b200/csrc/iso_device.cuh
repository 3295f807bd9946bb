// iso_device.cuh -- isothermal-EOS Riemann solvers (reference
// src/mhd/rsolvers/llf_mhd_singlestate.hpp:27-89, hlle_mhd.hpp:24-179, hlld_mhd.hpp:350-546 and
// src/hydro/rsolvers/llf_hyd_singlestate.hpp:27-82, hlle_hyd.hpp:26-129, roe_hyd.hpp:38-176,
// 290-345: the `else` branches of `eos.is_ideal`).  Four fluid variables (no energy); State7.e
// and Flux7.e are unused.  Same conventions as mhd_device.cuh.
#pragma once
#include "mhd_device.cuh"

namespace akb {

// internal solver codes: public code (AK_RSOLVER_*) + 20 for the isothermal variant
enum RSolverIso { RS_ISO = 20, RS_LLF_ISO = 20, RS_HLLE_ISO = 21, RS_HLLD_ISO = 22,
                  RS_HYD_LLF_ISO = 31, RS_HYD_HLLE_ISO = 32, RS_HYD_ROE_ISO = 34 };
__host__ __device__ constexpr bool rs_is_iso(int rs) { return rs >= RS_ISO; }
__host__ __device__ constexpr bool rs_is_hydro(int rs) { return (rs % 20) >= 10; }

// eos.hpp:60-68
__device__ __forceinline__ double fast_speed_iso(double iso_cs, double d, double bx, double by, double bz, bool *oor = nullptr) {
  double asq = (iso_cs * iso_cs) * d;
  double ct2 = by * by + bz * bz;
  double qsq = bx * bx + ct2 + asq;
  double tmp = bx * bx + ct2 - asq;
  return dsqrt(ddiv(0.5 * (qsq + dsqrt(tmp * tmp + 4.0 * asq * ct2)), d, oor), oor);
}

__device__ __forceinline__ void llf_iso(const State7 &l, const State7 &r, double bxi, const EosParams &eos, Flux7 &f, bool *oor = nullptr) {
  double qa = l.d * l.vx;
  double qb = r.d * r.vx;
  double qc = 0.5 * (sqr(l.by) + sqr(l.bz) - sqr(bxi));
  double qd = 0.5 * (sqr(r.by) + sqr(r.bz) - sqr(bxi));
  double s_d = qa + qb;
  double s_mx = qa * l.vx + qb * r.vx + qc + qd;
  double s_my = qa * l.vy + qb * r.vy - bxi * (l.by + r.by);
  double s_mz = qa * l.vz + qb * r.vz - bxi * (l.bz + r.bz);
  double s_by = l.by * l.vx + r.by * r.vx - bxi * (l.vy + r.vy);
  double s_bz = l.bz * l.vx + r.bz * r.vx - bxi * (l.vz + r.vz);
  s_mx += sqr(eos.iso_cs) * (l.d + r.d);
  qa = fast_speed_iso(eos.iso_cs, l.d, bxi, l.by, l.bz, oor);
  qb = fast_speed_iso(eos.iso_cs, r.d, bxi, r.by, r.bz, oor);
  double a = dmax((fabs(l.vx) + qa), (fabs(r.vx) + qb));
  double du_d = a * (r.d - l.d);
  double du_mx = a * (r.d * r.vx - l.d * l.vx);
  double du_my = a * (r.d * r.vy - l.d * l.vy);
  double du_mz = a * (r.d * r.vz - l.d * l.vz);
  double du_by = a * (r.by - l.by);
  double du_bz = a * (r.bz - l.bz);
  f.d = 0.5 * (s_d - du_d);
  f.mx = 0.5 * (s_mx - du_mx);
  f.my = 0.5 * (s_my - du_my);
  f.mz = 0.5 * (s_mz - du_mz);
  f.ey = -0.5 * (s_by - du_by);
  f.ez = 0.5 * (s_bz - du_bz);
}

__device__ __forceinline__ void hlle_iso(const State7 &l, const State7 &r, double bxi, const EosParams &eos, Flux7 &f, bool *oor = nullptr) {
  const double iso_cs = eos.iso_cs;
  double sqrtdl = dsqrt(l.d, oor);
  double sqrtdr = dsqrt(r.d, oor);
  double isdlpdr = drcp(sqrtdl + sqrtdr, oor);
  double roe_d = sqrtdl * sqrtdr;
  double roe_vx = (sqrtdl * l.vx + sqrtdr * r.vx) * isdlpdr;
  double roe_by = (sqrtdr * l.by + sqrtdl * r.by) * isdlpdr;
  double roe_bz = (sqrtdr * l.bz + sqrtdl * r.bz) * isdlpdr;
  double x = ddiv(0.5 * (sqr(l.by - r.by) + sqr(l.bz - r.bz)), sqr(sqrtdl + sqrtdr), oor);
  double y = ddiv(0.5 * (l.d + r.d), roe_d, oor);
  double pbl = 0.5 * (bxi * bxi + sqr(l.by) + sqr(l.bz));
  double pbr = 0.5 * (bxi * bxi + sqr(r.by) + sqr(r.bz));
  double cl = fast_speed_iso(iso_cs, l.d, bxi, l.by, l.bz, oor);
  double cr = fast_speed_iso(iso_cs, r.d, bxi, r.by, r.bz, oor);
  double btsq = sqr(roe_by) + sqr(roe_bz);
  double vaxsq = ddiv(bxi * bxi, roe_d, oor);
  double bt_starsq = btsq * y;
  double twid_asq = iso_cs * iso_cs + x;
  double ct2 = ddiv(bt_starsq, roe_d, oor);
  double tsum = vaxsq + ct2 + twid_asq;
  double tdif = vaxsq + ct2 - twid_asq;
  double cf2_cs2 = dsqrt(tdif * tdif + 4.0 * twid_asq * ct2, oor);
  double cfsq = 0.5 * (tsum + cf2_cs2);
  double a = dsqrt(cfsq, oor);
  double al = dmin((roe_vx - a), (l.vx - cl));
  double ar = dmax((roe_vx + a), (r.vx + cr));
  double bp = ar > 0.0 ? ar : 1.0e-20;
  double bm = al < 0.0 ? al : -1.0e-20;
  double vxl = l.vx - bm;
  double vxr = r.vx - bp;
  double fl_d = l.d * vxl, fr_d = r.d * vxr;
  double fl_mx = l.d * l.vx * vxl + pbl - sqr(bxi);
  double fr_mx = r.d * r.vx * vxr + pbr - sqr(bxi);
  double fl_my = l.d * l.vy * vxl - bxi * l.by;
  double fr_my = r.d * r.vy * vxr - bxi * r.by;
  double fl_mz = l.d * l.vz * vxl - bxi * l.bz;
  double fr_mz = r.d * r.vz * vxr - bxi * r.bz;
  fl_mx += (iso_cs * iso_cs) * l.d;
  fr_mx += (iso_cs * iso_cs) * r.d;
  double fl_by = l.by * vxl - bxi * l.vy;
  double fr_by = r.by * vxr - bxi * r.vy;
  double fl_bz = l.bz * vxl - bxi * l.vz;
  double fr_bz = r.bz * vxr - bxi * r.vz;
  double tmp = 0.0;
  if (bp != bm) tmp = ddiv(0.5 * (bp + bm), bp - bm, oor);
  f.d = 0.5 * (fl_d + fr_d) + (fl_d - fr_d) * tmp;
  f.mx = 0.5 * (fl_mx + fr_mx) + (fl_mx - fr_mx) * tmp;
  f.my = 0.5 * (fl_my + fr_my) + (fl_my - fr_my) * tmp;
  f.mz = 0.5 * (fl_mz + fr_mz) + (fl_mz - fr_mz) * tmp;
  f.ey = -0.5 * (fl_by + fr_by) - (fl_by - fr_by) * tmp;
  f.ez = 0.5 * (fl_bz + fr_bz) + (fl_bz - fr_bz) * tmp;
}

// hlld_mhd.hpp:350-546 (Mignone 2007)
__device__ __forceinline__ void hlld_iso(const State7 &l, const State7 &r, double bxi, const EosParams &eos, Flux7 &f, bool *oor = nullptr) {
  const double small = 1.0e-4;   // HLLD_SMALL_NUMBER
  const double iso_cs = eos.iso_cs;
  Cons7 ul, ur;
  ul.d = l.d; ul.mx = l.vx * ul.d; ul.my = l.vy * ul.d; ul.mz = l.vz * ul.d; ul.by = l.by; ul.bz = l.bz;
  ur.d = r.d; ur.mx = r.vx * ur.d; ur.my = r.vy * ur.d; ur.mz = r.vz * ur.d; ur.by = r.by; ur.bz = r.bz;
  const double cfl = fast_speed_iso(iso_cs, l.d, bxi, l.by, l.bz, oor);
  const double cfr = fast_speed_iso(iso_cs, r.d, bxi, r.by, r.bz, oor);
  const double s0 = dmin(l.vx - cfl, r.vx - cfr);
  const double s4 = dmax(l.vx + cfl, r.vx + cfr);
  const double bxsq = bxi * bxi;
  const double ptl = sqr(iso_cs) * l.d + 0.5 * (bxsq + sqr(l.by) + sqr(l.bz));
  const double ptr = sqr(iso_cs) * r.d + 0.5 * (bxsq + sqr(r.by) + sqr(r.bz));
  Cons7 fl, fr;
  fl.d = ul.mx;
  fl.mx = ul.mx * l.vx + ptl - bxsq;
  fl.my = ul.my * l.vx - bxi * ul.by;
  fl.mz = ul.mz * l.vx - bxi * ul.bz;
  fl.by = ul.by * l.vx - bxi * l.vy;
  fl.bz = ul.bz * l.vx - bxi * l.vz;
  fr.d = ur.mx;
  fr.mx = ur.mx * r.vx + ptr - bxsq;
  fr.my = ur.my * r.vx - bxi * ur.by;
  fr.mz = ur.mz * r.vx - bxi * ur.bz;
  fr.by = ur.by * r.vx - bxi * r.vy;
  fr.bz = ur.bz * r.vx - bxi * r.vz;
  if (s0 >= 0.0) { f.d = fl.d; f.mx = fl.mx; f.my = fl.my; f.mz = fl.mz; f.ey = -fl.by; f.ez = fl.bz; return; }
  if (s4 <= 0.0) { f.d = fr.d; f.mx = fr.mx; f.my = fr.my; f.mz = fr.mz; f.ey = -fr.by; f.ez = fr.bz; return; }
  const double idspd = drcp(s4 - s0, oor);
  double dhll = (s4 * ur.d - s0 * ul.d - fr.d + fl.d) * idspd;
  dhll = dmax(dhll, eos.dfloor);
  const double sqrtdhll = dsqrt(dhll, oor);
  const double fdhll = (s4 * fl.d - s0 * fr.d + s4 * s0 * (ur.d - ul.d)) * idspd;
  const double fmxhll = (s4 * fl.mx - s0 * fr.mx + s4 * s0 * (ur.mx - ul.mx)) * idspd;
  const double ustar = ddiv(fdhll, dhll, oor);
  const double mxhll = (s4 * ur.mx - s0 * ul.mx - fr.mx + fl.mx) * idspd;
  const double s1 = ustar - ddiv(fabs(bxi), sqrtdhll, oor);
  const double s3 = ustar + ddiv(fabs(bxi), sqrtdhll, oor);
  Cons7 ulst, urst;
  ulst.d = dhll; ulst.mx = mxhll;
  double tmp = (s0 - s1) * (s0 - s3);
  if (fabs(s0 - s1) < (small)*iso_cs) {
    ulst.my = ul.my; ulst.mz = ul.mz; ulst.by = ul.by; ulst.bz = ul.bz;
  } else {
    const double mfact = ddiv(bxi * (ustar - l.vx), tmp, oor);
    const double bfact = ddiv(ul.d * sqr(s0 - l.vx) - bxsq, dhll * tmp, oor);
    ulst.my = dhll * l.vy - ul.by * mfact;
    ulst.mz = dhll * l.vz - ul.bz * mfact;
    ulst.by = ul.by * bfact;
    ulst.bz = ul.bz * bfact;
  }
  urst.d = dhll; urst.mx = mxhll;
  tmp = (s4 - s1) * (s4 - s3);
  if (fabs(s4 - s3) < (small)*iso_cs) {
    urst.my = ur.my; urst.mz = ur.mz; urst.by = ur.by; urst.bz = ur.bz;
  } else {
    const double mfact = ddiv(bxi * (ustar - r.vx), tmp, oor);
    const double bfact = ddiv(ur.d * sqr(s4 - r.vx) - bxsq, dhll * tmp, oor);
    urst.my = dhll * r.vy - ur.by * mfact;
    urst.mz = dhll * r.vz - ur.bz * mfact;
    urst.by = ur.by * bfact;
    urst.bz = ur.bz * bfact;
  }
  if (s1 >= 0.0) {
    f.d = fl.d + s0 * (ulst.d - ul.d);
    f.mx = fl.mx + s0 * (ulst.mx - ul.mx);
    f.my = fl.my + s0 * (ulst.my - ul.my);
    f.mz = fl.mz + s0 * (ulst.mz - ul.mz);
    f.ey = -(fl.by + s0 * (ulst.by - ul.by));
    f.ez = fl.bz + s0 * (ulst.bz - ul.bz);
  } else if (s3 <= 0.0) {
    f.d = fr.d + s4 * (urst.d - ur.d);
    f.mx = fr.mx + s4 * (urst.mx - ur.mx);
    f.my = fr.my + s4 * (urst.my - ur.my);
    f.mz = fr.mz + s4 * (urst.mz - ur.mz);
    f.ey = -(fr.by + s4 * (urst.by - ur.by));
    f.ez = fr.bz + s4 * (urst.bz - ur.bz);
  } else {
    const double x = sqrtdhll * (bxi > 0.0 ? 1.0 : -1.0);
    const double c_my = 0.5 * (ulst.my + urst.my + (urst.by - ulst.by) * x);
    const double c_mz = 0.5 * (ulst.mz + urst.mz + (urst.bz - ulst.bz) * x);
    const double c_by = 0.5 * (ulst.by + urst.by + ddiv(urst.my - ulst.my, x, oor));
    const double c_bz = 0.5 * (ulst.bz + urst.bz + ddiv(urst.mz - ulst.mz, x, oor));
    f.d = dhll * ustar;
    f.mx = fmxhll;
    f.my = c_my * ustar - bxi * c_by;
    f.mz = c_mz * ustar - bxi * c_bz;
    f.ey = -(c_by * ustar - ddiv(bxi * c_my, dhll, oor));
    f.ez = c_bz * ustar - ddiv(bxi * c_mz, dhll, oor);
  }
}

// ---- hydro -------------------------------------------------------------------------------
__device__ __forceinline__ void hyd_llf_iso(const State7 &l, const State7 &r, const EosParams &eos, Flux7 &f, bool *oor = nullptr) {
  double qa = l.d * l.vx;
  double qb = r.d * r.vx;
  double s_d = qa + qb;
  double s_mx = qa * l.vx + qb * r.vx;
  double s_my = qa * l.vy + qb * r.vy;
  double s_mz = qa * l.vz + qb * r.vz;
  s_mx += sqr(eos.iso_cs) * (l.d + r.d);
  double a = dmax((fabs(l.vx) + eos.iso_cs), (fabs(r.vx) + eos.iso_cs));
  double du_d = a * (r.d - l.d);
  double du_mx = a * (r.d * r.vx - l.d * l.vx);
  double du_my = a * (r.d * r.vy - l.d * l.vy);
  double du_mz = a * (r.d * r.vz - l.d * l.vz);
  f.d = 0.5 * (s_d - du_d);
  f.mx = 0.5 * (s_mx - du_mx);
  f.my = 0.5 * (s_my - du_my);
  f.mz = 0.5 * (s_mz - du_mz);
}

__device__ __forceinline__ void hyd_hlle_iso(const State7 &l, const State7 &r, const EosParams &eos, Flux7 &f, bool *oor = nullptr) {
  const double iso_cs = eos.iso_cs;
  double sqrtdl = dsqrt(l.d, oor);
  double sqrtdr = dsqrt(r.d, oor);
  double isdlpdr = drcp(sqrtdl + sqrtdr, oor);
  double roe_vx = (sqrtdl * l.vx + sqrtdr * r.vx) * isdlpdr;
  double al = dmin((roe_vx - iso_cs), (l.vx - iso_cs));
  double ar = dmax((roe_vx + iso_cs), (r.vx + iso_cs));
  double bp = (ar > 0.0) ? ar : 1.0e-20;
  double bm = (al < 0.0) ? al : -1.0e-20;
  double qa = l.vx - bm;
  double qb = r.vx - bp;
  double fl_d = l.d * qa, fr_d = r.d * qb;
  double fl_mx = l.d * l.vx * qa, fr_mx = r.d * r.vx * qb;
  double fl_my = l.d * l.vy * qa, fr_my = r.d * r.vy * qb;
  double fl_mz = l.d * l.vz * qa, fr_mz = r.d * r.vz * qb;
  fl_mx += (iso_cs * iso_cs) * l.d;
  fr_mx += (iso_cs * iso_cs) * r.d;
  qa = 0.0;
  if (bp != bm) qa = ddiv(0.5 * (bp + bm), bp - bm, oor);
  f.d = 0.5 * (fl_d + fr_d) + qa * (fl_d - fr_d);
  f.mx = 0.5 * (fl_mx + fr_mx) + qa * (fl_mx - fr_mx);
  f.my = 0.5 * (fl_my + fr_my) + qa * (fl_my - fr_my);
  f.mz = 0.5 * (fl_mz + fr_mz) + qa * (fl_mz - fr_mz);
}

__device__ __forceinline__ void hyd_roe_iso(const State7 &l, const State7 &r, const EosParams &eos, Flux7 &f, bool *oor = nullptr) {
  const double iso_cs = eos.iso_cs;
  double sqrtdl = dsqrt(l.d, oor);
  double sqrtdr = dsqrt(r.d, oor);
  double isdlpdr = drcp(sqrtdl + sqrtdr, oor);
  const double v1 = (sqrtdl * l.vx + sqrtdr * r.vx) * isdlpdr;
  const double v2 = (sqrtdl * l.vy + sqrtdr * r.vy) * isdlpdr;
  const double v3 = (sqrtdl * l.vz + sqrtdr * r.vz) * isdlpdr;
  double mxl = l.d * l.vx;
  double mxr = r.d * r.vx;
  double fl[4], fr[4], flx[4], du[4];
  fl[0] = mxl;        fr[0] = mxr;
  fl[1] = mxl * l.vx; fr[1] = mxr * r.vx;
  fl[2] = mxl * l.vy; fr[2] = mxr * r.vy;
  fl[3] = mxl * l.vz; fr[3] = mxr * r.vz;
  fl[1] += (iso_cs * iso_cs) * l.d;
  fr[1] += (iso_cs * iso_cs) * r.d;
  du[0] = r.d - l.d;
  du[1] = r.d * r.vx - l.d * l.vx;
  du[2] = r.d * r.vy - l.d * l.vy;
  du[3] = r.d * r.vz - l.d * l.vz;
#pragma unroll
  for (int n = 0; n < 4; ++n) flx[n] = 0.5 * (fl[n] + fr[n]);
  int llf_flag = 0;
  const double ev0 = v1 - iso_cs, ev3 = v1 + iso_cs;
  {
    double a0 = du[0] * (0.5 + ddiv(0.5 * v1, iso_cs, oor));
    a0 -= ddiv(du[1] * 0.5, iso_cs, oor);
    double a1 = du[0] * (-v2);
    a1 += du[2];
    double a2 = du[0] * (-v3);
    a2 += du[3];
    double a3 = du[0] * (0.5 - ddiv(0.5 * v1, iso_cs, oor));
    a3 += ddiv(du[1] * 0.5, iso_cs, oor);
    const double c0 = -0.5 * fabs(ev0) * a0;
    const double c1 = -0.5 * fabs(v1) * a1;
    const double c2 = -0.5 * fabs(v1) * a2;
    const double c3 = -0.5 * fabs(ev3) * a3;
    double dens = l.d + a0;
    if (dens < 0.0) llf_flag = 1;
    dens += a3;
    if (dens < 0.0) llf_flag = 1;
    flx[0] += c0;
    flx[0] += c3;
    flx[1] += c0 * (v1 - iso_cs);
    flx[1] += c3 * (v1 + iso_cs);
    flx[2] += c0 * v2;
    flx[2] += c1;
    flx[2] += c3 * v2;
    flx[3] += c0 * v3;
    flx[3] += c2;
    flx[3] += c3 * v3;
  }
  if (ev0 >= 0.0) {
#pragma unroll
    for (int n = 0; n < 4; ++n) flx[n] = fl[n];
  }
  if (ev3 <= 0.0) {
#pragma unroll
    for (int n = 0; n < 4; ++n) flx[n] = fr[n];
  }
  if (llf_flag != 0) {
    double a = 0.5 * dmax((fabs(l.vx) + iso_cs), (fabs(r.vx) + iso_cs));
#pragma unroll
    for (int n = 0; n < 4; ++n) flx[n] = 0.5 * (fl[n] + fr[n]) - a * du[n];
  }
  f.d = flx[0]; f.mx = flx[1]; f.my = flx[2]; f.mz = flx[3];
}

}  // namespace akb
