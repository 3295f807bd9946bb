// c2p_device.cuh -- one cell of ConsToPrim and of the new-dt reduction for ideal / isothermal MHD, shared by cons2prim_kernel
// (mhd_kernels.cu) and the fused stage kernels (stage_tail.cu).
#pragma once
#include "mhd_device.cuh"
#include "iso_device.cuh"

#ifndef AKB_C2P_SCALED
#define AKB_C2P_SCALED 1      // ConsToPrim divisions through ddiv_x (exact for every numerator, no repeat): 0.52 -> 0.42 ms at 256^3
#endif

namespace akb {

__device__ __forceinline__ double c2p_div(double a, double b, bool *oor) {
#if AKB_C2P_SCALED
  return oor ? ddiv_x(a, b, oor) : ddiv(a, b);
#else
  return ddiv(a, b, oor);
#endif
}

// One cell of ConsToPrim (ideal_mhd.cpp:47-122 + ideal_c2p_mhd.hpp:20-67; isothermal_mhd.cpp:28-46,95-135), shared by the
// fused tail and the shell kernel.  u[]: conserved (floors applied in place, flags say which), b: cell-centred field,
// w[]: primitives out.  oor: deferred range guard of the division / square root (mhd_device.cuh).
template <bool IDEAL>
__device__ __forceinline__ void c2p_cell(const EosParams &eos, double &ud, double umx, double umy, double umz, double &ue,
                                         double ubx, double uby, double ubz, double &wd, double &wvx, double &wvy, double &wvz,
                                         double &we, bool &dfl, bool &efl, bool &tfl, bool *oor) {
  const double gm1 = eos.gamma - 1.0;
  const double b2 = sqr(ubx) + sqr(uby) + sqr(ubz);
  const double dfloor_ = dmax(eos.dfloor, c2p_div(b2, eos.sigma_max, oor));
  if (ud < dfloor_) { ud = dfloor_; dfl = true; }
  wd = ud;
  const double di = drcp(ud, oor);
  wvx = di * umx; wvy = di * umy; wvz = di * umz;
  we = 0.0;
  if constexpr (IDEAL) {
    const double efloor = c2p_div(eos.pfloor, eos.gamma - 1.0, oor);
    const double e_k = 0.5 * di * (sqr(umx) + sqr(umy) + sqr(umz));
    const double e_m = 0.5 * (sqr(ubx) + sqr(uby) + sqr(ubz));
    we = (ue - e_k - e_m);
    if (we < efloor) { we = efloor; ue = efloor + e_k + e_m; efl = true; }
    if (gm1 * we * di < eos.tfloor) { we = wd * eos.tfloor / gm1; ue = we + e_k + e_m; tfl = true; }
    // entropy floor: pow() only when the floor can fire (see cons2prim_kernel, mhd_kernels.cu)
    const double pow_ub = dmax(1.0, wd * wd);
    const bool need_pow = !(gm1 > 0.0 && gm1 <= 2.0 && (gm1 * we * di) > 4.0 * eos.sfloor * pow_ub);
    if (need_pow) {
      const double spe_over_eps = gm1 / pow(wd, gm1);
      const double spe = spe_over_eps * we * di;
      if (spe <= eos.sfloor) { we = wd * eos.sfloor / spe_over_eps; efl = true; }
    }
  }
}

// new-dt contribution of one active cell (mhd_newdt.cpp:128-138,146-148)
template <bool IDEAL>
__device__ __forceinline__ void newdt_cell(const EosParams &eos, double dx1, double dx2, double dx3, double wd, double wvx, double wvy,
                                           double wvz, double we, double bx, double by, double bz, double &d1, double &d2, double &d3,
                                           bool *oor) {
  const double p = (eos.gamma - 1.0) * we;
  auto cf = [&](double b1, double b2, double b3) {
    if constexpr (IDEAL) return fast_speed(eos.gamma, wd, p, b1, b2, b3, oor);
    else return fast_speed_iso(eos.iso_cs, wd, b1, b2, b3, oor);
  };
  d1 = c2p_div(dx1, fabs(wvx) + cf(bx, by, bz), oor);
  d2 = c2p_div(dx2, fabs(wvy) + cf(by, bz, bx), oor);
  d3 = c2p_div(dx3, fabs(wvz) + cf(bz, bx, by), oor);
}

}  // namespace akb
