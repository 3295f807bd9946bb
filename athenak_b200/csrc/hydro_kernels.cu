// hydro_kernels.cu -- the Newtonian ideal-gas HYDRO twin of the stage update (SURVEY.md 8f item 1;
// reference src/hydro/hydro_fluxes.cpp, hydro_update.cpp, hydro_newdt.cpp, src/eos/ideal_hyd.cpp).
// The sweep kernels are the MHD ones (flux_kernels.cuh) instantiated with the hydro Riemann
// solvers: five variables, no face field, no EMF outputs, no transverse extension of the sweep
// ranges; 1-D, 2-D and 3-D MeshBlocks.  The RK update is update_ct_kernel<U> (mhd_kernels.cu).
#include "flux_kernels.cuh"
#include "reduce.cuh"

namespace akb {

// ConsToPrim (ideal_hyd.cpp:28-113 + ideal_c2p_hyd.hpp:20-67) over a cell box, optionally fused
// with the new-dt reduction over the active cells (hydro_newdt.cpp:75-120, Newtonian branch).
template <bool NEWDT, bool IDEAL>
__global__ void __launch_bounds__(256)
hydro_cons2prim_kernel(Geom g, EosParams eos, double *__restrict__ u0, double *__restrict__ w0, int il, int iu,
                       int jl, int ju, int kl, int ku, int *__restrict__ counters,
                       const double *__restrict__ mb_size, unsigned long long *__restrict__ dtmin) {
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ni = iu - il + 1, nj = ju - jl + 1, nk = ku - kl + 1;
  const long total = (long)g.nmb * nk * nj * ni;
  int nd = 0, ne = 0, nt = 0;
  double d1 = DT_IDENT, d2 = DT_IDENT, d3 = DT_IDENT;
  if (gt < total) {
    long t = gt;
    const int i = il + (int)(t % ni); t /= ni;
    const int j = jl + (int)(t % nj); t /= nj;
    const int k = kl + (int)(t % nk);
    const int m = (int)(t / nk);
    const long cc = g.cc;
    const long o = ((long)k * g.n2 + j) * g.n1 + i;
    double *u = u0 + (long)m * g.nvar * cc + o;
    double ud = u[IDN * cc], ue = 0.0;
    if constexpr (IDEAL) ue = u[IEN * cc];
    const double umx = u[IM1 * cc], umy = u[IM2 * cc], umz = u[IM3 * cc];
    bool dfl = false, efl = false, tfl = false;
    const double gm1 = eos.gamma - 1.0;
    if (ud < eos.dfloor) { ud = eos.dfloor; dfl = true; }
    const double wd = ud;
    const double di = drcp(ud);
    const double wvx = di * umx, wvy = di * umy, wvz = di * umz;
    double we = 0.0;
    double *w = w0 + (long)m * g.nvar * cc + o;
    if constexpr (IDEAL) {
    const double efloor = ddiv(eos.pfloor, eos.gamma - 1.0);
    const double e_k = 0.5 * di * (sqr(umx) + sqr(umy) + sqr(umz));
    we = (ue - e_k);
    if (we < efloor) { we = efloor; ue = efloor + e_k; efl = true; }
    if (gm1 * we * di < eos.tfloor) { we = wd * eos.tfloor / gm1; ue = we + e_k; tfl = true; }
    // entropy floor: pow() only when the floor can fire (same argument as cons2prim_kernel)
    const double pow_ub = dmax(1.0, wd * wd);
    const bool need_pow = !(gm1 > 0.0 && gm1 <= 2.0 && (gm1 * we * di) > 4.0 * eos.sfloor * pow_ub);
    if (need_pow) {
      const double spe_over_eps = gm1 / pow(wd, gm1);
      const double spe = spe_over_eps * we * di;
      if (spe <= eos.sfloor) { we = wd * eos.sfloor / spe_over_eps; efl = true; }
    }
    if (efl) { u[IEN * cc] = ue; ne = 1; }
    if (tfl) { u[IEN * cc] = ue; nt = 1; }
    w[IEN * cc] = we;
    }  // IDEAL; isothermal_hyd.cpp:25-45,60-130 has the density floor only
    if (dfl) { u[IDN * cc] = ud; nd = 1; }
    w[IDN * cc] = wd; w[IVX * cc] = wvx; w[IVY * cc] = wvy; w[IVZ * cc] = wvz;
    for (int n = g.nfl; n < g.nvar; ++n) {   // ideal_hyd.cpp:95-101: scalar floor, then s / d (floored density)
      double sc = u[n * cc];
      if (IDEAL && sc < 0.0) { sc = 0.0; u[n * cc] = 0.0; }   // isothermal_hyd.cpp:116-118: no scalar floor
      w[n * cc] = sc / ud;
    }
    if (NEWDT) {
      if (i >= g.is && i <= g.ie && j >= g.js && j <= g.je && k >= g.ks && k <= g.ke) {
        double cs = eos.iso_cs;                                   // hydro_newdt.cpp:107-112
        if constexpr (IDEAL) cs = sound_speed(eos.gamma, wd, gm1 * we);
        d1 = ddiv(__ldg(mb_size + m * 9 + 6), fabs(wvx) + cs);
        d2 = ddiv(__ldg(mb_size + m * 9 + 7), fabs(wvy) + cs);
        d3 = ddiv(__ldg(mb_size + m * 9 + 8), fabs(wvz) + cs);
      }
    }
  }
  warp_count3(nd, ne, nt, counters);
  if (NEWDT) block_min3(d1, d2, d3, dtmin);
}

// Hydro::NewTimeStep stand-alone (hydro_newdt.cpp:75-120): reads w0 over the active cells.
template <bool IDEAL>
__global__ void __launch_bounds__(256)
hydro_newdt_kernel(Geom g, EosParams eos, const double *__restrict__ w0, const double *__restrict__ mb_size,
                   unsigned long long *__restrict__ dtmin) {
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
    const double *w = w0 + (long)m * g.nvar * cc + o;
    double cs = eos.iso_cs;
    if constexpr (IDEAL) cs = sound_speed(eos.gamma, __ldg(w), (eos.gamma - 1.0) * __ldg(w + IEN * cc));
    d1 = ddiv(__ldg(mb_size + m * 9 + 6), fabs(__ldg(w + IVX * cc)) + cs);
    d2 = ddiv(__ldg(mb_size + m * 9 + 7), fabs(__ldg(w + IVY * cc)) + cs);
    d3 = ddiv(__ldg(mb_size + m * 9 + 8), fabs(__ldg(w + IVZ * cc)) + cs);
  }
  block_min3(d1, d2, d3, dtmin);
}

// ---------------------------------------------------------------------------------------
template <int RECON>
static cudaError_t launch_hydro_fluxes_r(int rs, const Geom &g, const EosParams &eos, const Arrays &a, cudaStream_t st) {
  switch (rs) {
    case RS_HYD_ADVECT: return launch_fluxes_t<RECON, RS_HYD_ADVECT>(g, eos, a, st);
    case RS_HYD_LLF: return launch_fluxes_t<RECON, RS_HYD_LLF>(g, eos, a, st);
    case RS_HYD_HLLE: return launch_fluxes_t<RECON, RS_HYD_HLLE>(g, eos, a, st);
    case RS_HYD_HLLC: return launch_fluxes_t<RECON, RS_HYD_HLLC>(g, eos, a, st);
    case RS_HYD_ROE: return launch_fluxes_t<RECON, RS_HYD_ROE>(g, eos, a, st);
    case RS_HYD_LLF_ISO: return launch_fluxes_t<RECON, RS_HYD_LLF_ISO>(g, eos, a, st);
    case RS_HYD_HLLE_ISO: return launch_fluxes_t<RECON, RS_HYD_HLLE_ISO>(g, eos, a, st);
    case RS_HYD_ROE_ISO: return launch_fluxes_t<RECON, RS_HYD_ROE_ISO>(g, eos, a, st);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_hydro_fluxes(int recon, int rs, const Geom &g, const EosParams &eos, const Arrays &a, cudaStream_t st) {
  switch (recon) {
    case RECON_DC: return launch_hydro_fluxes_r<RECON_DC>(rs, g, eos, a, st);
    case RECON_PLM: return launch_hydro_fluxes_r<RECON_PLM>(rs, g, eos, a, st);
    case RECON_PPM4: return launch_hydro_fluxes_r<RECON_PPM4>(rs, g, eos, a, st);
    case RECON_PPMX: return launch_hydro_fluxes_r<RECON_PPMX>(rs, g, eos, a, st);
    case RECON_WENOZ: return launch_hydro_fluxes_r<RECON_WENOZ>(rs, g, eos, a, st);
    case RECON_TENO: return launch_hydro_fluxes_r<RECON_TENO>(rs, g, eos, a, st);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_hydro_cons2prim(const Geom &g, const EosParams &eos, const Arrays &a, int il, int iu, int jl, int ju,
                                   int kl, int ku, int *counters, bool newdt, unsigned long long *dtmin, cudaStream_t st) {
  const long n = (long)g.nmb * (ku - kl + 1) * (ju - jl + 1) * (iu - il + 1);
  if (n <= 0) return cudaSuccess;
#define AKB_HC2P_ARGS g, eos, a.u0, a.w0, il, iu, jl, ju, kl, ku, counters, a.mb_size, dtmin
  const bool ideal = (g.nfl == 5);
  if (newdt && ideal) hydro_cons2prim_kernel<true, true><<<nblocks(n, 256), 256, 0, st>>>(AKB_HC2P_ARGS);
  else if (ideal) hydro_cons2prim_kernel<false, true><<<nblocks(n, 256), 256, 0, st>>>(AKB_HC2P_ARGS);
  else if (newdt) hydro_cons2prim_kernel<true, false><<<nblocks(n, 256), 256, 0, st>>>(AKB_HC2P_ARGS);
  else hydro_cons2prim_kernel<false, false><<<nblocks(n, 256), 256, 0, st>>>(AKB_HC2P_ARGS);
#undef AKB_HC2P_ARGS
  return cudaGetLastError();
}

cudaError_t launch_hydro_newdt(const Geom &g, const EosParams &eos, const Arrays &a, unsigned long long *dtmin, cudaStream_t st) {
  const long n = (long)g.nmb * g.nx3 * g.nx2 * g.nx1;
  if (g.nfl == 5) hydro_newdt_kernel<true><<<nblocks(n, 256), 256, 0, st>>>(g, eos, a.w0, a.mb_size, dtmin);
  else hydro_newdt_kernel<false><<<nblocks(n, 256), 256, 0, st>>>(g, eos, a.w0, a.mb_size, dtmin);
  return cudaGetLastError();
}

}  // namespace akb
