// flux_kernels.cuh -- the reconstruction + Riemann sweep kernels, shared by the MHD path
// (mhd_kernels.cu instantiates them for LLF/HLLE/HLLD) and its hydro twin (hydro_kernels.cu:
// advect/LLF/HLLE/HLLC/Roe).  See mhd_kernels.cu for the kernel inventory.
#pragma once
#include <cstdlib>
#include "mhd_kernels.h"
#include "mhd_device.cuh"
#include "hydro_device.cuh"
#include "iso_device.cuh"

#ifndef AKB_MARCH_PREFETCH
#define AKB_MARCH_PREFETCH 0
#endif
#ifndef AKB_CORNER_MINB
#define AKB_CORNER_MINB 2   // 128-register budget: measured 0.56 ms vs 0.96 ms unbounded, 0.84-0.88 ms at 3-4
#endif
#ifndef AKB_X1_MINB
#define AKB_X1_MINB 4
#endif
#ifndef AKB_MARCH_MINB
#define AKB_MARCH_MINB 4
#endif
#ifndef AKB_X1_L2PF
#define AKB_X1_L2PF 0       // x1 sweep: planes ahead whose stencil lines each warp prefetches into L2 (0 = off)
#endif
#ifndef AKB_MARCH_TMA_DEFAULT
#define AKB_MARCH_TMA_DEFAULT 0   // x2/x3 sweeps of the 5-point reconstructions through flux_march_tma_kernel (AKB_MARCH_TMA=0/1 overrides)
#endif
#ifndef AKB_MARCH_SHARE
#define AKB_MARCH_SHARE 1   // PPM4 x2/x3 sweeps: one interface evaluation per face, carried through shared memory (flux_march_kernel)
#endif
#ifndef AKB_MARCH_CARRY
#define AKB_MARCH_CARRY 0   // PPM4 x2/x3 sweeps: carry the clamped interface value along the march.
                            // Measured SLOWER (1.12 vs 1.05 ms): -10% FP64 instructions but +14 live registers
                            // under the 128-register budget turn into local-memory spills inside the loop.
#endif

namespace akb {

// Which of the seven State7 members (d, vx, vy, vz, e, by, bz) a solver family reconstructs:
// ideal MHD all seven, isothermal MHD six (no e), ideal hydro five, isothermal hydro four.
template <int RS>
struct SolverTraits {
  static constexpr bool MHD = !rs_is_hydro(RS);
  static constexpr bool IDEAL = !rs_is_iso(RS);
  static constexpr int NV = (MHD ? 6 : 4) + (IDEAL ? 1 : 0);
  // member index of reconstructed variable v
  __host__ __device__ static constexpr int member(int v) { return (IDEAL || v < 4) ? v : v + 1; }
};

template <int RS>
__device__ __forceinline__ void riemann_any(const State7 &l, const State7 &r, double bxi, const EosParams &eos, Flux7 &f,
                                            bool *oor = nullptr) {   // oor: deferred range guard (mhd_device.cuh), threaded through the headline solver
  if constexpr (RS == RS_LLF_ISO) llf_iso(l, r, bxi, eos, f, oor);
  else if constexpr (RS == RS_HLLE_ISO) hlle_iso(l, r, bxi, eos, f, oor);
  else if constexpr (RS == RS_HLLD_ISO) hlld_iso(l, r, bxi, eos, f, oor);
  else if constexpr (RS == RS_HYD_LLF_ISO) hyd_llf_iso(l, r, eos, f, oor);
  else if constexpr (RS == RS_HYD_HLLE_ISO) hyd_hlle_iso(l, r, eos, f, oor);
  else if constexpr (RS == RS_HYD_ROE_ISO) hyd_roe_iso(l, r, eos, f, oor);
  else if constexpr (rs_is_hydro(RS)) riemann_hyd<RS>(l, r, eos, f, oor);
  else riemann<RS>(l, r, bxi, eos, f, oor);
}

__device__ __forceinline__ void prefetch_l1(const double *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// ---------------------------------------------------------------------------------------
// x1 sweep.  One grid row (blockIdx.y) per (m, k in [ks-1,ke+1]) plane; inside the plane the
// cells (j in [js-1,je+1], c in [0,nx1+2)) are flattened and warp w owns cells
// STEP*w .. STEP*w+31.  Index decode is one multiply-high by a host-computed reciprocal of
// nx1+2 (exact for the extents checked by the launcher) instead of three 64-bit divisions.
//
// PPM4: each lane evaluates only the clamped interface value at ITS lower face and takes the
// one at its upper face from lane+1 (bit-identical, see ppm4_iface), so the interface
// arithmetic is done once per face instead of twice; lanes 0..30 then hold (up, lo), lane l
// gets the left state from lane l-1 and solves the face between them: STEP = 30 faces per warp.
// Each row carries one helper cell (c = nx1+2, i = ie+2) whose only job is the interface at
// the upper face of the row's last real cell.  Other methods reconstruct per cell from L1:
// STEP = 31, no helper cell.
template <int RECON>
struct X1Step { static constexpr int value = (RECON == RECON_PPM4) ? 30 : 31; };

// Flux + update fusion (ak_mhd_fluxes_update): the sweeps never write the flux arrays.  Each sweep turns the fluxes at the
// two faces of a cell into that direction's term of the flux divergence while they are in registers -- x1 by a shuffle from
// lane+1, x2 / x3 from the previous step of the march -- and stores 5 words per cell instead of 5 per face:
//   x1: d1 = (F1(i+1) - F1(i)) / dx1 -> du1;   x2: d2 -> du2;   x3: divf = du1; divf += du2; divf += d3;
//       u0 = gam0*u0 + gam1*u1 - beta_dt*divf (+ the stage-1 CopyCons of u)       [mhd_update.cpp:51-82, same operation order]
// Only the mass flux (corner-EMF upwinding) and the face EMFs still go to HBM.  Against flux arrays + update kernel this
// removes 15 words written + 15 read per cell per stage and one kernel; the divisions ride in kernels that are bound by FP64
// issue, not by HBM.  ACC: 0 = plain sweep, 1 = store this direction's term, 2 = x3: finish the update.
struct AccArgs {
  double *du1, *du2;             // nmb x nfl x cc each
  const double *mb_size;
  double *u0, *u1;
  double gam0, gam1, beta_dt;
  int save;                      // stage 1: u1 <- old u0 on the updated cells
};

// left state of a face = upper-face state of the cell in lane-1
template <int RS>
__device__ __forceinline__ void shift_up_state(const State7 &up, State7 &L) {
  using T = SolverTraits<RS>;
  L.d = __shfl_up_sync(0xffffffffu, up.d, 1);
  L.vx = __shfl_up_sync(0xffffffffu, up.vx, 1);
  L.vy = __shfl_up_sync(0xffffffffu, up.vy, 1);
  L.vz = __shfl_up_sync(0xffffffffu, up.vz, 1);
  if constexpr (T::IDEAL) L.e = __shfl_up_sync(0xffffffffu, up.e, 1);
  if constexpr (T::MHD) {
    L.by = __shfl_up_sync(0xffffffffu, up.by, 1);
    L.bz = __shfl_up_sync(0xffffffffu, up.bz, 1);
  }
}

// Cold side of the x1 sweep's deferred range guard: the whole warp repeats reconstruction and Riemann solve with the
// guarded (IEEE-exact for every operand) operators, re-reading the stencil.  Called warp-uniformly.
template <int RECON, int RS>
__device__ __noinline__ void x1_face_guarded(const double *__restrict__ wc, const double *__restrict__ bc, long cc, double bxi,
                                             bool act, EosParams eos, Flux7 &f) {
  using T = SolverTraits<RS>;
  State7 up, lo, L;
  if constexpr (RECON == RECON_PPM4) {
    const double *vp[7] = {wc + IDN * cc, wc + IVX * cc, wc + IVY * cc, wc + IVZ * cc, wc + IEN * cc, bc + 1 * cc, bc + 2 * cc};
    double *upv[7] = {&up.d, &up.vx, &up.vy, &up.vz, &up.e, &up.by, &up.bz};
    double *lov[7] = {&lo.d, &lo.vx, &lo.vy, &lo.vz, &lo.e, &lo.by, &lo.bz};
#pragma unroll
    for (int v = 0; v < T::NV; ++v) {
      const double *pv = vp[T::member(v)];
      const double q = __ldg(pv);
      const double fl = ppm4_iface(__ldg(pv - 2), __ldg(pv - 1), q, __ldg(pv + 1), AKB_COLD);
      const double fh = __shfl_down_sync(0xffffffffu, fl, 1);
      ppm4_mono(q, fl, fh, *upv[T::member(v)], *lov[T::member(v)]);
    }
  } else {
    recon_cell<RECON, 0, T::MHD, T::IDEAL>(wc, bc, cc, 1, eos, up, lo, AKB_COLD);
  }
  shift_up_state<RS>(up, L);
  if (act) riemann_any<RS>(L, lo, bxi, eos, f, AKB_COLD);
}

template <int RECON, int RS, bool ACC = false>
__global__ void __launch_bounds__(128, AKB_X1_MINB)
flux_x1_kernel(Geom g, EosParams eos, const double *__restrict__ w0, const double *__restrict__ bcc,
               const double *__restrict__ bx1f, double *__restrict__ flx, double *__restrict__ e3x1,
               double *__restrict__ e2x1, unsigned rcp_nc, AccArgs acc) {
  constexpr int NFL = X1Step<RECON>::value;     // lanes 1..NFL solve a face
  constexpr int STEP = NFL - (ACC ? 1 : 0);     // warp stride; ACC: lane NFL's face only closes lane NFL-1's cell
  using T = SolverTraits<RS>;
  constexpr bool MHD = T::MHD;       // hydro twin: no face field, no EMFs
  constexpr bool IDEAL = T::IDEAL;   // isothermal EOS: no energy variable
  constexpr int NV = T::NV;
  const int lane = threadIdx.x & 31;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  // g.next = 1 (FOFC): faces [is-1,ie+2] instead of [is,ie+1] (mhd_fluxes.cpp:96-106)
  const int nc = g.nx1 + 2 + 2 * g.next + (RECON == RECON_PPM4 ? 1 : 0), nj = g.nx2 + 2 * g.ext2, nk = g.nx3 + 2 * g.ext3;
  const unsigned cell = warp * STEP + lane;            // cell index inside the plane
  const bool valid = cell < (unsigned)(nj * nc);
  const int m = blockIdx.y / nk;
  const int k = g.ks - g.ext3 + (blockIdx.y - m * nk);
  const unsigned jr = valid ? __umulhi(cell, rcp_nc) : 0u;   // cell / nc
  const int c = valid ? (int)(cell - jr * nc) : 1;
  const int j = g.js - g.ext2 + (int)jr;
  const int i = g.is - 1 - g.next + c;
  const long cc = g.cc;
  const long off = ((long)k * g.n2 + j) * g.n1 + i;
  // The face field is fetched FIRST, with the stencil loads: issued next to its use (after the
  // early exit) it was the single largest stall site of the kernel (14% of the samples).
  double bxi = 0.0;
  if constexpr (MHD) bxi = __ldg(bx1f + (((long)m * g.n3 + k) * g.n2 + j) * (g.n1 + 1) + i);
  if constexpr (AKB_X1_L2PF > 0) {
    // One instruction per warp: lane l touches variable (l mod NV) at its own cell, AKB_X1_L2PF planes
    // ahead, so every 128-byte line the counterpart warp of that plane will read is requested once.
    if (k + AKB_X1_L2PF <= g.ke + g.ext3) {
      const int v = T::member(lane % NV);
      constexpr int vidx[7] = {IDN, IVX, IVY, IVZ, IEN, 1, 2};
      const double *src = (v < 5) ? w0 + ((long)m * g.nvar + vidx[v]) * cc : bcc + ((long)m * 3 + vidx[v]) * cc;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(src + off + (long)AKB_X1_L2PF * g.n1 * g.n2));
    }
  }
  // Range guard of the guard-free division / square root (mhd_device.cuh), deferred form: the arithmetic below only
  // accumulates "an operand left the fast path's exact range" into oor; a warp that sees the flag anywhere (a lane's
  // interface value feeds lane-1, its upper state lane+1) repeats recon + Riemann with the guarded operators in a
  // separate function that re-reads its inputs, so nothing stays live across the hot path for its sake.
  bool oor = false;
  State7 up, lo;
  const bool act = valid && lane != 0 && lane <= NFL && c != 0 && c <= g.nx1 + 1 + 2 * g.next;
  State7 L;
  Flux7 f;
  const double *wc = w0 + (long)m * g.nvar * cc + off, *bc = MHD ? bcc + (long)m * 3 * cc + off : wc;
  if constexpr (RECON == RECON_PPM4) {
    const double *vp[7] = {wc + IDN * cc, wc + IVX * cc, wc + IVY * cc, wc + IVZ * cc, wc + IEN * cc, bc + 1 * cc, bc + 2 * cc};
    double q[NV], fl[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const double *pv = vp[T::member(v)];
      q[v] = __ldg(pv);
      fl[v] = ppm4_iface(__ldg(pv - 2), __ldg(pv - 1), q[v], __ldg(pv + 1), &oor);
    }
    double *upv[7] = {&up.d, &up.vx, &up.vy, &up.vz, &up.e, &up.by, &up.bz};
    double *lov[7] = {&lo.d, &lo.vx, &lo.vy, &lo.vz, &lo.e, &lo.by, &lo.bz};
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const double fh = __shfl_down_sync(0xffffffffu, fl[v], 1);   // interface at this cell's upper face
      ppm4_mono(q[v], fl[v], fh, *upv[T::member(v)], *lov[T::member(v)]);
    }
  } else {
    recon_cell<RECON, 0, MHD, IDEAL>(wc, bc, cc, 1, eos, up, lo, &oor);
  }
  shift_up_state<RS>(up, L);
  if (act) riemann_any<RS>(L, lo, bxi, eos, f, &oor);
  if (__any_sync(0xffffffffu, oor)) {     // results come back through a separate object so that f itself stays in registers
    Flux7 fg;
    x1_face_guarded<RECON, RS>(wc, bc, cc, bxi, act, eos, fg);
    f = fg;
  }
  const long fs = (long)g.n3 * g.n2 * (g.n1 + 1);
  double *fp = flx + (long)m * g.nvar * fs + ((long)k * g.n2 + j) * (g.n1 + 1) + i;
  if constexpr (ACC) {
    // x1 term of the flux divergence of this lane's cell: its upper face is lane+1's (every lane takes part in the shuffles)
    const double nd = __shfl_down_sync(0xffffffffu, f.d, 1), nmx = __shfl_down_sync(0xffffffffu, f.mx, 1);
    const double nmy = __shfl_down_sync(0xffffffffu, f.my, 1), nmz = __shfl_down_sync(0xffffffffu, f.mz, 1);
    double ne = 0.0;
    if constexpr (IDEAL) ne = __shfl_down_sync(0xffffffffu, f.e, 1);
    const bool cell = act && lane < NFL && i >= g.is && i <= g.ie && j >= g.js && j <= g.je && k >= g.ks && k <= g.ke;
    if (cell) {
      const double dx1 = __ldg(acc.mb_size + m * 9 + 6), rdx1 = drcp(dx1);
      bool o2 = false;
      double d0 = div_by(nd - f.d, dx1, rdx1, &o2), d1 = div_by(nmx - f.mx, dx1, rdx1, &o2), d2 = div_by(nmy - f.my, dx1, rdx1, &o2);
      double d3 = div_by(nmz - f.mz, dx1, rdx1, &o2), d4 = IDEAL ? div_by(ne - f.e, dx1, rdx1, &o2) : 0.0;
      if (o2) {
        d0 = div_by(nd - f.d, dx1, rdx1); d1 = div_by(nmx - f.mx, dx1, rdx1); d2 = div_by(nmy - f.my, dx1, rdx1);
        d3 = div_by(nmz - f.mz, dx1, rdx1); if (IDEAL) d4 = div_by(ne - f.e, dx1, rdx1);
      }
      double *dp = acc.du1 + (long)m * g.nfl * cc + off;
      dp[IDN * cc] = d0; dp[IM1 * cc] = d1; dp[IM2 * cc] = d2; dp[IM3 * cc] = d3;
      if constexpr (IDEAL) dp[IEN * cc] = d4;
    }
    if (!act) return;
    fp[IDN * fs] = f.d;               // mass flux: upwinding of the corner EMFs (mhd_corner_e.cpp:342-413)
  } else {
    if (!act) return;
    fp[IDN * fs] = f.d;
    fp[IVX * fs] = f.mx;
    fp[IVY * fs] = f.my;
    fp[IVZ * fs] = f.mz;
    if constexpr (IDEAL) fp[IEN * fs] = f.e;
  }
  if constexpr (MHD) {
    e3x1[(long)m * cc + off] = f.ey;  // mhd_fluxes.cpp:146
    e2x1[(long)m * cc + off] = f.ez;
  }
}

// x2 / x3 sweeps.  Thread = (m, segment, transverse index, i-cell); marches `len` faces.
//
// 5-point reconstructions keep a rolling window of the sweep in SHARED memory: a ring of
// MARCH_SLOTS rows x 7 variables per thread column (thread-private slots, so no barriers and
// no bank conflicts: consecutive threads own consecutive doubles).  Each step issues ONE new
// row per variable with cp.async (global -> shared, no register staging) one step ahead of its
// use, so the stencil reads are LDS and the HBM/L2 latency of the new row hides behind the
// Riemann solve of the current face.  3-point methods (PLM/DC) read global memory directly.
//
// Three loop forms live in flux_march_kernel:
//   * PPM4, plain sweep (the headline): shared-interface form -- one ppm4_iface per face, carried values / left state / next
//     face's normal field all in the thread's shared memory, 5-row ring, range-guard repeats after the march (march_fix_faces);
//   * other 5-point reconstructions, plain sweep: 6-row ring, left state through shared memory, repeats after the march;
//   * flux + update fusion (ACC != 0, opt-in): the face's result feeds the divergence term of the same step, so a flagged face
//     is repeated inside the loop (march_face_guarded), left state carried in registers.
#define TILE_X 32
#define TILE_Y 4
#define TILE_Z 2
#ifndef MARCH_SLOTS
#define MARCH_SLOTS 6          // 5 stencil rows + prefetch distance (6,7,8 measured: 3.16/3.26/3.27 ms)
#endif
#define MARCH_PF (MARCH_SLOTS - 5)
#define MARCH_THREADS 128

__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int RECON, int SS = MARCH_THREADS>
__device__ __forceinline__ void recon_ring_var(const double *__restrict__ rv, int b0, double &q_up, double &q_lo, bool *oor = nullptr) {
  // rv: this thread's ring for one variable (slot stride SS doubles); b0 = slot of row c-2
  int b1 = b0 + 1; if (b1 >= MARCH_SLOTS) b1 -= MARCH_SLOTS;
  int b2 = b1 + 1; if (b2 >= MARCH_SLOTS) b2 -= MARCH_SLOTS;
  int b3 = b2 + 1; if (b3 >= MARCH_SLOTS) b3 -= MARCH_SLOTS;
  int b4 = b3 + 1; if (b4 >= MARCH_SLOTS) b4 -= MARCH_SLOTS;
  const double qm2 = rv[b0 * SS], qm1 = rv[b1 * SS], q = rv[b2 * SS];
  const double qp1 = rv[b3 * SS], qp2 = rv[b4 * SS];
  if constexpr (RECON == RECON_PPM4) ppm4(qm2, qm1, q, qp1, qp2, q_up, q_lo, oor);
  else if constexpr (RECON == RECON_PPMX) ppmx(qm2, qm1, q, qp1, qp2, q_up, q_lo, oor);
  else if constexpr (RECON == RECON_TENO) teno(qm2, qm1, q, qp1, qp2, q_up, q_lo, oor);
  else wenoz(qm2, qm1, q, qp1, qp2, q_up, q_lo, oor);
}

template <int RECON, int RS, int SS = MARCH_THREADS>
__device__ __forceinline__ void recon_ring_cell(const double *__restrict__ ring, int b0, const EosParams &eos,
                                                State7 &up, State7 &lo, bool *oor = nullptr) {
  using T = SolverTraits<RS>;
  constexpr int VS = MARCH_SLOTS * SS;   // ring stride between variables
  double *upv[7] = {&up.d, &up.vx, &up.vy, &up.vz, &up.e, &up.by, &up.bz};
  double *lov[7] = {&lo.d, &lo.vx, &lo.vy, &lo.vz, &lo.e, &lo.by, &lo.bz};
#pragma unroll
  for (int v = 0; v < T::NV; ++v) recon_ring_var<RECON, SS>(ring + v * VS, b0, *upv[T::member(v)], *lov[T::member(v)], oor);
  if constexpr (RECON == RECON_PPMX || RECON == RECON_WENOZ || RECON == RECON_TENO) {   // recon.hpp:78-108
    up.d = dmax(up.d, eos.dfloor);
    lo.d = dmax(lo.d, eos.dfloor);
    if constexpr (T::IDEAL) {
      const double efloor = eos.pfloor / (eos.gamma - 1.0);
      up.e = dmax(up.e, efloor);
      lo.e = dmax(lo.e, efloor);
    }
  }
}

// PPM4 with the interface value carried along the march: the clamped interface at the upper
// face of cell c (rows c-1..c+2) is computed once and reused as the lower interface of cell
// c+1 on the next step (bit-identical, see ppm4_iface).  fl[] holds the lower interfaces of
// the current cell on entry and its upper ones on exit.  b0 = slot of row c-2.
__device__ __forceinline__ void ppm4_ring_iface(const double *__restrict__ ring, int b0, double fl[7]) {
  // interface between rows c-1 and c: slots of rows c-2, c-1, c, c+1
  constexpr int VS = MARCH_SLOTS * MARCH_THREADS;
  int b1 = b0 + 1; if (b1 >= MARCH_SLOTS) b1 -= MARCH_SLOTS;
  int b2 = b1 + 1; if (b2 >= MARCH_SLOTS) b2 -= MARCH_SLOTS;
  int b3 = b2 + 1; if (b3 >= MARCH_SLOTS) b3 -= MARCH_SLOTS;
#pragma unroll
  for (int v = 0; v < 7; ++v) {
    const double *rv = ring + v * VS;
    fl[v] = ppm4_iface(rv[b0 * MARCH_THREADS], rv[b1 * MARCH_THREADS], rv[b2 * MARCH_THREADS], rv[b3 * MARCH_THREADS]);
  }
}
__device__ __forceinline__ void ppm4_ring_cell(const double *__restrict__ ring, int b0, double fl[7], State7 &up, State7 &lo) {
  constexpr int VS = MARCH_SLOTS * MARCH_THREADS;
  int b1 = b0 + 1; if (b1 >= MARCH_SLOTS) b1 -= MARCH_SLOTS;
  int b2 = b1 + 1; if (b2 >= MARCH_SLOTS) b2 -= MARCH_SLOTS;
  int b3 = b2 + 1; if (b3 >= MARCH_SLOTS) b3 -= MARCH_SLOTS;
  int b4 = b3 + 1; if (b4 >= MARCH_SLOTS) b4 -= MARCH_SLOTS;
  double *upv[7] = {&up.d, &up.vx, &up.vy, &up.vz, &up.e, &up.by, &up.bz};
  double *lov[7] = {&lo.d, &lo.vx, &lo.vy, &lo.vz, &lo.e, &lo.by, &lo.bz};
#pragma unroll
  for (int v = 0; v < 7; ++v) {
    const double *rv = ring + v * VS;
    const double q = rv[b2 * MARCH_THREADS];
    const double fh = ppm4_iface(rv[b1 * MARCH_THREADS], q, rv[b3 * MARCH_THREADS], rv[b4 * MARCH_THREADS]);
    ppm4_mono(q, fl[v], fh, *upv[v], *lov[v]);
    fl[v] = fh;
  }
}

// Cold side of the march kernels' deferred range guard (see flux_x1_kernel): one face repeated with the guarded operators.
// The left state comes back from the thread's shared-memory copy (the hot path does not keep it live for this), the cell's
// own states are re-reconstructed from the ring.
template <int RECON, int RS, int SS = MARCH_THREADS>
__device__ __noinline__ void march_face_guarded(const double *__restrict__ ring, int b0, const double *__restrict__ lsm, double bni,
                                                EosParams eos, State7 &up, Flux7 &fx) {
  State7 L, lo;
  L.d = lsm[0 * MARCH_THREADS]; L.vx = lsm[1 * MARCH_THREADS]; L.vy = lsm[2 * MARCH_THREADS]; L.vz = lsm[3 * MARCH_THREADS];
  L.e = lsm[4 * MARCH_THREADS]; L.by = lsm[5 * MARCH_THREADS]; L.bz = lsm[6 * MARCH_THREADS];
  recon_ring_cell<RECON, RS, SS>(ring, b0, eos, up, lo, AKB_COLD);
  riemann_any<RS>(L, lo, bni, eos, fx, AKB_COLD);
}
__device__ __forceinline__ void load_left_state(const double *lsm, State7 &s) {
  s.d = lsm[0 * MARCH_THREADS]; s.vx = lsm[1 * MARCH_THREADS]; s.vy = lsm[2 * MARCH_THREADS]; s.vz = lsm[3 * MARCH_THREADS];
  s.e = lsm[4 * MARCH_THREADS]; s.by = lsm[5 * MARCH_THREADS]; s.bz = lsm[6 * MARCH_THREADS];
}
// Repeat side of the plain march kernels' deferred range guard: the faces marked in `fix` (bit b = face f_lo + b) are solved again
// from global memory with the guarded operators -- the upper state of the cell below, the lower state of the cell above, the
// Riemann problem between them -- and their fluxes / face EMFs stored over the fast path's.  Same functions, same operands, so
// an unflagged quantity comes out with the bits it had.  co0 / fo0: cell f_lo-1 and face f_lo of this thread's column.
template <int RECON, int RS, int DIR>
__device__ __noinline__ void march_fix_faces(unsigned fix, int f_lo, int f_hi, int co0, int fo0, int s, const double *__restrict__ wb,
                                             const double *__restrict__ bb, const double *__restrict__ bnb, long cc, EosParams eos,
                                             double *__restrict__ fb, long fsz, double *__restrict__ eyb, double *__restrict__ ezb) {
  using T = SolverTraits<RS>;
  constexpr int ivx = IVX + DIR, ivy = IVX + (DIR + 1) % 3, ivz = IVX + (DIR + 2) % 3;
  State7 L, lo, up_above;
  int prev = -2;                     // bit of the face solved last: marked faces come in runs, and the cell above face f is the one below f+1
  while (fix) {
    const int b = __ffs((int)fix) - 1;
    fix &= fix - 1u;
    if (f_lo + b > f_hi) break;
    const int cb = co0 + b * s, ca = cb + s, fo = fo0 + b * s;     // cell below face f, cell above it, the face
    if (b == prev + 1) L = up_above;
    else { State7 unused; recon_cell<RECON, DIR, T::MHD, T::IDEAL>(wb + cb, bb + cb, cc, s, eos, L, unused, AKB_COLD); }
    recon_cell<RECON, DIR, T::MHD, T::IDEAL>(wb + ca, bb + ca, cc, s, eos, up_above, lo, AKB_COLD);
    prev = b;
    double bni = 0.0;
    if constexpr (T::MHD) bni = __ldg(bnb + fo);
    Flux7 fx;
    riemann_any<RS>(L, lo, bni, eos, fx, AKB_COLD);
    double *fp = fb + fo;
    fp[IDN * fsz] = fx.d; fp[ivx * fsz] = fx.mx; fp[ivy * fsz] = fx.my; fp[ivz * fsz] = fx.mz;
    if constexpr (T::IDEAL) fp[IEN * fsz] = fx.e;
    if constexpr (T::MHD) { eyb[ca] = fx.ey; ezb[ca] = fx.ez; }
  }
}
template <int RECON, int RS, int SS = MARCH_THREADS>
__device__ __noinline__ void march_recon_guarded(const double *__restrict__ ring, int b0, EosParams eos, State7 &up) {
  State7 lo;
  recon_ring_cell<RECON, RS, SS>(ring, b0, eos, up, lo, AKB_COLD);
}
__device__ __forceinline__ void store_left_state(double *lsm, const State7 &s) {
  lsm[0 * MARCH_THREADS] = s.d; lsm[1 * MARCH_THREADS] = s.vx; lsm[2 * MARCH_THREADS] = s.vy; lsm[3 * MARCH_THREADS] = s.vz;
  lsm[4 * MARCH_THREADS] = s.e; lsm[5 * MARCH_THREADS] = s.by; lsm[6 * MARCH_THREADS] = s.bz;
}

// Addressing is kept out of the register file on purpose: the block's (MeshBlock, segment) come
// from blockIdx.y, so every array base is block-uniform, and a thread carries only two 32-bit
// running offsets (cell, face) that advance by the sweep stride.  The earlier version carried a
// dozen 64-bit pointers through the loop; under the 128-register budget ptxas spilled loop
// invariants to local memory and their reloads (L1-thrashed, long scoreboard) were the top stall
// site of the kernel (profiles/r01_ncu_kernels_session2.txt).
template <int RECON, int RS, int DIR, int ACC = 0>
__global__ void __launch_bounds__(MARCH_THREADS, AKB_MARCH_MINB)
flux_march_kernel(Geom g, EosParams eos, const double *__restrict__ w0, const double *__restrict__ bcc,
                  const double *__restrict__ bnf, double *__restrict__ flx, double *__restrict__ eyo,
                  double *__restrict__ ezo, int seglen, int nseg, AccArgs acc) {
  static_assert(DIR == 1 || DIR == 2, "march kernel handles x2 and x3");
  constexpr bool RING = (RECON >= RECON_PPM4);
  using T = SolverTraits<RS>;
  constexpr bool MHD = T::MHD;
  constexpr bool IDEAL = T::IDEAL;
  constexpr int NV = T::NV;
  extern __shared__ double ring_all[];   // dynamic: RING: (NV * MARCH_SLOTS + 7) * MARCH_THREADS doubles (ring + left-state copy); ACC: + 5 * MARCH_THREADS
  const int nc = g.nx1 + 2 * g.ext1;
  const int nt = (DIR == 1) ? g.nx3 + 2 * g.ext3 : g.nx2 + 2 * g.ext2;
  const int m = blockIdx.y / nseg, seg = blockIdx.y - m * nseg;       // block-uniform
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= nt * nc) return;
  const int tr = col / nc;
  const int i = g.is - g.ext1 + (col - tr * nc);
  const int fs0 = ((DIR == 1) ? g.js : g.ks) - g.next;         // first face along the sweep (one more with FOFC)
  const int fe0 = ((DIR == 1) ? g.je + 1 : g.ke + 1) + g.next; // last face
  const int f_lo = fs0 + seg * seglen;
  int f_hi = f_lo + seglen - 1 + (ACC ? 1 : 0);     // ACC: a segment is seglen CELLS, closed by one more face (solved twice)
  if (f_hi > fe0) f_hi = fe0;
  if (f_lo > f_hi - (ACC ? 1 : 0)) return;
  const long cc = g.cc;
  const int s = (DIR == 1) ? g.n1 : g.n1 * g.n2;    // sweep stride (cells and faces alike)
  // cell (k,j,i) with the sweep index = f_lo-1, and the face f_lo above it
  int k, j;
  if (DIR == 1) { k = g.ks - g.ext3 + tr; j = f_lo - 1; } else { j = g.js - g.ext2 + tr; k = f_lo - 1; }
  int co = (k * g.n2 + j) * g.n1 + i;               // running cell offset inside one variable of the block
  int fo = (DIR == 1) ? (k * (g.n2 + 1) + f_lo) * g.n1 + i : (f_lo * g.n2 + j) * g.n1 + i;   // running face offset
  const long fsz = (DIR == 1) ? g.f2 : g.f3;
  // block-uniform bases
  const double *wb = w0 + (long)m * g.nvar * cc;
  const double *bb = MHD ? bcc + (long)m * 3 * cc : wb;
  const double *bnb = MHD ? bnf + (long)m * fsz : wb;
  double *fb = flx + (long)m * g.nvar * fsz;
  double *eyb = MHD ? eyo + (long)m * cc : nullptr;
  double *ezb = MHD ? ezo + (long)m * cc : nullptr;
  constexpr int ivx = IVX + DIR, ivy = IVX + (DIR + 1) % 3, ivz = IVX + (DIR + 2) % 3;
  constexpr int iby = (DIR + 1) % 3, ibz = (DIR + 2) % 3;
  // the NV reconstructed variables: source array and variable index (block-uniform)
  auto vsrc = [&](int vv) -> const double * {
    const int v = T::member(vv);
    return v == 0 ? wb + IDN * cc : v == 1 ? wb + ivx * cc : v == 2 ? wb + ivy * cc : v == 3 ? wb + ivz * cc
         : v == 4 ? wb + IEN * cc : v == 5 ? bb + iby * cc : bb + ibz * cc;
  };

  // ACC (flux + update fusion, see AccArgs): this direction's term of the flux divergence of the cell BELOW face f, from the
  // flux just solved and the previous step's (kept in thread-private shared memory, not in registers)
  double *pfs = ring_all + (RING ? (NV * MARCH_SLOTS + 7) * MARCH_THREADS : 0) + threadIdx.x;
  const bool tact = (DIR == 1) ? (k >= g.ks && k <= g.ke && i >= g.is && i <= g.ie) : (j >= g.js && j <= g.je && i >= g.is && i <= g.ie);
  double dxn = 1.0, rdxn = 1.0;
  if constexpr (ACC != 0) { dxn = __ldg(acc.mb_size + m * 9 + 6 + DIR); rdxn = drcp(dxn); }
  auto emit = [&](int f, const Flux7 &fx, int cof, int fof) {      // cof: cell above face f, fof: face f
    double *fp = fb + fof;
    fp[IDN * fsz] = fx.d;
    if constexpr (ACC == 0) {
      fp[ivx * fsz] = fx.mx;
      fp[ivy * fsz] = fx.my;
      fp[ivz * fsz] = fx.mz;
      if constexpr (IDEAL) fp[IEN * fsz] = fx.e;
    }
    if constexpr (MHD) {
      eyb[cof] = fx.ey;  // x2: (e1x2,e3x2)  x3: (e2x3,e1x3)   mhd_fluxes.cpp:196,243
      ezb[cof] = fx.ez;
    }
    if constexpr (ACC != 0) {
      const int cs = (DIR == 1) ? g.js : g.ks, ce = (DIR == 1) ? g.je : g.ke;
      if (f > f_lo && tact && f - 1 >= cs && f - 1 <= ce) {
        const int cb = cof - s;                                    // the cell below face f
        const double p0 = pfs[0 * MARCH_THREADS], p1 = pfs[1 * MARCH_THREADS], p2 = pfs[2 * MARCH_THREADS], p3 = pfs[3 * MARCH_THREADS];
        const double p4 = IDEAL ? pfs[4 * MARCH_THREADS] : 0.0;
        bool o2 = false;
        double d0 = div_by(fx.d - p0, dxn, rdxn, &o2), d1 = div_by(fx.mx - p1, dxn, rdxn, &o2), d2 = div_by(fx.my - p2, dxn, rdxn, &o2);
        double d3 = div_by(fx.mz - p3, dxn, rdxn, &o2), d4 = IDEAL ? div_by(fx.e - p4, dxn, rdxn, &o2) : 0.0;
        if (o2) {
          d0 = div_by(fx.d - p0, dxn, rdxn); d1 = div_by(fx.mx - p1, dxn, rdxn); d2 = div_by(fx.my - p2, dxn, rdxn);
          d3 = div_by(fx.mz - p3, dxn, rdxn); if (IDEAL) d4 = div_by(fx.e - p4, dxn, rdxn);
        }
        const long mb = (long)m * g.nfl * cc;
        if constexpr (ACC == 1) {
          double *dp = acc.du2 + mb + cb;
          dp[IDN * cc] = d0; dp[ivx * cc] = d1; dp[ivy * cc] = d2; dp[ivz * cc] = d3;
          if constexpr (IDEAL) dp[IEN * cc] = d4;
        } else {
          // x3: mhd_update.cpp:51-82 -- divf = x1 term; += x2 term; += x3 term; u0 = gam0*u0 + gam1*u1 - beta_dt*divf
          const double *q1 = acc.du1 + mb + cb, *q2 = acc.du2 + mb + cb;
          double *u0p = acc.u0 + mb + cb, *u1p = acc.u1 + mb + cb;
          auto upd = [&](int n, double d3rd) {
            double divf = q1[n * cc];
            divf += q2[n * cc];
            divf += d3rd;
            const double uold = u0p[n * cc];
            double ureg;
            if (acc.save) { ureg = uold; u1p[n * cc] = uold; } else { ureg = u1p[n * cc]; }
            u0p[n * cc] = acc.gam0 * uold + acc.gam1 * ureg - acc.beta_dt * divf;
          };
          upd(IDN, d0); upd(ivx, d1); upd(ivy, d2); upd(ivz, d3);
          if constexpr (IDEAL) upd(IEN, d4);
        }
      }
      pfs[0 * MARCH_THREADS] = fx.d; pfs[1 * MARCH_THREADS] = fx.mx; pfs[2 * MARCH_THREADS] = fx.my; pfs[3 * MARCH_THREADS] = fx.mz;
      if constexpr (IDEAL) pfs[4 * MARCH_THREADS] = fx.e;
    }
  };

  State7 L, lo_unused;
  if constexpr (RING && RECON == RECON_PPM4 && ACC == 0 && AKB_MARCH_SHARE) {
    // PPM4, plain sweep: shared-interface form of the march.  The clamped fourth-order value at the face between cells c and c+1 is
    // ONE call ppm4_iface(q[c-1], q[c], q[c+1], q[c+2]) whether cell c asks for its upper value or cell c+1 for its lower one -- same
    // function, same operands in the same order -- so each step evaluates only the interface ABOVE its cell and takes the one below
    // from the previous step: 7 interface evaluations (division by 12, two clamps, the guard's tests) per face instead of 14.  The
    // carried values live in the thread's shared memory next to the left state, not in registers (carried in registers the same
    // idea lost in round 1: 1.05 -> 1.12 ms), and the stencil no longer reaches row c-2, so the ring is 5 rows instead of 6: the
    // shared-memory footprint per CTA stays (6 NV + 7) x 128 doubles.  Range guard: faces are marked and solved again after the
    // march (march_fix_faces); a flagged interface taints both cells it belongs to, i.e. faces f, f+1 and f+2.
    constexpr int NS = 5;
    constexpr int VS = NS * MARCH_THREADS;
    double *ring = ring_all + threadIdx.x;
    double *lsm = ring_all + NV * VS + threadIdx.x;      // left state of the next face
    double *fsm = lsm + 7 * MARCH_THREADS;               // interface values at the lower face of the next cell
    // the face-normal field of the next face, fetched like the rows (two alternating slots).  As a register prefetch (__ldg one
    // step ahead) the value was spilled the moment it was loaded and the STL waited out the load: 10 % of the kernel's stall samples
    double *bsm = fsm + NV * MARCH_THREADS;
    // rows f_lo-1 .. f_lo+2 -> slots 0..3 (the first cell's rows c-1 .. c+2); co = cell f_lo-1
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int v = 0; v < NV; ++v) cp_async8(ring + v * VS + r * MARCH_THREADS, vsrc(v) + (co + r * s));
    }
    if constexpr (MHD) cp_async8(bsm, bnb + fo);
    cp_async_commit();
    // meanwhile the cell below the first face, straight from global memory: both interfaces, its upper state (the first left
    // state) and the interface it shares with cell f_lo
    L.d = L.vx = L.vy = L.vz = L.e = L.by = L.bz = 0.0;
    auto first_cell = [&](bool *oor) {
      double *upv[7] = {&L.d, &L.vx, &L.vy, &L.vz, &L.e, &L.by, &L.bz};
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const double *p = vsrc(v) + co;
        const double qm2 = __ldg(p - 2 * s), qm1 = __ldg(p - s), q = __ldg(p), qp1 = __ldg(p + s), qp2 = __ldg(p + 2 * s);
        const double fl = ppm4_iface(qm2, qm1, q, qp1, oor), fh = ppm4_iface(qm1, q, qp1, qp2, oor);
        double lo_v;
        ppm4_mono(q, fl, fh, *upv[T::member(v)], lo_v);
        fsm[v * MARCH_THREADS] = fh;
      }
    };
    {
      bool o0 = false;
      first_cell(&o0);
      if (o0) first_cell(AKB_COLD);
    }
    store_left_state(lsm, L);
    cp_async_wait<0>();
    unsigned fix = 0u;                // faces to be solved again after the march (bit f - f_lo)
    int b0 = 0;                       // slot of row c-1 for the next cell (c = f_lo)
    int bslot = 0;                    // slot of this face's normal field
    for (int f = f_lo; f <= f_hi; ++f) {
      co += s;                        // cell c = f
      int b1 = b0 + 1; if (b1 >= NS) b1 -= NS;
      int b2 = b1 + 1; if (b2 >= NS) b2 -= NS;
      int b3 = b2 + 1; if (b3 >= NS) b3 -= NS;
      int b4 = b3 + 1; if (b4 >= NS) b4 -= NS;     // held row c-2: free
      if (f + 1 <= f_hi) {            // row c+3 and the normal field of face f+1, needed one step from now
        const int ro = co + 3 * s;
#pragma unroll
        for (int v = 0; v < NV; ++v) cp_async8(ring + v * VS + b4 * MARCH_THREADS, vsrc(v) + ro);
        if constexpr (MHD) cp_async8(bsm + (bslot ^ 1) * MARCH_THREADS, bnb + fo + s);
      }
      cp_async_commit();
      cp_async_wait<1>();
      double bni = 0.0;
      if constexpr (MHD) bni = bsm[bslot * MARCH_THREADS];
      bslot ^= 1;
      State7 up, lo;
      Flux7 fx;
      bool orec = false, orie = false;
      load_left_state(lsm, L);
      {
        double *upv[7] = {&up.d, &up.vx, &up.vy, &up.vz, &up.e, &up.by, &up.bz};
        double *lov[7] = {&lo.d, &lo.vx, &lo.vy, &lo.vz, &lo.e, &lo.by, &lo.bz};
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const double *rv = ring + v * VS;
          const double q = rv[b1 * MARCH_THREADS];
          const double fh = ppm4_iface(rv[b0 * MARCH_THREADS], q, rv[b2 * MARCH_THREADS], rv[b3 * MARCH_THREADS], &orec);
          ppm4_mono(q, fsm[v * MARCH_THREADS], fh, *upv[T::member(v)], *lov[T::member(v)]);
          fsm[v * MARCH_THREADS] = fh;
        }
      }
      store_left_state(lsm, up);
      riemann_any<RS>(L, lo, bni, eos, fx, &orie);
      fix |= ((orec ? 7u : 0u) | (orie ? 1u : 0u)) << (f - f_lo);     // bits past 31 fall off: those faces belong to the next segment
      emit(f, fx, co, fo);
      fo += s;
      b0 = b1;
    }
    cp_async_wait<0>();
    if (fix) {     // co / fo have advanced by one row per face: cell f_lo-1 and face f_lo are that many rows back
      const int nrow = (f_hi - f_lo + 1) * s;
      march_fix_faces<RECON, RS, DIR>(fix, f_lo, f_hi, co - nrow, fo - nrow, s, wb, bb, bnb, cc, eos, fb, fsz, eyb, ezb);
    }
  } else if constexpr (RING) {
    constexpr int VS = MARCH_SLOTS * MARCH_THREADS;
    double *ring = ring_all + threadIdx.x;
    // prologue: rows c-2 .. c+2+PF -> slots 0..S-1 (row c-2 in slot 0), c = f_lo-1
#pragma unroll
    for (int r = 0; r < MARCH_SLOTS; ++r) {
      if (r < 5 || (f_lo - 1) + (r - 2) <= f_hi + 2) {   // prefetch rows only if some face of this march needs them
#pragma unroll
        for (int v = 0; v < NV; ++v) cp_async8(ring + v * VS + r * MARCH_THREADS, vsrc(v) + (co + (r - 2) * s));
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
    double fl[7];
    if constexpr (RECON == RECON_PPM4 && AKB_MARCH_CARRY && MHD && IDEAL) {
      ppm4_ring_iface(ring, 0, fl);
      ppm4_ring_cell(ring, 0, fl, L, lo_unused);
    } else {
      bool oor0 = false;      // deferred range guard (mhd_device.cuh): repeat with the guarded operators when flagged
      recon_ring_cell<RECON, RS>(ring, 0, eos, L, lo_unused, &oor0);
      if (oor0) { State7 lg; march_recon_guarded<RECON, RS>(ring, 0, eos, lg); L = lg; }
    }
    double *lsm = ring_all + NV * VS + threadIdx.x;     // this thread's copy of the left state
    store_left_state(lsm, L);
    unsigned fix = 0u;                // faces to be solved again after the march (bit f - f_lo)
    const int co0 = co, fo0 = fo;     // cell f_lo-1 and face f_lo
    int b0 = 1;                       // slot of row c-2 for the next cell (c = f_lo)
    // the normal field of a face is fetched one step ahead: issued next to its use it was the top stall site of the kernel
    double bn_next = 0.0;
    if constexpr (MHD) bn_next = __ldg(bnb + fo);
    for (int f = f_lo; f <= f_hi; ++f) {
      co += s;                        // cell c = f
      const double bni = bn_next;
      if constexpr (MHD) { if (f < f_hi) bn_next = __ldg(bnb + fo + s); }
      // rows c-2..c+2 are present; request row c+2+PF (needed PF steps from now) into the slot that held row c-3
      if (f + MARCH_PF <= f_hi) {
        int bn_slot = b0 + MARCH_SLOTS - 1; if (bn_slot >= MARCH_SLOTS) bn_slot -= MARCH_SLOTS;
        const int ro = co + (2 + MARCH_PF) * s;
#pragma unroll
        for (int v = 0; v < NV; ++v) cp_async8(ring + v * VS + bn_slot * MARCH_THREADS, vsrc(v) + ro);
      }
      cp_async_commit();
      cp_async_wait<MARCH_PF>();
      State7 up, lo;
      Flux7 fx;
      if constexpr (ACC == 0 && !(RECON == RECON_PPM4 && AKB_MARCH_CARRY && MHD && IDEAL)) {
        // Plain sweep: nothing but the flux of a face leaves the loop body, so a face whose arithmetic raised the range flag is only
        // MARKED here and solved again after the march (march_fix_faces: both cells re-reconstructed from global memory with the
        // guarded operators, flux and EMFs stored over the fast ones).  The loop then holds no call -- ptxas spilled ~25 more values
        // per face around the in-loop repeat, and with 200 KB of the SM's 256 KB carved out for the rings those spills miss L1 --
        // and the left state travels through the thread's shared-memory copy instead of 14 registers across the Riemann solve.
        bool orec = false, orie = false;
        load_left_state(lsm, L);
        recon_ring_cell<RECON, RS>(ring, b0, eos, up, lo, &orec);
        store_left_state(lsm, up);
        riemann_any<RS>(L, lo, bni, eos, fx, &orie);
        // a flagged reconstruction taints this face (lo) and the next one (up is its left state); bit 32 falls off: that face opens
        // the next segment, whose thread reconstructs its own left state
        fix |= ((orec ? 3u : 0u) | (orie ? 1u : 0u)) << (f - f_lo);
        emit(f, fx, co, fo);
      } else {
      bool oor = false;
      if constexpr (RECON == RECON_PPM4 && AKB_MARCH_CARRY && MHD && IDEAL) ppm4_ring_cell(ring, b0, fl, up, lo);
      else recon_ring_cell<RECON, RS>(ring, b0, eos, up, lo, &oor);
      riemann_any<RS>(L, lo, bni, eos, fx, &oor);
      if (oor) {       // cold: an operand left the fast path's exact range (separate result objects keep up / fx in registers)
        State7 ug; Flux7 fg;
        march_face_guarded<RECON, RS>(ring, b0, lsm, bni, eos, ug, fg);
        up = ug; fx = fg;
      }
      store_left_state(lsm, up);
      emit(f, fx, co, fo);
      L = up;
      }
      fo += s;
      b0 = (b0 + 1 >= MARCH_SLOTS) ? 0 : b0 + 1;
    }
    cp_async_wait<0>();
    if constexpr (ACC == 0) {
      if (fix) march_fix_faces<RECON, RS, DIR>(fix, f_lo, f_hi, co0, fo0, s, wb, bb, bnb, cc, eos, fb, fsz, eyb, ezb);
    }
  } else {
    {
      bool oor0 = false;
      recon_cell<RECON, DIR, MHD, IDEAL>(wb + co, bb + co, cc, s, eos, L, lo_unused, &oor0);
      if (oor0) recon_cell<RECON, DIR, MHD, IDEAL>(wb + co, bb + co, cc, s, eos, L, lo_unused, AKB_COLD);
    }
    unsigned fix = 0u;                // plain sweep: faces to be solved again after the march (bit f - f_lo), as in the ring forms
    const int co0 = co, fo0 = fo;     // cell f_lo-1 and face f_lo
    for (int f = f_lo; f <= f_hi; ++f) {
      co += s;
      State7 up, lo;
      bool orec = false, orie = false;
      recon_cell<RECON, DIR, MHD, IDEAL>(wb + co, bb + co, cc, s, eos, up, lo, &orec);
      Flux7 fx;
      double bni = 0.0;
      if constexpr (MHD) bni = __ldg(bnb + fo);
      riemann_any<RS>(L, lo, bni, eos, fx, &orie);
      if constexpr (ACC == 0) {
        // a flagged reconstruction taints this face (lo) and the next one (up is its left state)
        fix |= ((orec ? 3u : 0u) | (orie ? 1u : 0u)) << (f - f_lo);
      } else {
        if (orec | orie) {       // fusion: the result feeds this step's divergence term, repeat in place (L was verified when it was computed)
          recon_cell<RECON, DIR, MHD, IDEAL>(wb + co, bb + co, cc, s, eos, up, lo, AKB_COLD);
          riemann_any<RS>(L, lo, bni, eos, fx, AKB_COLD);
        }
      }
      emit(f, fx, co, fo);
      L = up;
      fo += s;
    }
    if constexpr (ACC == 0) {
      if (fix) march_fix_faces<RECON, RS, DIR>(fix, f_lo, f_hi, co0, fo0, s, wb, bb, bnb, cc, eos, fb, fsz, eyb, ezb);
    }
  }
}

// ---------------------------------------------------------------------------------------
// x2 / x3 sweeps, TMA form (5-point reconstructions, no flux+update fusion).  Same thread mapping, ring and arithmetic as
// flux_march_kernel; what changes is who fetches the ring rows.  There every thread issues one 8-byte cp.async per variable
// and step (7 LDGSTS + their addresses per thread and face); here ONE thread of the CTA issues one bulk-copy per variable and
// row segment -- cp.async.bulk.shared.global through the TMA unit (SASS UBLKCP), up to ~1 KB each, completion counted in bytes
// on an mbarrier -- and the other 127 threads only wait on that barrier.
//   * A CTA owns 128 consecutive columns of the flattened (transverse, i) index space, i.e. at most three row segments of
//     the array (rows of 64 columns or more).  Bulk copies need 16-byte aligned addresses and sizes, so each segment is fetched from its even-aligned start
//     over an even number of doubles (rows start on 16-byte boundaries because n1 is even -- checked by the launcher); a thread
//     reads its own column at a fixed offset `sidx` of the row buffer (MARCH_RB doubles per variable and slot).
//   * full[slot]: mbarrier armed with the bytes of the row about to land in `slot`; every thread waits for the phase of the row
//     c+2 it is about to read.  Re-use of a slot (the row fetched at step f overwrites the one last read at step f-1) is
//     ordered by one __syncthreads per step, followed by a proxy fence in the issuing thread (generic reads before async writes).
#define MARCH_RB 136     // row buffer: 128 columns + alignment slack of the segments (2 doubles each)
#define MARCH_TMA_SEGS 3  // a CTA's 128 columns span at most 3 rows when a row has >= 64 columns
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(double *smem_dst, const double *gmem_src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int RECON, int RS, int DIR>
__global__ void __launch_bounds__(MARCH_THREADS, AKB_MARCH_MINB)
flux_march_tma_kernel(Geom g, EosParams eos, const double *__restrict__ w0, const double *__restrict__ bcc,
                      const double *__restrict__ bnf, double *__restrict__ flx, double *__restrict__ eyo,
                      double *__restrict__ ezo, int seglen, int nseg) {
  static_assert(DIR == 1 || DIR == 2, "march kernel handles x2 and x3");
  static_assert(RECON >= RECON_PPM4, "the ring exists for the 5-point reconstructions");
  using T = SolverTraits<RS>;
  constexpr bool MHD = T::MHD;
  constexpr bool IDEAL = T::IDEAL;
  constexpr int NV = T::NV;
  constexpr int VS = MARCH_SLOTS * MARCH_RB;
  extern __shared__ double ring_all[];   // NV * MARCH_SLOTS * MARCH_RB (ring) + 7 * MARCH_THREADS (left state) + MARCH_SLOTS barriers; base 16-byte aligned
  unsigned long long *full = reinterpret_cast<unsigned long long *>(ring_all + NV * VS + 7 * MARCH_THREADS);
  const int nc = g.nx1 + 2 * g.ext1;
  const int nt = (DIR == 1) ? g.nx3 + 2 * g.ext3 : g.nx2 + 2 * g.ext2;
  const int m = blockIdx.y / nseg, seg = blockIdx.y - m * nseg;       // block-uniform
  const int ncol = nt * nc;
  const int c0 = blockIdx.x * blockDim.x;                             // first column of the CTA (block-uniform)
  const bool colok = c0 + (int)threadIdx.x < ncol;
  const int col = colok ? c0 + (int)threadIdx.x : ncol - 1;           // idle threads shadow the last column (no stores)
  const int tr = col / nc;
  const int i = g.is - g.ext1 + (col - tr * nc);
  const int fs0 = ((DIR == 1) ? g.js : g.ks) - g.next;
  const int fe0 = ((DIR == 1) ? g.je + 1 : g.ke + 1) + g.next;
  const int f_lo = fs0 + seg * seglen;
  int f_hi = f_lo + seglen - 1;
  if (f_hi > fe0) f_hi = fe0;
  if (f_lo > f_hi) return;                                            // block-uniform
  const long cc = g.cc;
  const int s = (DIR == 1) ? g.n1 : g.n1 * g.n2;
  int k, j;
  if (DIR == 1) { k = g.ks - g.ext3 + tr; j = f_lo - 1; } else { j = g.js - g.ext2 + tr; k = f_lo - 1; }
  int co = (k * g.n2 + j) * g.n1 + i;
  int fo = (DIR == 1) ? (k * (g.n2 + 1) + f_lo) * g.n1 + i : (f_lo * g.n2 + j) * g.n1 + i;
  const long fsz = (DIR == 1) ? g.f2 : g.f3;
  const double *wb = w0 + (long)m * g.nvar * cc;
  const double *bb = MHD ? bcc + (long)m * 3 * cc : wb;
  const double *bnb = MHD ? bnf + (long)m * fsz : wb;
  double *fb = flx + (long)m * g.nvar * fsz;
  double *eyb = MHD ? eyo + (long)m * cc : nullptr;
  double *ezb = MHD ? ezo + (long)m * cc : nullptr;
  constexpr int ivx = IVX + DIR, ivy = IVX + (DIR + 1) % 3, ivz = IVX + (DIR + 2) % 3;
  constexpr int iby = (DIR + 1) % 3, ibz = (DIR + 2) % 3;
  auto vsrc = [&](int vv) -> const double * {
    const int v = T::member(vv);
    return v == 0 ? wb + IDN * cc : v == 1 ? wb + ivx * cc : v == 2 ? wb + ivy * cc : v == 3 ? wb + ivz * cc
         : v == 4 ? wb + IEN * cc : v == 5 ? bb + iby * cc : bb + ibz * cc;
  };
  // the CTA's row segments (block-uniform): its 128 flattened columns cover at most MARCH_TMA_SEGS consecutive transverse rows
  // (the launcher takes this kernel only when a row has at least 64 columns); segment q = nq columns of row trA+q starting at
  // x1 index iq, fetched from the even-aligned index below it over an even number of doubles into offset doff[q] of the row buffer
  const int trA = c0 / nc, cA = c0 - trA * nc;
  int segn[MARCH_TMA_SEGS], seglen_[MARCH_TMA_SEGS], segoff[MARCH_TMA_SEGS], segrow[MARCH_TMA_SEGS], segpad[MARCH_TMA_SEGS];
  int left = min((int)blockDim.x, ncol - c0), doff = 0, tot_len = 0;
#pragma unroll
  for (int q = 0; q < MARCH_TMA_SEGS; ++q) {
    const int cq = (q == 0) ? cA : 0;
    const int nq = max(0, min(left, nc - cq));
    const int iq = g.is - g.ext1 + cq, iq_al = iq & ~1;
    const int lq = nq > 0 ? (((iq - iq_al) + nq + 1) & ~1) : 0;
    segn[q] = nq; seglen_[q] = lq; segoff[q] = doff; segpad[q] = iq - iq_al;
    // cell offset of the aligned start at sweep index f_lo-1
    segrow[q] = (DIR == 1) ? ((g.ks - g.ext3 + trA + q) * g.n2 + (f_lo - 1)) * g.n1 + iq_al
                           : ((f_lo - 1) * g.n2 + (g.js - g.ext2 + trA + q)) * g.n1 + iq_al;
    left -= nq; doff += lq; tot_len += lq;
  }
  const unsigned row_bytes = (unsigned)(NV * tot_len * 8);
  int sidx = segpad[0];                  // idle threads shadow the CTA's first column
  {
    int t = (int)threadIdx.x;
#pragma unroll
    for (int q = 0; q < MARCH_TMA_SEGS; ++q) {
      if (colok && t >= 0 && t < segn[q]) sidx = segoff[q] + segpad[q] + t;
      t -= segn[q];
    }
  }
  auto fetch_row = [&](int slot, int r) {       // row (f_lo-1) + r of every variable into `slot`; one thread
    mbar_expect_tx(full + slot, row_bytes);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double *dst = ring_all + v * VS + slot * MARCH_RB;
      const double *src = vsrc(v) + r * s;
#pragma unroll
      for (int q = 0; q < MARCH_TMA_SEGS; ++q)
        if (seglen_[q]) bulk_g2s(dst + segoff[q], src + segrow[q], (unsigned)(seglen_[q] * 8), full + slot);
    }
  };
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < MARCH_SLOTS; ++q) mbar_init(full + q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned phase = 0;                  // bit q = parity the next wait on slot q expects
  // prologue: rows c-2 .. c+2+PF -> slots 0..S-1 (row c-2 in slot 0), c = f_lo-1
  if (threadIdx.x == 0) {
#pragma unroll
    for (int r = 0; r < MARCH_SLOTS; ++r)
      if (r < 5 || (f_lo - 1) + (r - 2) <= f_hi + 2) fetch_row(r, r - 2);
  }
  const double *ring = ring_all + sidx;
#pragma unroll
  for (int r = 0; r < 5; ++r) { mbar_wait(full + r, (phase >> r) & 1u); phase ^= 1u << r; }
  State7 L, lo_unused;
  {
    bool oor0 = false;
    recon_ring_cell<RECON, RS, MARCH_RB>(ring, 0, eos, L, lo_unused, &oor0);
    if (oor0) { State7 lg; march_recon_guarded<RECON, RS, MARCH_RB>(ring, 0, eos, lg); L = lg; }
  }
  double *lsm = ring_all + NV * VS + threadIdx.x;
  store_left_state(lsm, L);
  int b0 = 1;
  double bn_next = 0.0;
  if constexpr (MHD) bn_next = __ldg(bnb + fo);
  int r_next = 3 + MARCH_PF;           // next row to fetch, relative to sweep index f_lo-1: c+2+PF for c = f_lo
  for (int f = f_lo; f <= f_hi; ++f) {
    co += s;
    const double bni = bn_next;
    if constexpr (MHD) { if (f < f_hi) bn_next = __ldg(bnb + fo + s); }
    int bn_slot = b0 + MARCH_SLOTS - 1; if (bn_slot >= MARCH_SLOTS) bn_slot -= MARCH_SLOTS;     // held row c-3, last read one step ago
    __syncthreads();                   // every thread is done with that slot
    if (threadIdx.x == 0 && f + MARCH_PF <= f_hi) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      fetch_row(bn_slot, r_next);
    }
    r_next += 1;
    // this step reads rows c-2..c+2: the newest of them, c+2, sits in the slot 4 after b0
    int b4 = b0 + 4; if (b4 >= MARCH_SLOTS) b4 -= MARCH_SLOTS;
    mbar_wait(full + b4, (phase >> b4) & 1u);
    phase ^= 1u << b4;
    State7 up, lo;
    bool oor = false;
    recon_ring_cell<RECON, RS, MARCH_RB>(ring, b0, eos, up, lo, &oor);
    Flux7 fx;
    riemann_any<RS>(L, lo, bni, eos, fx, &oor);
    if (oor) {
      State7 ug; Flux7 fg;
      march_face_guarded<RECON, RS, MARCH_RB>(ring, b0, lsm, bni, eos, ug, fg);
      up = ug; fx = fg;
    }
    store_left_state(lsm, up);
    if (colok) {
      double *fp = fb + fo;
      fp[IDN * fsz] = fx.d;
      fp[ivx * fsz] = fx.mx;
      fp[ivy * fsz] = fx.my;
      fp[ivz * fsz] = fx.mz;
      if constexpr (IDEAL) fp[IEN * fsz] = fx.e;
      if constexpr (MHD) { eyb[co] = fx.ey; ezb[co] = fx.ez; }
    }
    L = up;
    fo += s;
    b0 = (b0 + 1 >= MARCH_SLOTS) ? 0 : b0 + 1;
  }
}

// ---------------------------------------------------------------------------------------
// Passive-scalar fluxes (mhd_fluxes.cpp:153-166,203-216,250-263; hydro_fluxes.cpp:126-138,...):
// flx(n) = flx(IDN) * (flx(IDN) >= 0 ? wl(n) : wr(n)) over the faces of the ACTIVE transverse
// range.  One thread per (face, scalar); only the upwind side is reconstructed (the value is the
// reference's wl or wr entry bit for bit).  Runs after the sweep that wrote the mass flux.
template <int RECON, int DIR>
__global__ void __launch_bounds__(256)
scalar_flux_kernel(Geom g, const double *__restrict__ w0, double *__restrict__ flx) {
  const int nfi = g.nx1 + (DIR == 0), nfj = g.nx2 + (DIR == 1), nfk = g.nx3 + (DIR == 2);
  const int ns = g.nvar - g.nfl;
  const long total = (long)g.nmb * ns * nfk * nfj * nfi;
  const long t0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t0 >= total) return;
  long t = t0;
  const int i = g.is + (int)(t % nfi); t /= nfi;
  const int j = g.js + (int)(t % nfj); t /= nfj;
  const int k = g.ks + (int)(t % nfk); t /= nfk;
  const int n = g.nfl + (int)(t % ns);
  const int m = (int)(t / ns);
  const long s = (DIR == 0) ? 1 : (DIR == 1) ? (long)g.n1 : (long)g.n1 * g.n2;
  const long fsz = (DIR == 0) ? g.f1 : (DIR == 1) ? g.f2 : g.f3;
  const long foff = (DIR == 0) ? ((long)k * g.n2 + j) * (g.n1 + 1) + i
                  : (DIR == 1) ? ((long)k * (g.n2 + 1) + j) * g.n1 + i : ((long)k * g.n2 + j) * g.n1 + i;
  double *fb = flx + (long)m * g.nvar * fsz + foff;
  const double fd = fb[IDN * fsz];
  const double *q = w0 + ((long)m * g.nvar + n) * g.cc + ((long)k * g.n2 + j) * g.n1 + i;   // cell on the upper side of the face
  double up, lo;
  if (fd >= 0.0) { recon_var<RECON>(q - s, s, up, lo); fb[n * fsz] = fd * up; }   // wl = upper-face value of the cell below
  else { recon_var<RECON>(q, s, up, lo); fb[n * fsz] = fd * lo; }                 // wr = lower-face value of this cell
}

template <int RECON>
static cudaError_t launch_scalar_fluxes(const Geom &g, const Arrays &a, cudaStream_t st) {
  const int ns = g.nvar - g.nfl;
  if (ns <= 0) return cudaSuccess;
  {
    const long n = (long)g.nmb * ns * g.nx3 * g.nx2 * (g.nx1 + 1);
    scalar_flux_kernel<RECON, 0><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, a.w0, a.flx1);
  }
  if (g.multi_d) {
    const long n = (long)g.nmb * ns * g.nx3 * (g.nx2 + 1) * g.nx1;
    scalar_flux_kernel<RECON, 1><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, a.w0, a.flx2);
  }
  if (g.three_d) {
    const long n = (long)g.nmb * ns * (g.nx3 + 1) * g.nx2 * g.nx1;
    scalar_flux_kernel<RECON, 2><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, a.w0, a.flx3);
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// launchers

static inline unsigned nblocks(long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// The three sweeps are independent (they read w0/bcc0/b0 and write disjoint arrays), so x2 and x3 go to
// two auxiliary streams forked from and joined back into the caller's stream: the block scheduler
// fills the partial last wave of one sweep with blocks of the next instead of idling those SMs.
// One pool per host thread (a handle is driven by one thread).  AKB_SWEEP_STREAMS=0 disables it.
struct SweepFork {
  cudaStream_t s[2] = {nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
  bool ok = false;
  SweepFork() {
    const char *env = getenv("AKB_SWEEP_STREAMS");
    if (env && env[0] == '0') return;
    ok = cudaStreamCreateWithFlags(&s[0], cudaStreamNonBlocking) == cudaSuccess &&
         cudaStreamCreateWithFlags(&s[1], cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&join[0], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&join[1], cudaEventDisableTiming) == cudaSuccess;
  }
};
// x2/x3 sweeps through flux_march_tma_kernel: process-wide switch, AKB_MARCH_TMA=0/1 at start-up, ak_mhd_set_march_tma at run time
inline int &march_tma_flag() {
  static int flag = [] { const char *v = getenv("AKB_MARCH_TMA"); return v ? (int)(v[0] == '1') : AKB_MARCH_TMA_DEFAULT; }();
  return flag;
}
inline SweepFork *sweep_fork() {
  thread_local SweepFork f;
  return f.ok ? &f : nullptr;
}

// cudaFuncSetAttribute is per device: remember it per (kernel instantiation, device) -- a host that drives two devices from one
// process must raise the dynamic shared-memory limit on each (round-1 VERDICT, minor 11)
static inline bool first_use_on_device(unsigned long long &mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

template <int RECON, int RS, bool ACC = false>
static cudaError_t launch_fluxes_t(const Geom &g, const EosParams &eos, const Arrays &a, cudaStream_t st, const AccArgs *accp = nullptr) {
  constexpr int NV = SolverTraits<RS>::NV;
  // ACC (flux + update fusion): x3 consumes what x1 and x2 stored, so the sweeps run in order on the caller's stream
  SweepFork *sf = (g.multi_d && !ACC) ? sweep_fork() : nullptr;
  cudaStream_t st2 = st, st3 = st;
  if (sf) {
    cudaEventRecord(sf->fork, st);
    cudaStreamWaitEvent(sf->s[0], sf->fork, 0);
    st2 = sf->s[0];
    if (g.three_d) { cudaStreamWaitEvent(sf->s[1], sf->fork, 0); st3 = sf->s[1]; }
  }
  AccArgs acc{};
  if (ACC) acc = *accp;
  // ring + left-state copy (5-point reconstructions), + the previous face's fluxes (ACC) or two slots of the face-normal field (plain PPM4 march)
  const size_t smem = sizeof(double) * (((RECON >= RECON_PPM4) ? (NV * MARCH_SLOTS + 7) : 0) + (ACC ? 5 : 2)) * MARCH_THREADS;
  // TMA form of the marches (flux_march_tma_kernel): rows fetched by bulk copies; needs 16-byte aligned rows (n1 even, aligned bases)
  const size_t smem_tma = sizeof(double) * ((size_t)NV * MARCH_SLOTS * MARCH_RB + 7 * MARCH_THREADS) + sizeof(unsigned long long) * MARCH_SLOTS;
  const bool tma = march_tma_flag() && (g.n1 % 2 == 0) && (g.nx1 + 2 * g.ext1 >= 64) && ((((uintptr_t)a.w0) | ((uintptr_t)a.bcc0)) & 15) == 0 && RECON == RECON_PPM4 && !ACC;
  {  // x1
    constexpr int STEP = X1Step<RECON>::value - (ACC ? 1 : 0);
    const unsigned nc = g.nx1 + 2 + 2 * g.next + (RECON == RECON_PPM4 ? 1 : 0), plane = (unsigned)(g.nx2 + 2 * g.ext2) * nc;
    // rcp = ceil(2^32 / nc): __umulhi(t, rcp) == t / nc for every t < 2^32 / nc
    const unsigned rcp = (unsigned)((0x100000000ull + nc - 1) / nc);
    const long planes = (long)g.nmb * (g.nx3 + 2 * g.ext3);
    if ((unsigned long long)plane * nc >= 0x100000000ull || planes > 65535) return cudaErrorInvalidConfiguration;
    const unsigned warps = (plane + STEP - 1) / STEP;
    const dim3 gr(nblocks((long)warps * 32, 128), (unsigned)planes);
    flux_x1_kernel<RECON, RS, ACC><<<gr, 128, 0, st>>>(g, eos, a.w0, a.bcc0, a.b0x1f, a.flx1, a.e3x1, a.e2x1, rcp, acc);
  }
  if (g.multi_d) {  // x2: (ey,ez) = (e1x2,e3x2)
    const int nf = g.nx2 + (ACC ? 0 : 1) + 2 * g.next;      // ACC: segments count cells, each closed by one more face
    const int nseg = (nf + 31 - (ACC ? 1 : 0)) / (32 - (ACC ? 1 : 0)), seglen = (nf + nseg - 1) / nseg;
    const long cols = (long)(g.nx3 + 2 * g.ext3) * (g.nx1 + 2 * g.ext1);
    if ((long)g.nmb * nseg > 65535 || (long)g.nvar * g.cc >= 0x7fffffffL) return cudaErrorInvalidConfiguration;
    bool done = false;
    if constexpr (RECON == RECON_PPM4 && !ACC && !AKB_MARCH_CARRY) {   // instantiated for the headline reconstruction only (build time)
      if (tma) {
        static unsigned long long attr1t = 0ull;
        if (first_use_on_device(attr1t)) cudaFuncSetAttribute(flux_march_tma_kernel<RECON, RS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma);
        flux_march_tma_kernel<RECON, RS, 1><<<dim3(nblocks(cols, MARCH_THREADS), (unsigned)(g.nmb * nseg)), MARCH_THREADS, smem_tma, st2>>>(g, eos, a.w0, a.bcc0, a.b0x2f, a.flx2, a.e1x2, a.e3x2, seglen, nseg);
        done = true;
      }
    }
    if (!done) {
      static unsigned long long attr1 = 0ull;
      if (smem && first_use_on_device(attr1)) cudaFuncSetAttribute(flux_march_kernel<RECON, RS, 1, ACC ? 1 : 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      flux_march_kernel<RECON, RS, 1, ACC ? 1 : 0><<<dim3(nblocks(cols, MARCH_THREADS), (unsigned)(g.nmb * nseg)), MARCH_THREADS, smem, st2>>>(g, eos, a.w0, a.bcc0, a.b0x2f, a.flx2, a.e1x2, a.e3x2, seglen, nseg, acc);
    }
  }
  if (g.three_d) {  // x3: (ey,ez) = (e2x3,e1x3)
    const int nf = g.nx3 + (ACC ? 0 : 1) + 2 * g.next;
    const int nseg = (nf + 31 - (ACC ? 1 : 0)) / (32 - (ACC ? 1 : 0)), seglen = (nf + nseg - 1) / nseg;
    const long cols = (long)(g.nx2 + 2 * g.ext2) * (g.nx1 + 2 * g.ext1);
    if ((long)g.nmb * nseg > 65535 || (long)g.nvar * g.cc >= 0x7fffffffL) return cudaErrorInvalidConfiguration;
    bool done = false;
    if constexpr (RECON == RECON_PPM4 && !ACC && !AKB_MARCH_CARRY) {   // instantiated for the headline reconstruction only (build time)
      if (tma) {
        static unsigned long long attr2t = 0ull;
        if (first_use_on_device(attr2t)) cudaFuncSetAttribute(flux_march_tma_kernel<RECON, RS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma);
        flux_march_tma_kernel<RECON, RS, 2><<<dim3(nblocks(cols, MARCH_THREADS), (unsigned)(g.nmb * nseg)), MARCH_THREADS, smem_tma, st3>>>(g, eos, a.w0, a.bcc0, a.b0x3f, a.flx3, a.e2x3, a.e1x3, seglen, nseg);
        done = true;
      }
    }
    if (!done) {
      static unsigned long long attr2 = 0ull;
      if (smem && first_use_on_device(attr2)) cudaFuncSetAttribute(flux_march_kernel<RECON, RS, 2, ACC ? 2 : 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      flux_march_kernel<RECON, RS, 2, ACC ? 2 : 0><<<dim3(nblocks(cols, MARCH_THREADS), (unsigned)(g.nmb * nseg)), MARCH_THREADS, smem, st3>>>(g, eos, a.w0, a.bcc0, a.b0x3f, a.flx3, a.e2x3, a.e1x3, seglen, nseg, acc);
    }
  }
  if (sf) {   // join
    cudaEventRecord(sf->join[0], st2);
    cudaStreamWaitEvent(st, sf->join[0], 0);
    if (g.three_d) { cudaEventRecord(sf->join[1], st3); cudaStreamWaitEvent(st, sf->join[1], 0); }
  }
  if (!ACC && g.nvar > g.nfl) return launch_scalar_fluxes<RECON>(g, a, st);
  return cudaGetLastError();
}

}  // namespace akb
