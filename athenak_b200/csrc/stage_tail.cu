// stage_tail.cu -- the second half of an RK stage as ONE kernel: corner EMFs (GS07) -> conserved-variable update ->
// constrained transport -> ConsToPrim (+ new-dt reduction on the last stage), for 3-D ideal / isothermal MHD on a
// same-level mesh.  Reference tasks fused here: MHD::CornerE (mhd_corner_e.cpp:206-215,309-318,338-414),
// MHD::RKUpdate (mhd_update.cpp:24-84), MHD::CT (mhd_ct.cpp:23-80), MHD::ConToPrim (eos/ideal_mhd.cpp:30-134,
// ideal_c2p_mhd.hpp:20-67, isothermal_mhd.cpp:20-135), MHD::NewTimeStep (mhd_newdt.cpp:31-174) and the stage-1
// CopyCons (mhd_tasks.cpp:162-170).  With the flux sweeps (flux_kernels.cuh) this is the "one or two kernels per stage"
// of the north star; what stays outside is what needs OTHER blocks' data: the EMF flux correction on block surfaces,
// the ghost-zone exchange, and therefore the outermost layer of cells and faces (shell kernels below).
//
// Mapping.  A CTA owns an (i,j) tile of TI x TJ cells of one MeshBlock and MARCHES along k; thread (ty,tx) sits on cell
// (j0-1+ty, i0-1+tx), so the thread tile is the cell tile plus a one-cell frame (W = TI+2, H = TJ+2).  Per plane p:
//   A  every thread loads the 15 plane-p values the corner EMFs need (v, Bcc, six face EMFs, three mass fluxes), forms the
//      cell-centred E = -v x B and shares what its i+1 / j+1 neighbours need through shared memory;
//   B  threads off the low frame compute the three corner EMFs of their lower-left corner (expressions of
//      corner_emf_kernel, plane p-1 operands carried in registers / the previous shared-memory plane); they stay ON CHIP
//      except within two edges of the block surface, where SendE/RecvE and the shell CT kernel need them in HBM;
//   C  CT: x3f(p) is complete; x1f / x2f of plane p get their E3 term now and their E2 / E1 k-difference at p+1, in the
//      reference's operation order.  Frame threads on the high side compute the tile's last x1f / x2f redundantly so
//      that every owned cell finds both of its faces in shared memory;
//   D  cell p-1: conserved update from the 15 flux values, then (interior cells) face-averaged B, ConsToPrim with floors
//      and counters, new-dt partial minima, and the stores.
// The face field and the primitives are written OUT OF PLACE (b_out, w_out, bcc_out): other tiles read the old values
// (redundant frame faces, cell-centred E of the frame) while this one writes.  The host rotates the buffers
// (host_driver.cpp); at stage 1 the old b0 buffer simply BECOMES b1, which removes CopyCons' face-field copy.
// Work = (tile, k) pairs flattened and cut into equal contiguous runs, one per resident CTA (persistent, 148 SMs x
// resident CTAs); a run that starts inside a tile pays one extra plane of loads.
//
// HBM traffic per cell per stage (FP64 words): reads u0 5 (+u1 5 after stage 1), b0 3 (+b1 3), fluxes 15, face EMFs 6,
// v 3, Bcc 3; writes u0 5 (+u1 5 at stage 1), b0 3, w0 5, bcc0 3: 51 words at stage 1 (56 with the u1 save), 59 after,
// against 81 for corner_emf + update_ct x2 + cons2prim as separate kernels.
#include "mhd_kernels.h"
#include "mhd_device.cuh"
#include "iso_device.cuh"
#include "reduce.cuh"
#include "c2p_device.cuh"

namespace akb {

#ifndef AKB_TAIL_W
#define AKB_TAIL_W 34
#endif
#ifndef AKB_TAIL_H
#define AKB_TAIL_H 10
#endif
#ifndef AKB_TAIL_MINB
#define AKB_TAIL_MINB 2
#endif
// whole warps: the W*H thread tile rounded up to a multiple of 32 (the extra threads only take part in barriers and shuffles)
#define AKB_TAIL_THREADS (((AKB_TAIL_W * AKB_TAIL_H + 31) / 32) * 32)
#ifndef AKB_CELL_MINB
#define AKB_CELL_MINB 3
#endif
#ifndef AKB_TAIL_MINB_FIELD
#define AKB_TAIL_MINB_FIELD 3
#endif

// CELL = false leaves phase D (conserved update + ConsToPrim) to stage_cell_kernel below: the "split" form of the tail.
template <bool STAGE1, bool IDEAL, bool NEWDT, bool CELL>
__global__ void __launch_bounds__(AKB_TAIL_THREADS, CELL ? AKB_TAIL_MINB : AKB_TAIL_MINB_FIELD)
stage_tail_kernel(Geom g, EosParams eos, TailArrays A, double gam0, double gam1, double beta_dt, int ntj, int nti,
                  int kchunk, long nitems, int *__restrict__ counters, unsigned long long *__restrict__ dtmin) {
  constexpr int W = AKB_TAIL_W, H = AKB_TAIL_H, NT = AKB_TAIL_THREADS, TI = W - 2, TJ = H - 2;   // NT: shared-memory plane stride
  extern __shared__ double sm[];
  double *S_c1 = sm, *S_c2 = sm + 2 * NT, *S_c3 = sm + 4 * NT, *S_e1x3 = sm + 5 * NT, *S_e2x3 = sm + 6 * NT;
  double *S_e3x2 = sm + 7 * NT, *S_e3x1 = sm + 8 * NT, *S_F1 = sm + 9 * NT, *S_F2 = sm + 10 * NT, *S_F3 = sm + 11 * NT;
  double *S_E1 = sm + 12 * NT, *S_E2 = sm + 13 * NT, *S_E3 = sm + 14 * NT, *S_x1n = sm + 15 * NT, *S_x2n = sm + 16 * NT;
  const int t = threadIdx.x, ty = t / W, tx = t - ty * W;
  const long n1 = g.n1, n2 = g.n2, n12 = n1 * n2, cc = g.cc;
  const int nx3 = g.nx3;
  int nd = 0, ne = 0, ntf = 0;
  double dmin1 = DT_IDENT, dmin2 = DT_IDENT, dmin3 = DT_IDENT;

  // Work items = (k-chunk, tile), chunk-major, dealt round-robin to the resident CTAs: CTAs that run side by side work on
  // neighbouring tiles of the SAME k-range, so the one-cell frames they share (and the plane a chunk re-reads from its
  // predecessor) are L2 hits instead of second trips to HBM.
  const long ntiles = (long)g.nmb * ntj * nti;
  for (long item = blockIdx.x; item < nitems; item += gridDim.x) {
    const long chunk = item / ntiles;
    const long tile = item - chunk * ntiles;
    const int kc = (int)chunk * kchunk;
    const int kcount = (nx3 - kc < kchunk) ? (nx3 - kc) : kchunk;
    const int ka = g.ks + kc, kb = ka + kcount - 1;
    const int m = (int)(tile / (ntj * nti));
    const int tr = (int)(tile - (long)m * ntj * nti);
    const int tj = tr / nti, ti = tr - tj * nti;
    const int j = g.js + tj * TJ - 1 + ty, i = g.is + ti * TI - 1 + tx;
    const bool inb = (t < W * H) && (j <= g.je + 1) && (i <= g.ie + 1);         // a thread of the tile, inside the corner range (frame included)
    const bool corner = inb && ty >= 1 && tx >= 1;
    const bool own = ty >= 1 && ty <= H - 2 && tx >= 1 && tx <= W - 2 && j <= g.je && i <= g.ie;   // an active cell of this tile
    const bool iin = (i >= g.is + 1 && i <= g.ie - 1), jin = (j >= g.js + 1 && j <= g.je - 1);
    // faces this thread evaluates (owned ones + the redundant high-frame column / row), interior ones only
    const bool f1t = ty >= 1 && ty <= H - 2 && tx >= 1 && jin && (i >= g.is + 1 && i <= g.ie);
    const bool f2t = tx >= 1 && tx <= W - 2 && ty >= 1 && ty <= H - 1 && iin && (j >= g.js + 1 && j <= g.je);
    const bool f3t = own && iin && jin;
    const double dx1 = __ldg(A.mb_size + m * 9 + 6), dx2 = __ldg(A.mb_size + m * 9 + 7), dx3 = __ldg(A.mb_size + m * 9 + 8);
    const double rdx1 = drcp(dx1), rdx2 = drcp(dx2), rdx3 = drcp(dx3);
    const double *wv = A.w_in + (long)m * g.nvar * cc, *bv = A.bcc_in + (long)m * 3 * cc;
    const long mo = (long)m * cc;
    const double *f1d = A.flx[0] + (long)m * g.nvar * g.f1, *f2d = A.flx[1] + (long)m * g.nvar * g.f2, *f3d = A.flx[2] + (long)m * g.nvar * g.f3;
    const long rc = (long)j * n1 + i;                       // in-plane offsets: cells, x1-faces, x2-faces
    const long r1 = (long)j * (n1 + 1) + i;
    // carried from plane p-1
    double p_e1x2 = 0.0, p_e2x1 = 0.0, pF1 = 0.0, pF2 = 0.0, pE1 = 0.0, pE2 = 0.0, x1part = 0.0, x2part = 0.0, x3prev = 0.0;

    __syncthreads();
    for (int p = ka - 1; p <= kb + 1; ++p) {
      const int cur = p & 1, prv = cur ^ 1;
      const bool pro = (p == ka - 1);                      // prologue plane: raw values only
      // ---- A: plane-p operands of the corner EMFs ------------------------------------------------------------------
      double a_e3x1 = 0.0, a_e2x1 = 0.0, a_e1x2 = 0.0, a_e3x2 = 0.0, a_e2x3 = 0.0, a_e1x3 = 0.0, F1 = 0.0, F2 = 0.0, F3 = 0.0;
      double c1 = 0.0, c2 = 0.0, c3 = 0.0;
      if (inb) {
        const long o = (long)p * n12 + rc;
        const double vx = __ldg(wv + IVX * cc + o), vy = __ldg(wv + IVY * cc + o), vz = __ldg(wv + IVZ * cc + o);
        const double bx = __ldg(bv + o), by = __ldg(bv + cc + o), bz = __ldg(bv + 2 * cc + o);
        a_e3x1 = __ldg(A.e3x1 + mo + o); a_e2x1 = __ldg(A.e2x1 + mo + o); a_e1x2 = __ldg(A.e1x2 + mo + o);
        a_e3x2 = __ldg(A.e3x2 + mo + o); a_e2x3 = __ldg(A.e2x3 + mo + o); a_e1x3 = __ldg(A.e1x3 + mo + o);
        F1 = __ldg(f1d + (long)p * n2 * (n1 + 1) + r1);
        F2 = __ldg(f2d + (long)p * (n2 + 1) * n1 + rc);
        F3 = __ldg(f3d + o);
        c1 = vz * by - vy * bz;       // cell-centred E = -(v x B), mhd_corner_e.cpp:310-317
        c2 = vx * bz - vz * bx;
        c3 = vy * bx - vx * by;
      }
      S_c1[cur * NT + t] = c1; S_c2[cur * NT + t] = c2;
      if (!pro) {
        S_c3[t] = c3; S_e1x3[t] = a_e1x3; S_e2x3[t] = a_e2x3; S_e3x2[t] = a_e3x2; S_e3x1[t] = a_e3x1;
        S_F1[t] = F1; S_F2[t] = F2; S_F3[t] = F3;
      }
      if (pro) { p_e1x2 = a_e1x2; p_e2x1 = a_e2x1; pF1 = F1; pF2 = F2; continue; }   // uniform: every thread sees the same p
      __syncthreads();
      // ---- B: corner EMFs of (p, j, i), mhd_corner_e.cpp:342-413 in the branch-free form of corner_emf_kernel ---------
      double E1 = 0.0, E2 = 0.0, E3 = 0.0;
      if (corner) {
        const int tjm = t - W, tim = t - 1, tjim = t - W - 1;
        {
          const double x3_jm = S_e1x3[tjm], x3_j = a_e1x3, x2_km = p_e1x2, x2_k = a_e1x2;
          const double c_mm = S_c1[prv * NT + tjm], c_m0 = S_c1[prv * NT + t], c_0m = S_c1[cur * NT + tjm], c_00 = c1;
          const double l3 = (pF2 >= 0.0) ? (x3_jm - c_mm) : (x3_j - c_m0);
          const double r3 = (F2 >= 0.0) ? (x3_jm - c_0m) : (x3_j - c_00);
          const double l2 = (S_F3[tjm] >= 0.0) ? (x2_km - c_mm) : (x2_k - c_0m);
          const double r2 = (F3 >= 0.0) ? (x2_km - c_m0) : (x2_k - c_00);
          E1 = 0.25 * (l3 + r3 + l2 + r2 + x2_km + x2_k + x3_jm + x3_j);
        }
        {
          const double x3_im = S_e2x3[tim], x3_i = a_e2x3, x1_km = p_e2x1, x1_k = a_e2x1;
          const double c_mm = S_c2[prv * NT + tim], c_m0 = S_c2[prv * NT + t], c_0m = S_c2[cur * NT + tim], c_00 = c2;
          const double l3 = (pF1 >= 0.0) ? (x3_im - c_mm) : (x3_i - c_m0);
          const double r3 = (F1 >= 0.0) ? (x3_im - c_0m) : (x3_i - c_00);
          const double l1 = (S_F3[tim] >= 0.0) ? (x1_km - c_mm) : (x1_k - c_0m);
          const double r1_ = (F3 >= 0.0) ? (x1_km - c_m0) : (x1_k - c_00);
          E2 = 0.25 * (l3 + r3 + l1 + r1_ + x3_im + x3_i + x1_km + x1_k);
        }
        {
          const double x2_im = S_e3x2[tim], x2_i = a_e3x2, x1_jm = S_e3x1[tjm], x1_j = a_e3x1;
          const double c_mm = S_c3[tjim], c_m0 = S_c3[tjm], c_0m = S_c3[tim], c_00 = c3;
          const double l2 = (S_F1[tjm] >= 0.0) ? (x2_im - c_mm) : (x2_i - c_m0);
          const double r2 = (F1 >= 0.0) ? (x2_im - c_0m) : (x2_i - c_00);
          const double l1 = (S_F2[tim] >= 0.0) ? (x1_jm - c_mm) : (x1_j - c_0m);
          const double r1_ = (F2 >= 0.0) ? (x1_jm - c_m0) : (x1_j - c_00);
          E3 = 0.25 * (l1 + r1_ + l2 + r2 + x2_im + x2_i + x1_jm + x1_j);
        }
        // the block surface and the layer next to it go to HBM: SendE/RecvE corrects the surface, the shell CT kernel
        // needs both (identical bits where two tiles or two k-runs evaluate the same corner)
        const bool near = (i <= g.is + 1) || (i >= g.ie) || (j <= g.js + 1) || (j >= g.je) || (p <= g.ks + 1) || (p >= g.ke);
        if (near) {
          if (i <= g.ie) A.e1[(long)m * g.e1 + ((long)p * (n2 + 1) + j) * n1 + i] = E1;
          if (j <= g.je) A.e2[(long)m * g.e2 + ((long)p * n2 + j) * (n1 + 1) + i] = E2;
          if (p <= g.ke) A.e3[(long)m * g.e3 + ((long)p * (n2 + 1) + j) * (n1 + 1) + i] = E3;
        }
      }
      S_E1[t] = E1; S_E2[t] = E2; S_E3[t] = E3;
      __syncthreads();
      // ---- C: constrained transport (mhd_ct.cpp:45-77), interior faces ----------------------------------------------
      const bool kin_f = (p >= g.ks + 1 && p <= g.ke - 1);          // plane of x1f / x2f faces that are interior in k
      double x1new = 0.0, x2new = 0.0, x3new = 0.0;
      {
        // complete the x1f / x2f of plane p-1 with the k-difference of E2 / E1 (third term of the reference expressions)
        const bool kprev = (p - 1 >= g.ks + 1 && p - 1 <= g.ke - 1 && p - 1 >= ka);
        if (f1t && kprev) {
          bool oor = false;
          const double de = beta_dt * (E2 - pE2);
          x1new = x1part + div_by(de, dx3, rdx3, &oor);
          if (oor) x1new = x1part + div_by(de, dx3, rdx3);
          if (tx <= W - 2) A.b_out[0][(long)m * g.f1 + (long)(p - 1) * n2 * (n1 + 1) + r1] = x1new;
        }
        if (f2t && kprev) {
          bool oor = false;
          const double de = beta_dt * (E1 - pE1);
          x2new = x2part - div_by(de, dx3, rdx3, &oor);
          if (oor) x2new = x2part - div_by(de, dx3, rdx3);
          if (ty <= H - 2) A.b_out[1][(long)m * g.f2 + (long)(p - 1) * (n2 + 1) * n1 + rc] = x2new;
        }
        if (CELL) { S_x1n[t] = x1new; S_x2n[t] = x2new; }
        // plane p: E3 terms of x1f / x2f now, all of x3f
        if (p <= kb) {
          if (f1t && kin_f) {
            const long o = (long)m * g.f1 + (long)p * n2 * (n1 + 1) + r1;
            const double bold = __ldg(A.b_in[0] + o);
            const double breg = STAGE1 ? bold : __ldg(A.b1[0] + o);
            const double de = beta_dt * (S_E3[t + W] - E3);
            bool oor = false;
            x1part = gam0 * bold + gam1 * breg;
            double q = div_by(de, dx2, rdx2, &oor);
            if (oor) q = div_by(de, dx2, rdx2);
            x1part -= q;
          }
          if (f2t && kin_f) {
            const long o = (long)m * g.f2 + (long)p * (n2 + 1) * n1 + rc;
            const double bold = __ldg(A.b_in[1] + o);
            const double breg = STAGE1 ? bold : __ldg(A.b1[1] + o);
            const double de = beta_dt * (S_E3[t + 1] - E3);
            bool oor = false;
            x2part = gam0 * bold + gam1 * breg;
            double q = div_by(de, dx1, rdx1, &oor);
            if (oor) q = div_by(de, dx1, rdx1);
            x2part += q;
          }
        }
        if (f3t && p >= g.ks + 1 && p <= g.ke) {
          const long o = (long)m * g.f3 + (long)p * n12 + rc;
          const double bold = __ldg(A.b_in[2] + o);
          const double breg = STAGE1 ? bold : __ldg(A.b1[2] + o);
          const double de2 = beta_dt * (S_E2[t + 1] - E2), de1 = beta_dt * (S_E1[t + W] - E1);
          bool oor = false;
          double q2 = div_by(de2, dx1, rdx1, &oor), q1 = div_by(de1, dx2, rdx2, &oor);
          if (oor) { q2 = div_by(de2, dx1, rdx1); q1 = div_by(de1, dx2, rdx2); }
          x3new = gam0 * bold + gam1 * breg;
          x3new -= q2;
          x3new += q1;
          if (p <= kb) A.b_out[2][o] = x3new;
        }
      }
      if (CELL) __syncthreads();
      // ---- D: cell (p-1, j, i): conserved update (mhd_update.cpp:51-82), then c2p for interior cells -------------------
      const int kcell = p - 1;
      if (CELL && own && kcell >= ka) {
        const long o = (long)kcell * n12 + rc;
        const long o1 = (long)kcell * n2 * (n1 + 1) + r1;
        const long o2 = (long)kcell * (n2 + 1) * n1 + rc;
        double un[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        auto update_var = [&](int n, bool *oor) -> double {
          const double *q1 = A.flx[0] + ((long)m * g.nvar + n) * g.f1 + o1;
          const double *q2 = A.flx[1] + ((long)m * g.nvar + n) * g.f2 + o2;
          const double *q3 = A.flx[2] + ((long)m * g.nvar + n) * g.f3 + o;
          double divf = div_by(__ldg(q1 + 1) - __ldg(q1), dx1, rdx1, oor);
          divf += div_by(__ldg(q2 + n1) - __ldg(q2), dx2, rdx2, oor);
          divf += div_by(__ldg(q3 + n12) - __ldg(q3), dx3, rdx3, oor);
          const long uo = ((long)m * g.nvar + n) * cc + o;
          const double uold = __ldg(A.u_in + uo);
          double ureg;
          if (STAGE1) { ureg = uold; if (A.u1w) A.u1w[uo] = uold; } else { ureg = __ldg(A.u1r + uo); }
          return gam0 * uold + gam1 * ureg - beta_dt * divf;
        };
        {
          bool oor = false;
#pragma unroll
          for (int n = 0; n < 4; ++n) un[n] = update_var(n, &oor);
          if (IDEAL) un[4] = update_var(IEN, &oor);
          if (oor) {
#pragma unroll
            for (int n = 0; n < 4; ++n) un[n] = update_var(n, nullptr);
            if (IDEAL) un[4] = update_var(IEN, nullptr);
          }
        }
        double *uo_ = A.u_out + (long)m * g.nvar * cc + o;
        const bool cint = iin && jin && (kcell >= g.ks + 1 && kcell <= g.ke - 1);
        if (!cint) {        // outermost layer: its faces wait for the corrected EMFs; the shell kernels finish it
          uo_[IDN * cc] = un[0]; uo_[IM1 * cc] = un[1]; uo_[IM2 * cc] = un[2]; uo_[IM3 * cc] = un[3];
          if (IDEAL) uo_[IEN * cc] = un[4];
          for (int n = g.nfl; n < g.nvar; ++n) uo_[n * cc] = update_var(n, nullptr);
        } else {
          const double ubx = 0.5 * (x1new + S_x1n[t + 1]);
          const double uby = 0.5 * (x2new + S_x2n[t + W]);
          const double ubz = 0.5 * (x3prev + x3new);
          double ud = un[0], ue = un[4], wd, wvx, wvy, wvz, we;
          bool dfl = false, efl = false, tfl = false, oor = false;
          c2p_cell<IDEAL>(eos, ud, un[1], un[2], un[3], ue, ubx, uby, ubz, wd, wvx, wvy, wvz, we, dfl, efl, tfl, &oor);
          if (oor) {
            ud = un[0]; ue = un[4]; dfl = efl = tfl = false;
            c2p_cell<IDEAL>(eos, ud, un[1], un[2], un[3], ue, ubx, uby, ubz, wd, wvx, wvy, wvz, we, dfl, efl, tfl, nullptr);
          }
          nd += dfl; ne += efl; ntf += tfl;
          uo_[IDN * cc] = ud; uo_[IM1 * cc] = un[1]; uo_[IM2 * cc] = un[2]; uo_[IM3 * cc] = un[3];
          double *w = A.w_out + (long)m * g.nvar * cc + o;
          if (IDEAL) { uo_[IEN * cc] = ue; w[IEN * cc] = we; }
          w[IDN * cc] = wd; w[IVX * cc] = wvx; w[IVY * cc] = wvy; w[IVZ * cc] = wvz;
          double *bc = A.bcc_out + (long)m * 3 * cc + o;
          bc[0] = ubx; bc[cc] = uby; bc[2 * cc] = ubz;
          for (int n = g.nfl; n < g.nvar; ++n) {     // passive scalars: update, floor (ideal gas only), s / d
            double sc = update_var(n, nullptr);
            if (IDEAL && sc < 0.0) sc = 0.0;
            uo_[n * cc] = sc;
            w[n * cc] = sc / ud;
          }
          if (NEWDT) {
            double d1, d2, d3;
            bool o2_ = false;
            newdt_cell<IDEAL>(eos, dx1, dx2, dx3, wd, wvx, wvy, wvz, we, ubx, uby, ubz, d1, d2, d3, &o2_);
            if (o2_) newdt_cell<IDEAL>(eos, dx1, dx2, dx3, wd, wvx, wvy, wvz, we, ubx, uby, ubz, d1, d2, d3, nullptr);
            dmin1 = dmin(dmin1, d1); dmin2 = dmin(dmin2, d2); dmin3 = dmin(dmin3, d3);
          }
        }
      }
      // carry
      p_e1x2 = a_e1x2; p_e2x1 = a_e2x1; pF1 = F1; pF2 = F2; pE1 = E1; pE2 = E2; x3prev = x3new;
    }
  }
  warp_count3(nd, ne, ntf, counters);
  if (NEWDT) block_min3(dmin1, dmin2, dmin3, dtmin);
}

// ---------------------------------------------------------------------------------------------------------------------
// Split form of the tail, second kernel: phase D as a streaming kernel of its own.  One thread per ACTIVE cell: conserved
// update from the fluxes (mhd_update.cpp:51-82, + the stage-1 CopyCons of u), and for cells off the block surface the
// face-averaged field from the NEW interior faces the field kernel (stage_tail_kernel<..., CELL = false>) just wrote,
// ConsToPrim with floors and counters, new-dt minima.  Reads 31 words per cell (15 flux, u 5 + 5, 6 faces through L1),
// writes 13 (u 5, w 5, bcc 3).
template <bool STAGE1, bool IDEAL, bool NEWDT>
__global__ void __launch_bounds__(256, AKB_CELL_MINB)
stage_cell_kernel(Geom g, EosParams eos, TailArrays A, double gam0, double gam1, double beta_dt, int *__restrict__ counters,
                  unsigned long long *__restrict__ dtmin) {
  const int nkb = (g.nx3 + 1) / 2;
  // 32 x 4 x 2 cell tile per block, as a 1-D thread block (the reductions at the end index warps by threadIdx.x)
  const int ii = blockIdx.x * 32 + (threadIdx.x & 31), jj = blockIdx.y * 4 + ((threadIdx.x >> 5) & 3);
  const int kk = (blockIdx.z % nkb) * 2 + (threadIdx.x >> 7);
  const int m = blockIdx.z / nkb;
  int nd = 0, ne = 0, ntf = 0;
  double dm1 = DT_IDENT, dm2 = DT_IDENT, dm3 = DT_IDENT;
  if (ii < g.nx1 && jj < g.nx2 && kk < g.nx3) {
    const int i = g.is + ii, j = g.js + jj, k = g.ks + kk;
    const long n1 = g.n1, n2 = g.n2, n12 = n1 * n2, cc = g.cc;
    const double dx1 = __ldg(A.mb_size + m * 9 + 6), dx2 = __ldg(A.mb_size + m * 9 + 7), dx3 = __ldg(A.mb_size + m * 9 + 8);
    const double rdx1 = drcp(dx1), rdx2 = drcp(dx2), rdx3 = drcp(dx3);
    const long o = ((long)k * n2 + j) * n1 + i;
    const long o1 = ((long)k * n2 + j) * (n1 + 1) + i;
    const long o2 = ((long)k * (n2 + 1) + j) * n1 + i;
    auto update_var = [&](int n, bool *oor) -> double {
      const double *q1 = A.flx[0] + ((long)m * g.nvar + n) * g.f1 + o1;
      const double *q2 = A.flx[1] + ((long)m * g.nvar + n) * g.f2 + o2;
      const double *q3 = A.flx[2] + ((long)m * g.nvar + n) * g.f3 + o;
      double divf = div_by(__ldg(q1 + 1) - __ldg(q1), dx1, rdx1, oor);
      divf += div_by(__ldg(q2 + n1) - __ldg(q2), dx2, rdx2, oor);
      divf += div_by(__ldg(q3 + n12) - __ldg(q3), dx3, rdx3, oor);
      const long uo = ((long)m * g.nvar + n) * cc + o;
      const double uold = __ldg(A.u_in + uo);
      double ureg;
      if (STAGE1) { ureg = uold; if (A.u1w) A.u1w[uo] = uold; } else { ureg = __ldg(A.u1r + uo); }
      return gam0 * uold + gam1 * ureg - beta_dt * divf;
    };
    double un[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    {
      bool oor = false;
#pragma unroll
      for (int n = 0; n < 4; ++n) un[n] = update_var(n, &oor);
      if (IDEAL) un[4] = update_var(IEN, &oor);
      if (oor) {
#pragma unroll
        for (int n = 0; n < 4; ++n) un[n] = update_var(n, nullptr);
        if (IDEAL) un[4] = update_var(IEN, nullptr);
      }
    }
    double *uo_ = A.u_out + (long)m * g.nvar * cc + o;
    const bool cint = (i >= g.is + 1 && i <= g.ie - 1) && (j >= g.js + 1 && j <= g.je - 1) && (k >= g.ks + 1 && k <= g.ke - 1);
    if (!cint) {
      uo_[IDN * cc] = un[0]; uo_[IM1 * cc] = un[1]; uo_[IM2 * cc] = un[2]; uo_[IM3 * cc] = un[3];
      if (IDEAL) uo_[IEN * cc] = un[4];
      for (int n = g.nfl; n < g.nvar; ++n) uo_[n * cc] = update_var(n, nullptr);
    } else {
      const double *p1 = A.b_out[0] + (long)m * g.f1 + o1, *p2 = A.b_out[1] + (long)m * g.f2 + o2, *p3 = A.b_out[2] + (long)m * g.f3 + o;
      const double ubx = 0.5 * (p1[0] + p1[1]);      // plain loads: written by the previous kernel of this stream
      const double uby = 0.5 * (p2[0] + p2[n1]);
      const double ubz = 0.5 * (p3[0] + p3[n12]);
      double ud = un[0], ue = un[4], wd, wvx, wvy, wvz, we;
      bool dfl = false, efl = false, tfl = false, oor = false;
      c2p_cell<IDEAL>(eos, ud, un[1], un[2], un[3], ue, ubx, uby, ubz, wd, wvx, wvy, wvz, we, dfl, efl, tfl, &oor);
      if (oor) {
        ud = un[0]; ue = un[4]; dfl = efl = tfl = false;
        c2p_cell<IDEAL>(eos, ud, un[1], un[2], un[3], ue, ubx, uby, ubz, wd, wvx, wvy, wvz, we, dfl, efl, tfl, nullptr);
      }
      nd += dfl; ne += efl; ntf += tfl;
      uo_[IDN * cc] = ud; uo_[IM1 * cc] = un[1]; uo_[IM2 * cc] = un[2]; uo_[IM3 * cc] = un[3];
      double *w = A.w_out + (long)m * g.nvar * cc + o;
      if (IDEAL) { uo_[IEN * cc] = ue; w[IEN * cc] = we; }
      w[IDN * cc] = wd; w[IVX * cc] = wvx; w[IVY * cc] = wvy; w[IVZ * cc] = wvz;
      double *bc = A.bcc_out + (long)m * 3 * cc + o;
      bc[0] = ubx; bc[cc] = uby; bc[2 * cc] = ubz;
      for (int n = g.nfl; n < g.nvar; ++n) {
        double sc = update_var(n, nullptr);
        if (IDEAL && sc < 0.0) sc = 0.0;
        uo_[n * cc] = sc;
        w[n * cc] = sc / ud;
      }
      if (NEWDT) {
        bool o2_ = false;
        newdt_cell<IDEAL>(eos, dx1, dx2, dx3, wd, wvx, wvy, wvz, we, ubx, uby, ubz, dm1, dm2, dm3, &o2_);
        if (o2_) newdt_cell<IDEAL>(eos, dx1, dx2, dx3, wd, wvx, wvy, wvz, we, ubx, uby, ubz, dm1, dm2, dm3, nullptr);
      }
    }
  }
  warp_count3(nd, ne, ntf, counters);
  if (NEWDT) block_min3(dm1, dm2, dm3, dtmin);
}

// ---------------------------------------------------------------------------------------------------------------------
// Shell kernels: the outermost layer, after SendE/RecvE corrected the surface EMFs (shell CT) and after the ghost-zone
// exchange (shell c2p).  Both enumerate "a box minus an inner box" as six slabs, one thread per position.
struct ShellBox {   // outer box [lo, hi], inner (excluded) box [ilo, ihi]; i, j, k
  int lo[3], hi[3], ilo[3], ihi[3];
};
__device__ __forceinline__ bool shell_decode(const ShellBox &b, long q, int &i, int &j, int &k) {
  const long ni = b.hi[0] - b.lo[0] + 1, nj = b.hi[1] - b.lo[1] + 1;
  // k-slabs: k in [lo, ilo) and (ihi, hi], all j, i
  const long nk_lo = b.ilo[2] - b.lo[2], nk_hi = b.hi[2] - b.ihi[2];
  long n = (nk_lo + nk_hi) * nj * ni;
  if (q < n) {
    i = b.lo[0] + (int)(q % ni); q /= ni;
    j = b.lo[1] + (int)(q % nj); q /= nj;
    k = (q < nk_lo) ? b.lo[2] + (int)q : b.ihi[2] + 1 + (int)(q - nk_lo);
    return true;
  }
  q -= n;
  const long nkm = b.ihi[2] - b.ilo[2] + 1;       // middle k range
  const long nj_lo = b.ilo[1] - b.lo[1], nj_hi = b.hi[1] - b.ihi[1];
  n = nkm * (nj_lo + nj_hi) * ni;
  if (q < n) {
    i = b.lo[0] + (int)(q % ni); q /= ni;
    const long jj = q % (nj_lo + nj_hi); q /= (nj_lo + nj_hi);
    j = (jj < nj_lo) ? b.lo[1] + (int)jj : b.ihi[1] + 1 + (int)(jj - nj_lo);
    k = b.ilo[2] + (int)q;
    return true;
  }
  q -= n;
  const long njm = b.ihi[1] - b.ilo[1] + 1;
  const long ni_lo = b.ilo[0] - b.lo[0], ni_hi = b.hi[0] - b.ihi[0];
  n = nkm * njm * (ni_lo + ni_hi);
  if (q < n) {
    const long ii = q % (ni_lo + ni_hi); q /= (ni_lo + ni_hi);
    i = (ii < ni_lo) ? b.lo[0] + (int)ii : b.ihi[0] + 1 + (int)(ii - ni_lo);
    j = b.ilo[1] + (int)(q % njm); q /= njm;
    k = b.ilo[2] + (int)q;
    return true;
  }
  return false;
}
static long shell_count(const ShellBox &b) {
  const long ni = b.hi[0] - b.lo[0] + 1, nj = b.hi[1] - b.lo[1] + 1, nk = b.hi[2] - b.lo[2] + 1;
  const long mi = b.ihi[0] - b.ilo[0] + 1, mj = b.ihi[1] - b.ilo[1] + 1, mk = b.ihi[2] - b.ilo[2] + 1;
  return ni * nj * nk - mi * mj * mk;
}

// CT of every active face the tail kernel left out (faces with an edge on the block surface), mhd_ct.cpp:45-77, from the
// corrected edge fields in HBM.  Position (k,j,i) in [s, e+1]^3 minus the deep interior; each thread updates the lower
// x1f / x2f / x3f of its position where that face is active and not an interior one.
template <bool STAGE1>
__global__ void __launch_bounds__(256)
shell_ct_kernel(Geom g, ShellBox box, long per_block, TailArrays A, double gam0, double gam1, double beta_dt) {
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gt >= per_block * g.nmb) return;
  const int m = (int)(gt / per_block);
  int i, j, k;
  if (!shell_decode(box, gt - (long)m * per_block, i, j, k)) return;
  const long n1 = g.n1, n2 = g.n2;
  const double dx1 = __ldg(A.mb_size + m * 9 + 6), dx2 = __ldg(A.mb_size + m * 9 + 7), dx3 = __ldg(A.mb_size + m * 9 + 8);
  const double rdx1 = drcp(dx1), rdx2 = drcp(dx2), rdx3 = drcp(dx3);
  const double *E1 = A.e1 + (long)m * g.e1, *E2 = A.e2 + (long)m * g.e2, *E3 = A.e3 + (long)m * g.e3;
  const long a1 = ((long)k * (n2 + 1) + j) * n1 + i;
  const long a2 = ((long)k * n2 + j) * (n1 + 1) + i;
  const long a3 = ((long)k * (n2 + 1) + j) * (n1 + 1) + i;
  const bool ia = (i <= g.ie), ja = (j <= g.je), ka = (k <= g.ke);
  const bool iin = (i >= g.is + 1 && i <= g.ie - 1), jin = (j >= g.js + 1 && j <= g.je - 1), kin = (k >= g.ks + 1 && k <= g.ke - 1);
  if (ja && ka && !(jin && kin && i >= g.is + 1 && i <= g.ie)) {     // x1f(k,j,i)
    const long o = (long)m * g.f1 + ((long)k * n2 + j) * (n1 + 1) + i;
    const double bold = __ldg(A.b_in[0] + o);
    const double breg = STAGE1 ? bold : __ldg(A.b1[0] + o);
    double b = gam0 * bold + gam1 * breg;
    b -= div_by(beta_dt * (__ldg(E3 + a3 + (n1 + 1)) - __ldg(E3 + a3)), dx2, rdx2);
    b += div_by(beta_dt * (__ldg(E2 + a2 + n2 * (n1 + 1)) - __ldg(E2 + a2)), dx3, rdx3);
    A.b_out[0][o] = b;
  }
  if (ia && ka && !(iin && kin && j >= g.js + 1 && j <= g.je)) {     // x2f(k,j,i)
    const long o = (long)m * g.f2 + ((long)k * (n2 + 1) + j) * n1 + i;
    const double bold = __ldg(A.b_in[1] + o);
    const double breg = STAGE1 ? bold : __ldg(A.b1[1] + o);
    double b = gam0 * bold + gam1 * breg;
    b += div_by(beta_dt * (__ldg(E3 + a3 + 1) - __ldg(E3 + a3)), dx1, rdx1);
    b -= div_by(beta_dt * (__ldg(E1 + a1 + (n2 + 1) * n1) - __ldg(E1 + a1)), dx3, rdx3);
    A.b_out[1][o] = b;
  }
  if (ia && ja && !(iin && jin && k >= g.ks + 1 && k <= g.ke)) {     // x3f(k,j,i)
    const long o = (long)m * g.f3 + ((long)k * n2 + j) * n1 + i;
    const double bold = __ldg(A.b_in[2] + o);
    const double breg = STAGE1 ? bold : __ldg(A.b1[2] + o);
    double b = gam0 * bold + gam1 * breg;
    b -= div_by(beta_dt * (__ldg(E2 + a2 + 1) - __ldg(E2 + a2)), dx1, rdx1);
    b += div_by(beta_dt * (__ldg(E1 + a1 + n1) - __ldg(E1 + a1)), dx2, rdx2);
    A.b_out[2][o] = b;
  }
}

// ConsToPrim of every cell the tail kernel left out: the ghost zones and the outermost active layer.
template <bool IDEAL, bool NEWDT>
__global__ void __launch_bounds__(256)
shell_c2p_kernel(Geom g, EosParams eos, ShellBox box, long per_block, double *__restrict__ u0, const double *__restrict__ b1f,
                 const double *__restrict__ b2f, const double *__restrict__ b3f, double *__restrict__ w0, double *__restrict__ bcc,
                 const double *__restrict__ mb_size, int *__restrict__ counters, unsigned long long *__restrict__ dtmin) {
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  int nd = 0, ne = 0, nt = 0;
  double d1 = DT_IDENT, d2 = DT_IDENT, d3 = DT_IDENT;
  int i, j, k;
  const int m = (int)(gt / per_block);
  if (gt < per_block * g.nmb && shell_decode(box, gt - (long)m * per_block, i, j, k)) {
    const long n1 = g.n1, n2 = g.n2, cc = g.cc;
    const long o = ((long)k * n2 + j) * n1 + i;
    double *u = u0 + (long)m * g.nvar * cc + o;
    double ud = u[IDN * cc], ue = 0.0;
    const double umx = u[IM1 * cc], umy = u[IM2 * cc], umz = u[IM3 * cc];
    if constexpr (IDEAL) ue = u[IEN * cc];
    const double *p1 = b1f + (long)m * g.f1 + ((long)k * n2 + j) * (n1 + 1) + i;
    const double *p2 = b2f + (long)m * g.f2 + ((long)k * (n2 + 1) + j) * n1 + i;
    const double *p3 = b3f + (long)m * g.f3 + o;
    const double ubx = 0.5 * (__ldg(p1) + __ldg(p1 + 1));
    const double uby = 0.5 * (__ldg(p2) + __ldg(p2 + n1));
    const double ubz = 0.5 * (__ldg(p3) + __ldg(p3 + n1 * n2));
    double wd, wvx, wvy, wvz, we;
    bool dfl = false, efl = false, tfl = false;
    c2p_cell<IDEAL>(eos, ud, umx, umy, umz, ue, ubx, uby, ubz, wd, wvx, wvy, wvz, we, dfl, efl, tfl, nullptr);
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
    for (int n = g.nfl; n < g.nvar; ++n) {
      double sc = u[n * cc];
      if (IDEAL && sc < 0.0) { sc = 0.0; u[n * cc] = 0.0; }
      w[n * cc] = sc / ud;
    }
    if (NEWDT && i >= g.is && i <= g.ie && j >= g.js && j <= g.je && k >= g.ks && k <= g.ke)
      newdt_cell<IDEAL>(eos, __ldg(mb_size + m * 9 + 6), __ldg(mb_size + m * 9 + 7), __ldg(mb_size + m * 9 + 8), wd, wvx, wvy, wvz, we,
                        ubx, uby, ubz, d1, d2, d3, nullptr);
  }
  warp_count3(nd, ne, nt, counters);
  if (NEWDT) block_min3(d1, d2, d3, dtmin);
}

// ---------------------------------------------------------------------------------------------------------------------
// launchers
bool stage_tail_supported(const Geom &g) {
  // 3-D blocks with an interior (nx >= 3) and room for the one-cell frame inside the ghost zones
  return g.three_d && g.multi_d && g.nx1 >= 3 && g.nx2 >= 3 && g.nx3 >= 3 && g.ng >= 2 && g.next == 0 && (g.nfl == 4 || g.nfl == 5);
}

// k-chunk length: as few round-robin rounds of (chunk, tile) items over the resident CTAs as possible, each chunk paying
// about one and a half extra planes (prologue loads + the plane that only completes the last cell)
static int tail_pick_chunk(long ntiles, int nx3, int grid) {
  int best = nx3;
  double bestc = 1e300;
  for (int kc = (nx3 < 8 ? nx3 : 8); kc <= nx3; ++kc) {
    const long nch = (nx3 + kc - 1) / kc, items = ntiles * nch;
    const long rounds = (items + grid - 1) / grid;
    const double cost = (double)rounds * (kc + 1.5);
    if (cost < bestc - 1e-9) { bestc = cost; best = kc; }
  }
  return best;
}

template <bool S1, bool ID, bool ND, bool CELL>
static cudaError_t tail_launch_t(const Geom &g, const EosParams &eos, const TailArrays &A, double gam0, double gam1, double beta_dt,
                                 int *counters, unsigned long long *dtmin, cudaStream_t st) {
  constexpr int W = AKB_TAIL_W, H = AKB_TAIL_H, NT = AKB_TAIL_THREADS;
  const int ntj = (g.nx2 + (H - 2) - 1) / (H - 2), nti = (g.nx1 + (W - 2) - 1) / (W - 2);
  const long ntiles = (long)g.nmb * ntj * nti;
  const size_t smem = sizeof(double) * 17 * NT;
  static int resident = 0;
  if (resident == 0) {
    int dev = 0, nsm = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(stage_tail_kernel<S1, ID, ND, CELL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(stage_tail_kernel<S1, ID, ND, CELL>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stage_tail_kernel<S1, ID, ND, CELL>, NT, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    resident = nsm * per_sm;
  }
  const int kchunk = tail_pick_chunk(ntiles, g.nx3, resident);
  const long nitems = ntiles * ((g.nx3 + kchunk - 1) / kchunk);
  const int grid = (int)(nitems < resident ? nitems : resident);
  stage_tail_kernel<S1, ID, ND, CELL><<<grid, NT, smem, st>>>(g, eos, A, gam0, gam1, beta_dt, ntj, nti, kchunk, nitems, counters, dtmin);
  return cudaGetLastError();
}

template <bool S1, bool ID, bool ND>
static cudaError_t cell_launch_t(const Geom &g, const EosParams &eos, const TailArrays &A, double gam0, double gam1, double beta_dt,
                                 int *counters, unsigned long long *dtmin, cudaStream_t st) {
  const dim3 gr((g.nx1 + 31) / 32, (g.nx2 + 3) / 4, ((g.nx3 + 1) / 2) * g.nmb);
  stage_cell_kernel<S1, ID, ND><<<gr, 256, 0, st>>>(g, eos, A, gam0, gam1, beta_dt, counters, dtmin);
  return cudaGetLastError();
}

// split = false: everything in stage_tail_kernel; split = true: field kernel (corner EMFs + CT) + cell kernel (update + c2p)
cudaError_t launch_stage_tail(const Geom &g, const EosParams &eos, const TailArrays &A, bool stage1, double gam0, double gam1,
                              double beta_dt, int *counters, bool newdt, unsigned long long *dtmin, bool split, cudaStream_t st) {
  const bool ideal = (g.nfl == 5);
  cudaError_t e = cudaSuccess;
#define AKB_TAIL_ARGS g, eos, A, gam0, gam1, beta_dt, counters, dtmin, st
#define AKB_TAIL_DISPATCH(S1, ID, ND)                                                         \
  do {                                                                                        \
    if (!split) e = tail_launch_t<S1, ID, ND, true>(AKB_TAIL_ARGS);                           \
    else {                                                                                    \
      e = tail_launch_t<S1, ID, false, false>(AKB_TAIL_ARGS);                                 \
      if (e == cudaSuccess) e = cell_launch_t<S1, ID, ND>(AKB_TAIL_ARGS);                     \
    }                                                                                         \
  } while (0)
  if (stage1) {
    if (ideal) { if (newdt) AKB_TAIL_DISPATCH(true, true, true); else AKB_TAIL_DISPATCH(true, true, false); }
    else { if (newdt) AKB_TAIL_DISPATCH(true, false, true); else AKB_TAIL_DISPATCH(true, false, false); }
  } else {
    if (ideal) { if (newdt) AKB_TAIL_DISPATCH(false, true, true); else AKB_TAIL_DISPATCH(false, true, false); }
    else { if (newdt) AKB_TAIL_DISPATCH(false, false, true); else AKB_TAIL_DISPATCH(false, false, false); }
  }
#undef AKB_TAIL_DISPATCH
#undef AKB_TAIL_ARGS
  return e;
}

cudaError_t launch_shell_ct(const Geom &g, const TailArrays &A, bool stage1, double gam0, double gam1, double beta_dt, cudaStream_t st) {
  ShellBox b;
  b.lo[0] = g.is; b.lo[1] = g.js; b.lo[2] = g.ks; b.hi[0] = g.ie + 1; b.hi[1] = g.je + 1; b.hi[2] = g.ke + 1;
  b.ilo[0] = g.is + 1; b.ilo[1] = g.js + 1; b.ilo[2] = g.ks + 1; b.ihi[0] = g.ie - 1; b.ihi[1] = g.je - 1; b.ihi[2] = g.ke - 1;
  const long per = shell_count(b), n = per * g.nmb;
  const unsigned nb = (unsigned)((n + 255) / 256);
  if (stage1) shell_ct_kernel<true><<<nb, 256, 0, st>>>(g, b, per, A, gam0, gam1, beta_dt);
  else shell_ct_kernel<false><<<nb, 256, 0, st>>>(g, b, per, A, gam0, gam1, beta_dt);
  return cudaGetLastError();
}

cudaError_t launch_shell_c2p(const Geom &g, const EosParams &eos, const Arrays &a, int *counters, bool newdt, unsigned long long *dtmin,
                             cudaStream_t st) {
  ShellBox b;
  b.lo[0] = 0; b.lo[1] = 0; b.lo[2] = 0; b.hi[0] = g.n1 - 1; b.hi[1] = g.n2 - 1; b.hi[2] = g.n3 - 1;
  b.ilo[0] = g.is + 1; b.ilo[1] = g.js + 1; b.ilo[2] = g.ks + 1; b.ihi[0] = g.ie - 1; b.ihi[1] = g.je - 1; b.ihi[2] = g.ke - 1;
  const long per = shell_count(b), n = per * g.nmb;
  const unsigned nb = (unsigned)((n + 255) / 256);
  const bool ideal = (g.nfl == 5);
#define AKB_SC2P_ARGS g, eos, b, per, a.u0, a.b0x1f, a.b0x2f, a.b0x3f, a.w0, a.bcc0, a.mb_size, counters, dtmin
  if (ideal) { if (newdt) shell_c2p_kernel<true, true><<<nb, 256, 0, st>>>(AKB_SC2P_ARGS); else shell_c2p_kernel<true, false><<<nb, 256, 0, st>>>(AKB_SC2P_ARGS); }
  else { if (newdt) shell_c2p_kernel<false, true><<<nb, 256, 0, st>>>(AKB_SC2P_ARGS); else shell_c2p_kernel<false, false><<<nb, 256, 0, st>>>(AKB_SC2P_ARGS); }
#undef AKB_SC2P_ARGS
  return cudaGetLastError();
}

}  // namespace akb
