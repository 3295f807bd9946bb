// smr_fc_body.h -- per-element bodies of the kernels that move FACE-CENTRED fields and EDGE EMFs across refinement
// levels (MHD on statically refined meshes).  The kernels in smr.cu are one-line wrappers
// around these functions; tests/emul/smr_fc_emul.cpp compiles the SAME functions with g++ and runs them in a host
// loop over the plan's tables, so the index logic is checked against the oracle in the CPU test suite (test
// infrastructure: the library itself never executes them on the host).
//
// Reference: src/mesh/mesh_refinement.cpp:1290-1383 (RestrictFC), src/bvals/bvals_fc.cpp:63-436 (PackAndSendFC /
// RecvAndUnpackFC across levels), src/bvals/prolongation.cpp:28-357,556-785 (FillCoarseInBndryFC, ProlongateFC),
// src/bvals/flux_correct_fc.cpp:29-1034 (SendE / RecvE with coarser and finer neighbours).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define AK_HD __host__ __device__ __forceinline__
#else
#define AK_HD inline
#endif

namespace akb_smr {

struct FcDims {
  int n[3], cn[3];              // cell extents of the fine / coarse arrays (i, j, k)
  int s[3], e[3], cs[3], ce[3];
  int multi_d, three_d, ng, nmb;
  long f[3], cf[3], ef[3];      // per-block sizes: fine faces, coarse faces, edge fields (component 0,1,2)
};

// ---- tables written by the host plan (smr.cu) -------------------------------------------------------------
struct FcSrc { int src_m, src_coarse, off[3]; };                 // per (dst block, dst slot, component)
struct FcBox { int m, v, lo[3], n[3], own_axis, own_sign; };     // fill / shared-face boxes (coarse indices)
struct FcCell { int m, lo[3], n[3]; };                           // interior-face boxes (coarse cells)
struct EmfPack { int src_m, v, coarser, lo[3], n[3]; long off; };
enum { EMF_SUM0 = 0, EMF_SUM1 = 1, EMF_MUL = 2, EMF_DIV = 3 };
struct EmfOp { int type, lo[3], hi[3]; long off; double div; };  // off: receive segment; div: divisor (EMF_DIV)
struct EmfRange { int sum0, sum1, avg, end; };                   // per (block, component): [sum0,sum1) [sum1,avg) [avg,end)

AK_HD long fidx(const FcDims &D, int v, int k, int j, int i) {   // fine face array of component v
  return ((long)k * (D.n[1] + (v == 1)) + j) * (D.n[0] + (v == 0)) + i;
}
AK_HD long cidx(const FcDims &D, int v, int k, int j, int i) {   // coarse face array
  return ((long)k * (D.cn[1] + (v == 1)) + j) * (D.cn[0] + (v == 0)) + i;
}
AK_HD long eidx(const FcDims &D, int v, int k, int j, int i) {   // edge field: node-like across its own direction
  return ((long)k * (D.n[1] + (v != 1)) + j) * (D.n[0] + (v != 0)) + i;
}

// prolongation.cpp:75-90
AK_HD bool fc_is_active(const FcDims &D, int v, int k, int j, int i) {
  return i >= D.s[0] && i <= D.e[0] + (v == 0) && j >= D.s[1] && j <= D.e[1] + (v == 1) && k >= D.s[2] && k <= D.e[2] + (v == 2);
}

// mesh_refinement.cpp:1290-1383: coarse face (k,j,i) of component v from the fine faces it covers.  a = this
// block's fine array of that component.  In inactive directions both coarse planes of a staggered component take
// the value computed from fine plane 0, as the reference's 1-D / 2-D branches do.
AK_HD double restrict_face(const FcDims &D, int v, const double *a, int k, int j, int i) {
  const int fi = 2 * i - D.cs[0];
  if (!D.multi_d) {
    if (v == 0) return a[fidx(D, 0, 0, 0, fi)];
    return 0.5 * (a[fidx(D, v, 0, 0, fi)] + a[fidx(D, v, 0, 0, fi + 1)]);
  }
  const int fj = 2 * j - D.cs[1];
  if (!D.three_d) {
    if (v == 0) return 0.5 * (a[fidx(D, 0, 0, fj, fi)] + a[fidx(D, 0, 0, fj + 1, fi)]);
    if (v == 1) return 0.5 * (a[fidx(D, 1, 0, fj, fi)] + a[fidx(D, 1, 0, fj, fi + 1)]);
    return 0.25 * (a[fidx(D, 2, 0, fj, fi)] + a[fidx(D, 2, 0, fj, fi + 1)] + a[fidx(D, 2, 0, fj + 1, fi)] + a[fidx(D, 2, 0, fj + 1, fi + 1)]);
  }
  const int fk = 2 * k - D.cs[2];
  if (v == 0) return 0.25 * (a[fidx(D, 0, fk, fj, fi)] + a[fidx(D, 0, fk, fj + 1, fi)] + a[fidx(D, 0, fk + 1, fj, fi)] + a[fidx(D, 0, fk + 1, fj + 1, fi)]);
  if (v == 1) return 0.25 * (a[fidx(D, 1, fk, fj, fi)] + a[fidx(D, 1, fk, fj, fi + 1)] + a[fidx(D, 1, fk + 1, fj, fi)] + a[fidx(D, 1, fk + 1, fj, fi + 1)]);
  return 0.25 * (a[fidx(D, 2, fk, fj, fi)] + a[fidx(D, 2, fk, fj, fi + 1)] + a[fidx(D, 2, fk, fj + 1, fi)] + a[fidx(D, 2, fk, fj + 1, fi + 1)]);
}

// RestrictFC: t over (block, active coarse faces of component v)
AK_HD void fc_restrict_elem(const FcDims &D, int v, long t, const double *b, double *cb) {
  const int e0 = (D.ce[0] - D.cs[0] + 1) + (v == 0), e1 = (D.ce[1] - D.cs[1] + 1) + (v == 1), e2 = (D.ce[2] - D.cs[2] + 1) + (v == 2);
  const int i = D.cs[0] + (int)(t % e0); t /= e0;
  const int j = D.cs[1] + (int)(t % e1); t /= e1;
  const int k = D.cs[2] + (int)(t % e2);
  const long m = t / e2;
  cb[m * D.cf[v] + cidx(D, v, k, j, i)] = restrict_face(D, v, b + m * D.f[v], k, j, i);
}
AK_HD long fc_restrict_count(const FcDims &D, int v) {
  return (long)D.nmb * ((D.ce[0] - D.cs[0] + 1) + (v == 0)) * ((D.ce[1] - D.cs[1] + 1) + (v == 1)) * ((D.ce[2] - D.cs[2] + 1) + (v == 2));
}

// SendB/RecvB across levels as a gather: every destination element knows (map byte) which neighbour slot wrote
// it LAST in the reference's serial unpack order; that slot's source element is read directly.  A coarse source
// (data sent to a coarser neighbour) is restricted on the fly from the sender's ACTIVE fine faces, which this
// step never writes, so the launch has no read-after-write hazard.
AK_HD double fc_exchange_value(const FcDims &D, int v, int dst_coarse, int m, int r, int n, const FcSrc *src, const double *b) {
  const int x0 = (dst_coarse ? D.cn[0] : D.n[0]) + (v == 0), x1 = (dst_coarse ? D.cn[1] : D.n[1]) + (v == 1);
  const int i = r % x0; r /= x0;              // r = element of block m's array, n = the winning slot (32-bit index math)
  const int j = r % x1;
  const int k = r / x1;
  const FcSrc S = src[((long)m * 56 + n) * 3 + v];
  const int si = i + S.off[0], sj = j + S.off[1], sk = k + S.off[2];
  const double *a = b + (long)S.src_m * D.f[v];
  return S.src_coarse ? restrict_face(D, v, a, sk, sj, si) : a[fidx(D, v, sk, sj, si)];
}
AK_HD void fc_exchange_elem(const FcDims &D, int v, int dst_coarse, int m, int r, const unsigned char *map, const FcSrc *src,
                            double *b, double *cb) {
  const long t = (long)m * (dst_coarse ? D.cf[v] : D.f[v]) + r;
  const unsigned char n = map[t];
  if (n == 255) return;
  const double val = fc_exchange_value(D, v, dst_coarse, m, r, n, src, b);
  if (dst_coarse) cb[t] = val; else b[t] = val;
}

// FillCoarseInBndryFC (prolongation.cpp:556-650): t over (descriptor box)
AK_HD void fc_fill_elem(const FcDims &D, const FcBox &d, int t, const double *b1, const double *b2, const double *b3,
                        double *cb1, double *cb2, double *cb3) {
  const int i = d.lo[0] + t % d.n[0]; t /= d.n[0];      // 32-bit index math: a box holds far fewer than 2^31 elements
  const int j = d.lo[1] + t % d.n[1];
  const int k = d.lo[2] + t / d.n[1];
  const double *a = (d.v == 0 ? b1 : d.v == 1 ? b2 : b3) + (long)d.m * D.f[d.v];
  double *c = (d.v == 0 ? cb1 : d.v == 1 ? cb2 : cb3) + (long)d.m * D.cf[d.v];
  c[cidx(D, d.v, k, j, i)] = restrict_face(D, d.v, a, k, j, i);
}

AK_HD double fc_sgn(double x) { return (x < 0.0) ? -1.0 : 1.0; }
AK_HD double fc_mm(double dl, double dr) {
  const double al = fabs(dl), ar = fabs(dr);
  return 0.125 * (fc_sgn(dl) + fc_sgn(dr)) * ((al < ar) ? al : ar);
}

// prolongation.cpp:92-133: ghost faces are always written; an ACTIVE face only on the fine/coarse interface normal
// to it, from the face slot of that side (own_axis / own_sign, set by the host when every neighbour across that
// face is coarser).
AK_HD bool fc_can_store(const FcDims &D, int v, int k, int j, int i, int own_axis, int own_sign) {
  if (!fc_is_active(D, v, k, j, i)) return true;
  if (own_axis != v) return false;
  const int c = (v == 0) ? i : (v == 1) ? j : k;
  return (own_sign < 0) ? (c == D.s[v]) : (c == D.e[v] + 1);
}

// ProlongateFC, shared faces (prolongation.cpp:651-760): t over (descriptor box of coarse faces)
AK_HD void fc_prolong_shared_elem(const FcDims &D, const FcBox &d, int t, double *b1, double *b2, double *b3,
                                  const double *cb1, const double *cb2, const double *cb3) {
  const int i = d.lo[0] + t % d.n[0]; t /= d.n[0];      // 32-bit index math: a box holds far fewer than 2^31 elements
  const int j = d.lo[1] + t % d.n[1];
  const int k = d.lo[2] + t / d.n[1];
  const int v = d.v;
  const int fi = (i - D.cs[0]) * 2 + D.s[0], fj = D.multi_d ? ((j - D.cs[1]) * 2 + D.s[1]) : j, fk = D.three_d ? ((k - D.cs[2]) * 2 + D.s[2]) : k;
  const double *c = (v == 0 ? cb1 : v == 1 ? cb2 : cb3) + (long)d.m * D.cf[v];
  double *a = (v == 0 ? b1 : v == 1 ? b2 : b3) + (long)d.m * D.f[v];
#define FC_ST(K_, J_, I_, val_) do { if (fc_can_store(D, v, K_, J_, I_, d.own_axis, d.own_sign)) a[fidx(D, v, K_, J_, I_)] = (val_); } while (0)
  const double q = c[cidx(D, v, k, j, i)];
  if (v == 0) {
    double dv2 = 0.0, dv3 = 0.0;
    if (D.multi_d) dv2 = fc_mm(q - c[cidx(D, 0, k, j - 1, i)], c[cidx(D, 0, k, j + 1, i)] - q);
    if (D.three_d) dv3 = fc_mm(q - c[cidx(D, 0, k - 1, j, i)], c[cidx(D, 0, k + 1, j, i)] - q);
    FC_ST(fk, fj, fi, q - dv2 - dv3);
    if (D.multi_d) FC_ST(fk, fj + 1, fi, q + dv2 - dv3);
    if (D.three_d) {
      FC_ST(fk + 1, fj, fi, q - dv2 + dv3);
      FC_ST(fk + 1, fj + 1, fi, q + dv2 + dv3);
    }
  } else if (v == 1) {
    const double dv1 = fc_mm(q - c[cidx(D, 1, k, j, i - 1)], c[cidx(D, 1, k, j, i + 1)] - q);
    double dv3 = 0.0;
    if (D.three_d) dv3 = fc_mm(q - c[cidx(D, 1, k - 1, j, i)], c[cidx(D, 1, k + 1, j, i)] - q);
    FC_ST(fk, fj, fi, q - dv1 - dv3);
    FC_ST(fk, fj, fi + 1, q + dv1 - dv3);
    if (D.three_d) {
      FC_ST(fk + 1, fj, fi, q - dv1 + dv3);
      FC_ST(fk + 1, fj, fi + 1, q + dv1 + dv3);
    }
  } else {
    const double dv1 = fc_mm(q - c[cidx(D, 2, k, j, i - 1)], c[cidx(D, 2, k, j, i + 1)] - q);
    double dv2 = 0.0;
    if (D.multi_d) dv2 = fc_mm(q - c[cidx(D, 2, k, j - 1, i)], c[cidx(D, 2, k, j + 1, i)] - q);
    FC_ST(fk, fj, fi, q - dv1 - dv2);
    FC_ST(fk, fj, fi + 1, q + dv1 - dv2);
    if (D.multi_d) {
      FC_ST(fk, fj + 1, fi, q - dv1 + dv2);
      FC_ST(fk, fj + 1, fi + 1, q + dv1 + dv2);
    }
  }
#undef FC_ST
}

// ProlongateFC, interior faces of the refined ghost cells (prolongation.cpp:253-357, Toth & Roe 2002): reads only
// faces shared with the coarse cell (even offsets), writes only the faces between them (odd offsets).
AK_HD void fc_prolong_interior_elem(const FcDims &D, const FcCell &d, int t, double *b1, double *b2, double *b3, bool force = false) {
  const int i = d.lo[0] + t % d.n[0]; t /= d.n[0];      // 32-bit index math: a box holds far fewer than 2^31 elements
  const int j = d.lo[1] + t % d.n[1];
  const int k = d.lo[2] + t / d.n[1];
  const int fi = (i - D.cs[0]) * 2 + D.s[0], fj = (j - D.cs[1]) * 2 + D.s[1], fk = (k - D.cs[2]) * 2 + D.s[2];
  double *x1 = b1 + (long)d.m * D.f[0], *x2 = b2 + (long)d.m * D.f[1], *x3 = b3 + (long)d.m * D.f[2];
#define B1(K_, J_, I_) x1[fidx(D, 0, K_, J_, I_)]
#define B2(K_, J_, I_) x2[fidx(D, 1, K_, J_, I_)]
#define B3(K_, J_, I_) x3[fidx(D, 2, K_, J_, I_)]
// force: write active faces too (blocks created by refinement, RepairAMRFC); ghost-zone prolongation never does
#define ST(v_, X_, K_, J_, I_, val_) do { if (force || !fc_is_active(D, v_, K_, J_, I_)) X_(K_, J_, I_) = (val_); } while (0)
  if (!D.multi_d) {
    ST(0, B1, fk, fj, fi + 1, 0.5 * (B1(fk, fj, fi) + B1(fk, fj, fi + 2)));
  } else if (D.three_d) {
    double Uxx = 0.0, Vyy = 0.0, Wzz = 0.0, Uxyz = 0.0, Vxyz = 0.0, Wxyz = 0.0;
    for (int jj = 0; jj < 2; jj++) {
      const int jsgn = 2 * jj - 1, fjj = fj + jj, fjp = fj + 2 * jj;
      for (int ii = 0; ii < 2; ii++) {
        const int isgn = 2 * ii - 1, fii = fi + ii, fip = fi + 2 * ii;
        Uxx += isgn * (jsgn * (B2(fk, fjp, fii) + B2(fk + 1, fjp, fii)) + (B3(fk + 2, fjj, fii) - B3(fk, fjj, fii)));
        Vyy += jsgn * ((B3(fk + 2, fjj, fii) - B3(fk, fjj, fii)) + isgn * (B1(fk, fjj, fip) + B1(fk + 1, fjj, fip)));
        Wzz += isgn * (B1(fk + 1, fjj, fip) - B1(fk, fjj, fip)) + jsgn * (B2(fk + 1, fjp, fii) - B2(fk, fjp, fii));
        Uxyz += isgn * jsgn * (B1(fk + 1, fjj, fip) - B1(fk, fjj, fip));
        Vxyz += isgn * jsgn * (B2(fk + 1, fjp, fii) - B2(fk, fjp, fii));
        Wxyz += isgn * jsgn * (B3(fk + 2, fjj, fii) - B3(fk, fjj, fii));
      }
    }
    Uxx *= 0.125; Vyy *= 0.125; Wzz *= 0.125;
    Uxyz *= 0.0625; Vxyz *= 0.0625; Wxyz *= 0.0625;
    ST(0, B1, fk, fj, fi + 1, 0.5 * (B1(fk, fj, fi) + B1(fk, fj, fi + 2)) + Uxx - Vxyz - Wxyz);
    ST(0, B1, fk, fj + 1, fi + 1, 0.5 * (B1(fk, fj + 1, fi) + B1(fk, fj + 1, fi + 2)) + Uxx - Vxyz + Wxyz);
    ST(0, B1, fk + 1, fj, fi + 1, 0.5 * (B1(fk + 1, fj, fi) + B1(fk + 1, fj, fi + 2)) + Uxx + Vxyz - Wxyz);
    ST(0, B1, fk + 1, fj + 1, fi + 1, 0.5 * (B1(fk + 1, fj + 1, fi) + B1(fk + 1, fj + 1, fi + 2)) + Uxx + Vxyz + Wxyz);
    ST(1, B2, fk, fj + 1, fi, 0.5 * (B2(fk, fj, fi) + B2(fk, fj + 2, fi)) + Vyy - Uxyz - Wxyz);
    ST(1, B2, fk, fj + 1, fi + 1, 0.5 * (B2(fk, fj, fi + 1) + B2(fk, fj + 2, fi + 1)) + Vyy - Uxyz + Wxyz);
    ST(1, B2, fk + 1, fj + 1, fi, 0.5 * (B2(fk + 1, fj, fi) + B2(fk + 1, fj + 2, fi)) + Vyy + Uxyz - Wxyz);
    ST(1, B2, fk + 1, fj + 1, fi + 1, 0.5 * (B2(fk + 1, fj, fi + 1) + B2(fk + 1, fj + 2, fi + 1)) + Vyy + Uxyz + Wxyz);
    ST(2, B3, fk + 1, fj, fi, 0.5 * (B3(fk + 2, fj, fi) + B3(fk, fj, fi)) + Wzz - Uxyz - Vxyz);
    ST(2, B3, fk + 1, fj, fi + 1, 0.5 * (B3(fk + 2, fj, fi + 1) + B3(fk, fj, fi + 1)) + Wzz - Uxyz + Vxyz);
    ST(2, B3, fk + 1, fj + 1, fi, 0.5 * (B3(fk + 2, fj + 1, fi) + B3(fk, fj + 1, fi)) + Wzz + Uxyz - Vxyz);
    ST(2, B3, fk + 1, fj + 1, fi + 1, 0.5 * (B3(fk + 2, fj + 1, fi + 1) + B3(fk, fj + 1, fi + 1)) + Wzz + Uxyz + Vxyz);
  } else {
    const double tmp1 = 0.25 * (B2(fk, fj + 2, fi + 1) - B2(fk, fj, fi + 1) - B2(fk, fj + 2, fi) + B2(fk, fj, fi));
    const double tmp2 = 0.25 * (B1(fk, fj, fi) - B1(fk, fj, fi + 2) - B1(fk, fj + 1, fi) + B1(fk, fj + 1, fi + 2));
    ST(0, B1, fk, fj, fi + 1, 0.5 * (B1(fk, fj, fi) + B1(fk, fj, fi + 2)) + tmp1);
    ST(0, B1, fk, fj + 1, fi + 1, 0.5 * (B1(fk, fj + 1, fi) + B1(fk, fj + 1, fi + 2)) + tmp1);
    ST(1, B2, fk, fj + 1, fi, 0.5 * (B2(fk, fj, fi) + B2(fk, fj + 2, fi)) + tmp2);
    ST(1, B2, fk, fj + 1, fi + 1, 0.5 * (B2(fk, fj, fi + 1) + B2(fk, fj + 2, fi + 1)) + tmp2);
  }
#undef ST
#undef B1
#undef B2
#undef B3
}

// SendE, pack (flux_correct_fc.cpp:29-366): the boundary EMFs toward a same-level neighbour as they are, toward a
// coarser neighbour averaged over the two fine edges of a coarse edge (inactive directions read index 0).
AK_HD void emf_pack_elem(const FcDims &D, const EmfPack &d, int t, const double *e1, const double *e2, const double *e3, double *buf) {
  const long slot = d.off + t;
  const int i = d.lo[0] + t % d.n[0]; t /= d.n[0];      // 32-bit index math: a box holds far fewer than 2^31 elements
  const int j = d.lo[1] + t % d.n[1];
  const int k = d.lo[2] + t / d.n[1];
  const int v = d.v;
  const double *a = (v == 0 ? e1 : v == 1 ? e2 : e3) + (long)d.src_m * D.ef[v];
  double r;
  if (!d.coarser) r = a[eidx(D, v, k, j, i)];
  else {
    const int fi = 2 * i - D.cs[0], fj = D.multi_d ? 2 * j - D.cs[1] : 0, fk = D.three_d ? 2 * k - D.cs[2] : 0;
    const bool act = (v == 0) || (v == 1 ? D.multi_d : D.three_d);
    if (act) r = 0.5 * (a[eidx(D, v, fk, fj, fi)] + a[eidx(D, v, fk + (v == 2), fj + (v == 1), fi + (v == 0))]);
    else r = a[eidx(D, v, fk, fj, fi)];
  }
  buf[slot] = r;
}

// The elements of one block's edge field that lie on the block surface (node-like coordinate at either end in at
// least one of the two directions across the field), enumerated without duplicates: first the two planes at the ends
// of direction a = (v+1)%3, then the two planes at the ends of b = (v+2)%3 minus the lines already counted.
AK_HD int emf_surface_count(const FcDims &D, int v) {
  const int a = (v + 1) % 3, b = (v + 2) % 3;
  const int na1 = D.e[a] - D.s[a] + 2, nb1 = D.e[b] - D.s[b] + 2, nc = D.e[v] - D.s[v] + 1;
  return 2 * nb1 * nc + 2 * (na1 - 2) * nc;
}
AK_HD void emf_surface_decode(const FcDims &D, int v, int q, int x[3]) {
  const int a = (v + 1) % 3, b = (v + 2) % 3;
  const int na1 = D.e[a] - D.s[a] + 2, nb1 = D.e[b] - D.s[b] + 2, nc = D.e[v] - D.s[v] + 1;
  const int s1 = 2 * nb1 * nc;
  if (q < s1) {
    const int side = q / (nb1 * nc), r = q - side * (nb1 * nc);
    x[a] = side ? D.e[a] + 1 : D.s[a];
    x[b] = D.s[b] + r / nc;
    x[v] = D.s[v] + r % nc;
  } else {
    q -= s1;
    const int per = (na1 - 2) * nc, side = q / per, r = q - side * per;
    x[a] = D.s[a] + 1 + r / nc;
    x[b] = side ? D.e[b] + 1 : D.s[b];
    x[v] = D.s[v] + r % nc;
  }
}

// RecvE (flux_correct_fc.cpp:374-1034) as a gather over the SURFACE elements of one block's edge field: same-level
// sums in slot order, zero where a finer neighbour covers the element, sums from finer neighbours in slot order,
// then the average (x0.5 on faces, / number of contributions on block edges).  q = surface element of (m, v).
AK_HD void emf_apply_elem(const FcDims &D, int v, int m, int q, const EmfRange *ranges, const EmfOp *ops, const double *buf, double *e) {
  int x[3];
  emf_surface_decode(D, v, q, x);
  const int i = x[0], j = x[1], k = x[2];
  const long t = (long)m * D.ef[v] + eidx(D, v, k, j, i);
  const EmfRange R = ranges[m * 3 + v];
  double val = e[t];
  bool touched = false;
  auto inside = [&](const EmfOp &o) { return i >= o.lo[0] && i <= o.hi[0] && j >= o.lo[1] && j <= o.hi[1] && k >= o.lo[2] && k <= o.hi[2]; };
  auto bufat = [&](const EmfOp &o) {
    const int ni = o.hi[0] - o.lo[0] + 1, nj = o.hi[1] - o.lo[1] + 1;
    return buf[o.off + (i - o.lo[0]) + (long)ni * ((j - o.lo[1]) + (long)nj * (k - o.lo[2]))];
  };
  for (int p = R.sum0; p < R.sum1; ++p) if (inside(ops[p])) { val += bufat(ops[p]); touched = true; }
  bool zero = false;
  for (int p = R.sum1; p < R.avg; ++p) zero |= inside(ops[p]);
  if (zero) {
    val = 0.0; touched = true;
    for (int p = R.sum1; p < R.avg; ++p) if (inside(ops[p])) val += bufat(ops[p]);
  }
  for (int p = R.avg; p < R.end; ++p) if (inside(ops[p])) {
    if (ops[p].type == EMF_MUL) val *= 0.5; else val /= ops[p].div;
    touched = true;
  }
  if (touched) e[t] = val;
}

}  // namespace akb_smr
