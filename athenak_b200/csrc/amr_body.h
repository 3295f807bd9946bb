// amr_body.h -- per-element bodies of the adaptive-mesh-refinement kernels (dynamic refinement).
// The kernels in amr.cu are thin wrappers; tests/emul/amr_emul.cpp compiles the SAME functions with g++ so the CPU test
// suite checks them against oracle/amr_oracle.c (test infrastructure: the library never runs them on the host).
//
// Reference: src/mesh/refinement_criteria.cpp:177-336 (block extrema), src/mesh/mesh_refinement.cpp:696-812 (Derefine*),
// :817-914 (Copy*), :919-1033 (CopyForRefinement*), :1038-1184 (RefineCC / RefineFC), :1189-1221 (RepairAMRFC).
//
// Design: the reference moves blocks IN PLACE inside over-allocated Views (move-left pass, move-right pass, staging
// through the coarse arrays).  Here the host owns an OLD and a NEW set of arrays and every element of the new arrays is
// GATHERED from the old ones in one launch per field: unchanged blocks are copied, a derefined block restricts its
// nleaf old siblings on the fly, a block created by refinement prolongates its octant of the old parent directly from the
// parent's fine array (the "coarse array" of the reference is exactly that octant plus ng cells).  Same expressions,
// same bits, no ordering between blocks, no staging copies.
#pragma once
#include "smr_fc_body.h"

namespace akb_smr {

enum { AMR_COPY = 0, AMR_DEREF = 1, AMR_REFINE = 2 };
struct AmrBlock { int kind, src, off[3]; };       // per NEW block: old block (same / first sibling / parent), octant offset in cells

AK_HD long ccidx(const FcDims &D, int k, int j, int i) { return ((long)k * D.n[1] + j) * D.n[0] + i; }

// cell-centred variable, one element (k,j,i) of new block m, variable v.  uo / un: old / new arrays [nmb][nvar][n3][n2][n1]
AK_HD void amr_move_cc_elem(const FcDims &D, const AmrBlock &B, int m, int nvar, int v, int k, int j, int i, const double *uo, double *un) {
  const long cc = (long)D.n[0] * D.n[1] * D.n[2];
  double *dst = un + ((long)m * nvar + v) * cc + ccidx(D, k, j, i);
  if (B.kind == AMR_COPY) { *dst = uo[((long)B.src * nvar + v) * cc + ccidx(D, k, j, i)]; return; }
  const bool act = i >= D.s[0] && i <= D.e[0] && j >= D.s[1] && j <= D.e[1] && k >= D.s[2] && k <= D.e[2];
  if (!act) return;                                // ghost zones of changed blocks are rebuilt by the exchange that follows
  const int cnx1 = D.ce[0] - D.cs[0] + 1, cnx2 = D.ce[1] - D.cs[1] + 1, cnx3 = D.ce[2] - D.cs[2] + 1;
  if (B.kind == AMR_DEREF) {                       // RestrictCC of the sibling that covers this cell (mesh_refinement.cpp:1226-1284)
    const int o1 = (i - D.s[0]) >= cnx1, o2 = D.multi_d && (j - D.s[1]) >= cnx2, o3 = D.three_d && (k - D.s[2]) >= cnx3;
    const double *a = uo + ((long)(B.src + o1 + 2 * o2 + 4 * o3) * nvar + v) * cc;
    const int fi = 2 * (i - D.s[0] - o1 * cnx1) + D.s[0], fj = D.multi_d ? 2 * (j - D.s[1] - o2 * cnx2) + D.s[1] : j,
              fk = D.three_d ? 2 * (k - D.s[2] - o3 * cnx3) + D.s[2] : k;
    if (!D.multi_d) *dst = 0.5 * (a[ccidx(D, fk, fj, fi)] + a[ccidx(D, fk, fj, fi + 1)]);
    else if (!D.three_d) *dst = 0.25 * (a[ccidx(D, fk, fj, fi)] + a[ccidx(D, fk, fj, fi + 1)] + a[ccidx(D, fk, fj + 1, fi)] + a[ccidx(D, fk, fj + 1, fi + 1)]);
    else *dst = 0.125 * (a[ccidx(D, fk, fj, fi)] + a[ccidx(D, fk, fj, fi + 1)] + a[ccidx(D, fk, fj + 1, fi)] + a[ccidx(D, fk, fj + 1, fi + 1)]
                       + a[ccidx(D, fk + 1, fj, fi)] + a[ccidx(D, fk + 1, fj, fi + 1)] + a[ccidx(D, fk + 1, fj + 1, fi)] + a[ccidx(D, fk + 1, fj + 1, fi + 1)]);
    return;
  }
  // AMR_REFINE: ProlongCC (mesh/prolongation.hpp:21-60) of the parent's cell that contains this fine cell
  const double *a = uo + ((long)B.src * nvar + v) * cc;
  const int ri = i - D.s[0], rj = j - D.s[1], rk = k - D.s[2];
  const int pi = D.s[0] + (ri >> 1) + B.off[0], pj = D.multi_d ? D.s[1] + (rj >> 1) + B.off[1] : j, pk = D.three_d ? D.s[2] + (rk >> 1) + B.off[2] : k;
  const double q = a[ccidx(D, pk, pj, pi)];
  const double dv1 = fc_mm(q - a[ccidx(D, pk, pj, pi - 1)], a[ccidx(D, pk, pj, pi + 1)] - q);
  const double dv2 = D.multi_d ? fc_mm(q - a[ccidx(D, pk, pj - 1, pi)], a[ccidx(D, pk, pj + 1, pi)] - q) : 0.0;
  const double dv3 = D.three_d ? fc_mm(q - a[ccidx(D, pk - 1, pj, pi)], a[ccidx(D, pk + 1, pj, pi)] - q) : 0.0;
  // the eight expressions of ProlongCC, selected by the position of the fine cell inside the coarse one
  const bool hi1 = ri & 1, hi2 = D.multi_d && (rj & 1), hi3 = D.three_d && (rk & 1);
  double r = hi1 ? q + dv1 : q - dv1;
  r = hi2 ? r + dv2 : r - dv2;
  r = hi3 ? r + dv3 : r - dv3;
  *dst = r;
}

// face field component v, one element of new block m: unchanged blocks are copied whole, a derefined block takes the
// RestrictFC value of the sibling that owns the face (at the plane where two siblings meet the reference's copy order lets
// the sibling on the + side win).  bo / bn = old / new arrays of that component.
AK_HD void amr_move_fc_elem(const FcDims &D, const AmrBlock &B, int m, int v, int k, int j, int i, const double *bo, double *bn) {
  double *dst = bn + (long)m * D.f[v] + fidx(D, v, k, j, i);
  if (B.kind == AMR_COPY) { *dst = bo[(long)B.src * D.f[v] + fidx(D, v, k, j, i)]; return; }
  if (B.kind != AMR_DEREF || !fc_is_active(D, v, k, j, i)) return;
  const int x[3] = {i, j, k};
  int o[3], c[3];
  for (int d = 0; d < 3; ++d) {
    const bool actd = (d == 0) || (d == 1 ? D.multi_d : D.three_d);
    const int cnx = D.ce[d] - D.cs[d] + 1;
    if (!actd) { o[d] = 0; c[d] = x[d]; continue; }
    const int rel = x[d] - D.s[d];
    o[d] = rel >= cnx;                                 // a face on the mid-plane belongs to the + sibling
    if (v == d && rel == 2 * cnx) o[d] = 1;            // the last face of the block
    c[d] = D.cs[d] + rel - o[d] * cnx;
  }
  const double *a = bo + (long)(B.src + o[0] + 2 * o[1] + 4 * o[2]) * D.f[v];
  *dst = restrict_face(D, v, a, c[2], c[1], c[0]);
}

// RefineFC, shared faces (mesh_refinement.cpp:1120-1158): t = coarse face (k,j,i) of the child, index space of its coarse
// array; the coarse values are the parent's fine faces at (index + octant offset)
AK_HD void amr_refine_fc_shared_elem(const FcDims &D, const AmrBlock &B, int m, int v, int k, int j, int i, const double *bo, double *bn) {
  const double *a = bo + (long)B.src * D.f[v];
  double *f = bn + (long)m * D.f[v];
  const int o1 = B.off[0], o2 = B.off[1], o3 = B.off[2];
  auto P = [&](int kk, int jj, int ii) { return a[fidx(D, v, kk + o3, jj + o2, ii + o1)]; };
  const int fi = (i - D.cs[0]) * 2 + D.s[0], fj = D.multi_d ? ((j - D.cs[1]) * 2 + D.s[1]) : j, fk = D.three_d ? ((k - D.cs[2]) * 2 + D.s[2]) : k;
  const double q = P(k, j, i);
  if (v == 0) {
    const double dv2 = D.multi_d ? fc_mm(q - P(k, j - 1, i), P(k, j + 1, i) - q) : 0.0;
    const double dv3 = D.three_d ? fc_mm(q - P(k - 1, j, i), P(k + 1, j, i) - q) : 0.0;
    f[fidx(D, 0, fk, fj, fi)] = q - dv2 - dv3;
    if (D.multi_d) f[fidx(D, 0, fk, fj + 1, fi)] = q + dv2 - dv3;
    if (D.three_d) { f[fidx(D, 0, fk + 1, fj, fi)] = q - dv2 + dv3; f[fidx(D, 0, fk + 1, fj + 1, fi)] = q + dv2 + dv3; }
  } else if (v == 1) {
    const double dv1 = fc_mm(q - P(k, j, i - 1), P(k, j, i + 1) - q);
    const double dv3 = D.three_d ? fc_mm(q - P(k - 1, j, i), P(k + 1, j, i) - q) : 0.0;
    f[fidx(D, 1, fk, fj, fi)] = q - dv1 - dv3;
    f[fidx(D, 1, fk, fj, fi + 1)] = q + dv1 - dv3;
    if (D.three_d) { f[fidx(D, 1, fk + 1, fj, fi)] = q - dv1 + dv3; f[fidx(D, 1, fk + 1, fj, fi + 1)] = q + dv1 + dv3; }
  } else {
    const double dv1 = fc_mm(q - P(k, j, i - 1), P(k, j, i + 1) - q);
    const double dv2 = D.multi_d ? fc_mm(q - P(k, j - 1, i), P(k, j + 1, i) - q) : 0.0;
    f[fidx(D, 2, fk, fj, fi)] = q - dv1 - dv2;
    f[fidx(D, 2, fk, fj, fi + 1)] = q + dv1 - dv2;
    if (D.multi_d) { f[fidx(D, 2, fk, fj + 1, fi)] = q - dv1 + dv2; f[fidx(D, 2, fk, fj + 1, fi + 1)] = q + dv1 + dv2; }
  }
}

// block extrema for the refinement criteria (refinement_criteria.cpp:177-336): value of one active cell.
// method 0 = the variable itself, 1 = slope, 2 = second derivative
AK_HD double amr_criterion_value(const FcDims &D, int method, const double *a, int k, int j, int i) {
  const double q = a[ccidx(D, k, j, i)];
  if (method == 0) return q;
  if (method == 1) {
    double t = a[ccidx(D, k, j, i + 1)] - a[ccidx(D, k, j, i - 1)];
    double d2 = t * t;
    if (D.multi_d) { t = a[ccidx(D, k, j + 1, i)] - a[ccidx(D, k, j - 1, i)]; d2 += t * t; }
    if (D.three_d) { t = a[ccidx(D, k + 1, j, i)] - a[ccidx(D, k - 1, j, i)]; d2 += t * t; }
    return 0.5 * sqrt(d2) / q;
  }
  double d2q = a[ccidx(D, k, j, i + 1)] - 2.0 * q + a[ccidx(D, k, j, i - 1)];
  if (D.multi_d) d2q += (a[ccidx(D, k, j + 1, i)] - 2.0 * q + a[ccidx(D, k, j - 1, i)]);
  if (D.three_d) d2q += (a[ccidx(D, k + 1, j, i)] - 2.0 * q + a[ccidx(D, k - 1, j, i)]);
  return fabs(d2q) / q;
}

// ---- flat-index entry points: the index decode lives here too, so that the g++ emulation harness of the tests exercises it and
// the CUDA wrappers in amr.cu contain nothing but their grid-stride loops
AK_HD void amr_move_cc_flat(const FcDims &D, const AmrBlock &B, int m, int nvar, int v, int t, const double *uo, double *un) {
  const int i = t % D.n[0], r = t / D.n[0];
  amr_move_cc_elem(D, B, m, nvar, v, r / D.n[1], r % D.n[1], i, uo, un);
}
AK_HD int amr_move_cc_count(const FcDims &D) { return D.n[0] * D.n[1] * D.n[2]; }
AK_HD void amr_move_fc_flat(const FcDims &D, const AmrBlock &B, int m, int v, int t, const double *bo, double *bn) {
  const int x0 = D.n[0] + (v == 0), x1 = D.n[1] + (v == 1);
  const int i = t % x0, r = t / x0;
  amr_move_fc_elem(D, B, m, v, r / x1, r % x1, i, bo, bn);
}
AK_HD int amr_refine_fc_shared_count(const FcDims &D, int v) {
  return ((D.ce[0] - D.cs[0] + 1) + (v == 0)) * ((D.ce[1] - D.cs[1] + 1) + (v == 1)) * ((D.ce[2] - D.cs[2] + 1) + (v == 2));
}
AK_HD void amr_refine_fc_shared_flat(const FcDims &D, const AmrBlock &B, int m, int v, int t, const double *bo, double *bn) {
  const int e0 = (D.ce[0] - D.cs[0] + 1) + (v == 0), e1 = (D.ce[1] - D.cs[1] + 1) + (v == 1);
  const int i = D.cs[0] + t % e0, r = t / e0;
  amr_refine_fc_shared_elem(D, B, m, v, D.cs[2] + r / e1, D.cs[1] + r % e1, i, bo, bn);
}
AK_HD int amr_active_count(const FcDims &D) { return (D.e[0] - D.s[0] + 1) * (D.e[1] - D.s[1] + 1) * (D.e[2] - D.s[2] + 1); }
AK_HD double amr_criterion_flat(const FcDims &D, int method, const double *a, int t) {
  const int nx1 = D.e[0] - D.s[0] + 1, nx2 = D.e[1] - D.s[1] + 1;
  const int i = D.s[0] + t % nx1, r = t / nx1;
  return amr_criterion_value(D, method, a, D.s[2] + r / nx2, D.s[1] + r % nx2, i);
}
AK_HD FcCell amr_whole_block(const FcDims &D, int m) {      // all coarse cells of block m (RefineFC / RepairAMRFC interior faces)
  FcCell d;
  d.m = m;
  for (int q = 0; q < 3; ++q) { d.lo[q] = D.cs[q]; d.n[q] = D.ce[q] - D.cs[q] + 1; }
  return d;
}

}  // namespace akb_smr
