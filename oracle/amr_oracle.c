/* amr_oracle.c -- TEST INFRASTRUCTURE ONLY (#included by mhd_oracle.c).
 * CPU restatement of the reference's refinement criteria (dynamic refinement, first part):
 *   src/mesh/refinement_criteria.cpp:177-238 (CheckMinMax), :243-287 (CheckSlope), :292-336 (CheckSecondDeriv),
 *   :341-386 (CheckLocation).  q = one variable of a cell-centred array (e.g. w0 with var = IDN for mhd_w_d);
 * flag[nmb] is updated in place exactly as the reference updates refine_flag (a block already flagged for
 * refinement by an earlier criterion is never turned into a derefinement).
 * Parity: tests/test_amr_cpu.py follows an unmodified reference AMR run dump by dump. */
enum { AMR_MIN_MAX = 0, AMR_SLOPE = 1, AMR_SECOND_DERIV = 2, AMR_LOCATION = 3 };

void oracle_amr_criterion(const ocfg *c, int method, const double *a, int nv, int var, double value_min, double value_max,
                          const double *loc4 /* x1, x2, x3, rad */, const double *mb_size, int *flag) {
  odims D = dims(c);
  const int multi_d = c->nx2 > 1, three_d = c->nx3 > 1;
  for (int m = 0; m < c->nmb; ++m) {
    if (method == AMR_LOCATION) {
      const double *sz = mb_size + (size_t)m * 9;
      const double x1 = loc4[0], x2 = loc4[1], x3 = loc4[2], rad = loc4[3];
      const double x1min = sz[0], x2min = sz[1], x3min = sz[2], x1max = sz[3], x2max = sz[4], x3max = sz[5];
      const int in1 = ((x1min < (x1 + rad)) && (x1min > (x1 - rad))) || ((x1max < (x1 + rad)) && (x1max > (x1 - rad))) || ((x1max > (x1 + rad)) && (x1min < (x1 - rad)));
      const int in2 = ((x2min < (x2 + rad)) && (x2min > (x2 - rad))) || ((x2max < (x2 + rad)) && (x2max > (x2 - rad))) || ((x2max > (x2 + rad)) && (x2min < (x2 - rad)));
      const int in3 = ((x3min < (x3 + rad)) && (x3min > (x3 - rad))) || ((x3max < (x3 + rad)) && (x3max > (x3 - rad))) || ((x3max > (x3 + rad)) && (x3min < (x3 - rad)));
      if (in1 && (!multi_d || in2) && (!three_d || in3)) flag[m] = 1;
      continue;
    }
    double qmax = -FLT_MAX, qmin = FLT_MAX;      /* Kokkos::Max / Kokkos::Min start from their identities */
    for (int k = D.ks; k <= D.ke; ++k)
      for (int j = D.js; j <= D.je; ++j)
        for (int i = D.is; i <= D.ie; ++i) {
          const double q = CC(a, nv, m, var, k, j, i);
          double v;
          if (method == AMR_MIN_MAX) { qmin = fmin(q, qmin); v = q; }
          else if (method == AMR_SLOPE) {
            double d2 = SQR(CC(a, nv, m, var, k, j, i + 1) - CC(a, nv, m, var, k, j, i - 1));
            if (multi_d) d2 += SQR(CC(a, nv, m, var, k, j + 1, i) - CC(a, nv, m, var, k, j - 1, i));
            if (three_d) d2 += SQR(CC(a, nv, m, var, k + 1, j, i) - CC(a, nv, m, var, k - 1, j, i));
            v = 0.5 * sqrt(d2) / q;
          } else {
            double d2q = CC(a, nv, m, var, k, j, i + 1) - 2.0 * q + CC(a, nv, m, var, k, j, i - 1);
            if (multi_d) d2q += (CC(a, nv, m, var, k, j + 1, i) - 2.0 * q + CC(a, nv, m, var, k, j - 1, i));
            if (three_d) d2q += (CC(a, nv, m, var, k + 1, j, i) - 2.0 * q + CC(a, nv, m, var, k - 1, j, i));
            v = fabs(d2q) / q;
          }
          qmax = fmax(v, qmax);
        }
    if (value_max < FLT_MAX) {
      if (qmax > value_max) flag[m] = 1;
      if ((qmax < value_max) && (flag[m] == 0)) flag[m] = -1;
    }
    if (method == AMR_MIN_MAX && value_min > -FLT_MAX) {
      if (qmin < value_min) flag[m] = 1;
      if ((qmin > value_min) && (flag[m] == 0)) flag[m] = -1;
    }
  }
}

/* ---------------------------------------------------------------------------------------------------------
 * Data movement of one AMR step on one rank (dynamic refinement, second part):
 * MeshRefinement::RedistAndRefineMeshBlocks steps 5-9 (src/mesh/mesh_refinement.cpp:523-603) and RepairAMRFC (:1189-1221).
 * Arrays hold max(old, new) blocks; the block-level copies are done IN PLACE in the reference's order (DerefineCC/FC
 * :696-812, CopyCC/FC :817-914 "move L" then "move R", CopyForRefinementCC/FC :919-1033), then RefineCC/FC (:1038-1184)
 * prolongates the blocks created by refinement from their coarse arrays.  flag = what happened to each OLD block
 * (+1 refined, -nleaf derefined, 0), lloc_old / lloc_new = [nmb][4] logical locations. */
static void amr_copy_block_cc(const ocfg *c, double *a, int nv, int dst, int src) {
  odims D = dims(c);
  memmove(a + (size_t)dst * nv * D.cc, a + (size_t)src * nv * D.cc, sizeof(double) * nv * D.cc);
}

void oracle_amr_redistribute(const ocfg *c_old, int new_nmb, double *u, double *cu, int nv, double *b1, double *b2, double *b3,
                             double *cb1, double *cb2, double *cb3, const int *flag, const int *newtoold, const int *oldtonew,
                             const int *lloc_old, const int *lloc_new) {
  ocfg cc_ = *c_old;
  const ocfg *c = &cc_;
  odims D = dims(c);
  sidx I = smr_indcs(c);
  cdims CD = smr_cdims(&I);
  cfdims CF = smr_cfdims(&I);
  const int old_nmb = c_old->nmb;
  const int nleaf = I.three_d ? 8 : (I.multi_d ? 4 : 2);
  const int ng = c->ng;
  int ndel = 0, nnew = 0;
  for (int m = 0; m < old_nmb; ++m) { if (flag[m] < -1) ndel++; if (flag[m] > 0) nnew++; }
  /* :519-521 */
  if (ndel > 0) oracle_smr_restrict_fc(c, b1, b2, b3, cb1, cb2, cb3);
  /* Step 5: derefine -- coarse arrays of the nleaf old siblings into the octants of the first sibling's fine array */
  for (int oldm = 0; oldm < old_nmb && ndel > 0; ++oldm) {
    if (flag[oldm] >= -1) continue;
    const int newm = oldtonew[oldm];
    if (oldm > 0 && oldtonew[oldm - 1] == newm) continue;
    for (int l = 0; l < nleaf; ++l) {
      if (oldm + l >= old_nmb) continue;
      const int *ll = lloc_old + 4 * (oldm + l);
      const int ox1 = ll[0] & 1, ox2 = ll[1] & 1, ox3 = ll[2] & 1;
      const int s = oldm + l, d = oldm;
      for (int v = 0; v < nv; ++v)
        for (int k = I.cks; k <= I.cke; ++k) for (int j = I.cjs; j <= I.cje; ++j) for (int i = I.cis; i <= I.cie; ++i)
          CC(u, nv, d, v, I.ks + ox3 * I.cnx3 * I.three_d + (k - I.cks), I.js + ox2 * I.cnx2 * I.multi_d + (j - I.cjs), I.is + ox1 * I.cnx1 + (i - I.cis)) = CCC(cu, nv, s, v, k, j, i);
      const int kd = I.ks + ox3 * I.cnx3 * I.three_d, jd = I.js + ox2 * I.cnx2 * I.multi_d, id = I.is + ox1 * I.cnx1;
      for (int k = I.cks; k <= I.cke; ++k) for (int j = I.cjs; j <= I.cje; ++j) for (int i = I.cis; i <= I.cie + 1; ++i)
        X1F(b1, d, kd + (k - I.cks), jd + (j - I.cjs), id + (i - I.cis)) = CX1F(cb1, s, k, j, i);
      for (int k = I.cks; k <= I.cke; ++k) for (int j = I.cjs; j <= I.cje + 1; ++j) for (int i = I.cis; i <= I.cie; ++i)
        X2F(b2, d, kd + (k - I.cks), jd + (j - I.cjs), id + (i - I.cis)) = CX2F(cb2, s, k, j, i);
      for (int k = I.cks; k <= I.cke + 1; ++k) for (int j = I.cjs; j <= I.cje; ++j) for (int i = I.cis; i <= I.cie; ++i)
        X3F(b3, d, kd + (k - I.cks), jd + (j - I.cjs), id + (i - I.cis)) = CX3F(cb3, s, k, j, i);
    }
  }
  /* Step 6: CopyCC / CopyFC -- move L (ascending), then move R (descending) */
  for (int pass = 0; pass < 2; ++pass) {
    for (int q = 0; q < old_nmb; ++q) {
      const int oldm = pass == 0 ? q : old_nmb - 1 - q;
      if (pass == 0 && oldm == 0) continue;
      const int newm = oldtonew[oldm];
      const int move = pass == 0 ? ((newm - oldm) < 0 && newm != oldtonew[oldm - 1]) : ((newm - oldm) > 0);
      if (!move) continue;
      amr_copy_block_cc(c, u, nv, newm, oldm);
      memmove(b1 + (size_t)newm * D.f1, b1 + (size_t)oldm * D.f1, sizeof(double) * D.f1);
      memmove(b2 + (size_t)newm * D.f2, b2 + (size_t)oldm * D.f2, sizeof(double) * D.f2);
      memmove(b3 + (size_t)newm * D.f3, b3 + (size_t)oldm * D.f3, sizeof(double) * D.f3);
    }
  }
  if (nnew == 0) return;
  /* Step 7: CopyForRefinementCC / FC -- the parent's octant (+ ng cells of the parent's resolution around it) into the
   * coarse array of every child */
  int il = I.cis - ng, iu = I.cie + ng, jl = I.cjs, ju = I.cje, kl = I.cks, ku = I.cke;
  if (I.multi_d) { jl -= ng; ju += ng; }
  if (I.three_d) { kl -= ng; ku += ng; }
  for (int newm = 0; newm < new_nmb; ++newm) {
    const int oldm = newtoold[newm];
    if (flag[oldm] <= 0) continue;
    const int msrc = oldtonew[oldm];
    const int *ll = lloc_new + 4 * newm;
    const int o1 = (ll[0] & 1) * I.cnx1, o2 = (ll[1] & 1) * I.cnx2 * I.multi_d, o3 = (ll[2] & 1) * I.cnx3 * I.three_d;
    for (int v = 0; v < nv; ++v)
      for (int k = kl; k <= ku; ++k) for (int j = jl; j <= ju; ++j) for (int i = il; i <= iu; ++i)
        CCC(cu, nv, newm, v, k, j, i) = CC(u, nv, msrc, v, k + o3, j + o2, i + o1);
    for (int k = kl; k <= ku; ++k) for (int j = jl; j <= ju; ++j) for (int i = il; i <= iu + 1; ++i)
      CX1F(cb1, newm, k, j, i) = X1F(b1, msrc, k + o3, j + o2, i + o1);
    for (int k = kl; k <= ku; ++k) for (int j = jl; j <= ju + 1; ++j) for (int i = il; i <= iu; ++i)
      CX2F(cb2, newm, k, j, i) = X2F(b2, msrc, k + o3, j + o2, i + o1);
    for (int k = kl; k <= ku + 1; ++k) for (int j = jl; j <= ju; ++j) for (int i = il; i <= iu; ++i)
      CX3F(cb3, newm, k, j, i) = X3F(b3, msrc, k + o3, j + o2, i + o1);
  }
  /* Step 9: RefineCC (mesh/prolongation.hpp:21-60) and RefineFC (shared faces, then the Toth-Roe interior faces) */
  for (int m = 0; m < new_nmb; ++m) {
    if (flag[newtoold[m]] <= 0) continue;
    for (int v = 0; v < nv; ++v)
      for (int k = I.cks; k <= I.cke; ++k) for (int j = I.cjs; j <= I.cje; ++j) for (int i = I.cis; i <= I.cie; ++i) {
        const int fi = 2 * i - I.cis, fj = 2 * j - I.cjs, fk = 2 * k - I.cks;
        const double q = CCC(cu, nv, m, v, k, j, i);
        double dl = q - CCC(cu, nv, m, v, k, j, i - 1), dr = CCC(cu, nv, m, v, k, j, i + 1) - q;
        const double dvar1 = 0.125 * (smr_sign(dl) + smr_sign(dr)) * fmin(fabs(dl), fabs(dr));
        double dvar2 = 0.0, dvar3 = 0.0;
        if (I.multi_d) { dl = q - CCC(cu, nv, m, v, k, j - 1, i); dr = CCC(cu, nv, m, v, k, j + 1, i) - q; dvar2 = 0.125 * (smr_sign(dl) + smr_sign(dr)) * fmin(fabs(dl), fabs(dr)); }
        if (I.three_d) { dl = q - CCC(cu, nv, m, v, k - 1, j, i); dr = CCC(cu, nv, m, v, k + 1, j, i) - q; dvar3 = 0.125 * (smr_sign(dl) + smr_sign(dr)) * fmin(fabs(dl), fabs(dr)); }
        CC(u, nv, m, v, fk, fj, fi) = q - dvar1 - dvar2 - dvar3;
        CC(u, nv, m, v, fk, fj, fi + 1) = q + dvar1 - dvar2 - dvar3;
        if (I.multi_d) { CC(u, nv, m, v, fk, fj + 1, fi) = q - dvar1 + dvar2 - dvar3; CC(u, nv, m, v, fk, fj + 1, fi + 1) = q + dvar1 + dvar2 - dvar3; }
        if (I.three_d) {
          CC(u, nv, m, v, fk + 1, fj, fi) = q - dvar1 - dvar2 + dvar3; CC(u, nv, m, v, fk + 1, fj, fi + 1) = q + dvar1 - dvar2 + dvar3;
          CC(u, nv, m, v, fk + 1, fj + 1, fi) = q - dvar1 + dvar2 + dvar3; CC(u, nv, m, v, fk + 1, fj + 1, fi + 1) = q + dvar1 + dvar2 + dvar3;
        }
      }
  }
}

/* RefineFC (mesh_refinement.cpp:1107-1184; operators mesh/prolongation.hpp) for the blocks whose flag is set, and the
 * interior-face part alone = RepairAMRFC (:1189-1221).  what: 1 = shared faces + interior (RefineFC), 2 = interior only */
void oracle_amr_refine_fc(const ocfg *c, double *b1, double *b2, double *b3, const double *cb1, const double *cb2, const double *cb3,
                          const int *blockflag, int what) {
  odims D = dims(c);
  sidx I = smr_indcs(c);
  cfdims CF = smr_cfdims(&I);
  const int multi_d = I.multi_d, three_d = I.three_d;
  for (int m = 0; m < c->nmb; ++m) {
    if (!blockflag[m]) continue;
    if (what == 1) {
      for (int k = I.cks; k <= I.cke; ++k) for (int j = I.cjs; j <= I.cje; ++j) for (int i = I.cis; i <= I.cie + 1; ++i) {
        const int fi = (i - I.cis) * 2 + I.is, fj = multi_d ? ((j - I.cjs) * 2 + I.js) : j, fk = three_d ? ((k - I.cks) * 2 + I.ks) : k;
        const double q = CX1F(cb1, m, k, j, i);
        double dv2 = 0.0, dv3 = 0.0;
        if (multi_d) dv2 = fc_mm(q - CX1F(cb1, m, k, j - 1, i), CX1F(cb1, m, k, j + 1, i) - q);
        if (three_d) dv3 = fc_mm(q - CX1F(cb1, m, k - 1, j, i), CX1F(cb1, m, k + 1, j, i) - q);
        X1F(b1, m, fk, fj, fi) = q - dv2 - dv3;
        if (multi_d) X1F(b1, m, fk, fj + 1, fi) = q + dv2 - dv3;
        if (three_d) { X1F(b1, m, fk + 1, fj, fi) = q - dv2 + dv3; X1F(b1, m, fk + 1, fj + 1, fi) = q + dv2 + dv3; }
      }
      for (int k = I.cks; k <= I.cke; ++k) for (int j = I.cjs; j <= I.cje + 1; ++j) for (int i = I.cis; i <= I.cie; ++i) {
        const int fi = (i - I.cis) * 2 + I.is, fj = multi_d ? ((j - I.cjs) * 2 + I.js) : j, fk = three_d ? ((k - I.cks) * 2 + I.ks) : k;
        const double q = CX2F(cb2, m, k, j, i);
        const double dv1 = fc_mm(q - CX2F(cb2, m, k, j, i - 1), CX2F(cb2, m, k, j, i + 1) - q);
        double dv3 = 0.0;
        if (three_d) dv3 = fc_mm(q - CX2F(cb2, m, k - 1, j, i), CX2F(cb2, m, k + 1, j, i) - q);
        X2F(b2, m, fk, fj, fi) = q - dv1 - dv3;
        X2F(b2, m, fk, fj, fi + 1) = q + dv1 - dv3;
        if (three_d) { X2F(b2, m, fk + 1, fj, fi) = q - dv1 + dv3; X2F(b2, m, fk + 1, fj, fi + 1) = q + dv1 + dv3; }
      }
      for (int k = I.cks; k <= I.cke + 1; ++k) for (int j = I.cjs; j <= I.cje; ++j) for (int i = I.cis; i <= I.cie; ++i) {
        const int fi = (i - I.cis) * 2 + I.is, fj = multi_d ? ((j - I.cjs) * 2 + I.js) : j, fk = three_d ? ((k - I.cks) * 2 + I.ks) : k;
        const double q = CX3F(cb3, m, k, j, i);
        const double dv1 = fc_mm(q - CX3F(cb3, m, k, j, i - 1), CX3F(cb3, m, k, j, i + 1) - q);
        double dv2 = 0.0;
        if (multi_d) dv2 = fc_mm(q - CX3F(cb3, m, k, j - 1, i), CX3F(cb3, m, k, j + 1, i) - q);
        X3F(b3, m, fk, fj, fi) = q - dv1 - dv2;
        X3F(b3, m, fk, fj, fi + 1) = q + dv1 - dv2;
        if (multi_d) { X3F(b3, m, fk, fj + 1, fi) = q - dv1 + dv2; X3F(b3, m, fk, fj + 1, fi + 1) = q + dv1 + dv2; }
      }
    }
    for (int k = I.cks; k <= I.cke; ++k) for (int j = I.cjs; j <= I.cje; ++j) for (int i = I.cis; i <= I.cie; ++i) {
      const int fi = (i - I.cis) * 2 + I.is, fj = (j - I.cjs) * 2 + I.js, fk = (k - I.cks) * 2 + I.ks;
      if (!multi_d) {
        X1F(b1, m, fk, fj, fi + 1) = 0.5 * (X1F(b1, m, fk, fj, fi) + X1F(b1, m, fk, fj, fi + 2));
      } else if (three_d) {
        double Uxx = 0.0, Vyy = 0.0, Wzz = 0.0, Uxyz = 0.0, Vxyz = 0.0, Wxyz = 0.0;
        for (int jj = 0; jj < 2; jj++) {
          int jsgn = 2 * jj - 1, fjj = fj + jj, fjp = fj + 2 * jj;
          for (int ii = 0; ii < 2; ii++) {
            int isgn = 2 * ii - 1, fii = fi + ii, fip = fi + 2 * ii;
            Uxx += isgn * (jsgn * (X2F(b2, m, fk, fjp, fii) + X2F(b2, m, fk + 1, fjp, fii)) + (X3F(b3, m, fk + 2, fjj, fii) - X3F(b3, m, fk, fjj, fii)));
            Vyy += jsgn * ((X3F(b3, m, fk + 2, fjj, fii) - X3F(b3, m, fk, fjj, fii)) + isgn * (X1F(b1, m, fk, fjj, fip) + X1F(b1, m, fk + 1, fjj, fip)));
            Wzz += isgn * (X1F(b1, m, fk + 1, fjj, fip) - X1F(b1, m, fk, fjj, fip)) + jsgn * (X2F(b2, m, fk + 1, fjp, fii) - X2F(b2, m, fk, fjp, fii));
            Uxyz += isgn * jsgn * (X1F(b1, m, fk + 1, fjj, fip) - X1F(b1, m, fk, fjj, fip));
            Vxyz += isgn * jsgn * (X2F(b2, m, fk + 1, fjp, fii) - X2F(b2, m, fk, fjp, fii));
            Wxyz += isgn * jsgn * (X3F(b3, m, fk + 2, fjj, fii) - X3F(b3, m, fk, fjj, fii));
          }
        }
        Uxx *= 0.125; Vyy *= 0.125; Wzz *= 0.125;
        Uxyz *= 0.0625; Vxyz *= 0.0625; Wxyz *= 0.0625;
        X1F(b1, m, fk, fj, fi + 1) = 0.5 * (X1F(b1, m, fk, fj, fi) + X1F(b1, m, fk, fj, fi + 2)) + Uxx - Vxyz - Wxyz;
        X1F(b1, m, fk, fj + 1, fi + 1) = 0.5 * (X1F(b1, m, fk, fj + 1, fi) + X1F(b1, m, fk, fj + 1, fi + 2)) + Uxx - Vxyz + Wxyz;
        X1F(b1, m, fk + 1, fj, fi + 1) = 0.5 * (X1F(b1, m, fk + 1, fj, fi) + X1F(b1, m, fk + 1, fj, fi + 2)) + Uxx + Vxyz - Wxyz;
        X1F(b1, m, fk + 1, fj + 1, fi + 1) = 0.5 * (X1F(b1, m, fk + 1, fj + 1, fi) + X1F(b1, m, fk + 1, fj + 1, fi + 2)) + Uxx + Vxyz + Wxyz;
        X2F(b2, m, fk, fj + 1, fi) = 0.5 * (X2F(b2, m, fk, fj, fi) + X2F(b2, m, fk, fj + 2, fi)) + Vyy - Uxyz - Wxyz;
        X2F(b2, m, fk, fj + 1, fi + 1) = 0.5 * (X2F(b2, m, fk, fj, fi + 1) + X2F(b2, m, fk, fj + 2, fi + 1)) + Vyy - Uxyz + Wxyz;
        X2F(b2, m, fk + 1, fj + 1, fi) = 0.5 * (X2F(b2, m, fk + 1, fj, fi) + X2F(b2, m, fk + 1, fj + 2, fi)) + Vyy + Uxyz - Wxyz;
        X2F(b2, m, fk + 1, fj + 1, fi + 1) = 0.5 * (X2F(b2, m, fk + 1, fj, fi + 1) + X2F(b2, m, fk + 1, fj + 2, fi + 1)) + Vyy + Uxyz + Wxyz;
        X3F(b3, m, fk + 1, fj, fi) = 0.5 * (X3F(b3, m, fk + 2, fj, fi) + X3F(b3, m, fk, fj, fi)) + Wzz - Uxyz - Vxyz;
        X3F(b3, m, fk + 1, fj, fi + 1) = 0.5 * (X3F(b3, m, fk + 2, fj, fi + 1) + X3F(b3, m, fk, fj, fi + 1)) + Wzz - Uxyz + Vxyz;
        X3F(b3, m, fk + 1, fj + 1, fi) = 0.5 * (X3F(b3, m, fk + 2, fj + 1, fi) + X3F(b3, m, fk, fj + 1, fi)) + Wzz + Uxyz - Vxyz;
        X3F(b3, m, fk + 1, fj + 1, fi + 1) = 0.5 * (X3F(b3, m, fk + 2, fj + 1, fi + 1) + X3F(b3, m, fk, fj + 1, fi + 1)) + Wzz + Uxyz + Vxyz;
      } else {
        const double tmp1 = 0.25 * (X2F(b2, m, fk, fj + 2, fi + 1) - X2F(b2, m, fk, fj, fi + 1) - X2F(b2, m, fk, fj + 2, fi) + X2F(b2, m, fk, fj, fi));
        const double tmp2 = 0.25 * (X1F(b1, m, fk, fj, fi) - X1F(b1, m, fk, fj, fi + 2) - X1F(b1, m, fk, fj + 1, fi) + X1F(b1, m, fk, fj + 1, fi + 2));
        X1F(b1, m, fk, fj, fi + 1) = 0.5 * (X1F(b1, m, fk, fj, fi) + X1F(b1, m, fk, fj, fi + 2)) + tmp1;
        X1F(b1, m, fk, fj + 1, fi + 1) = 0.5 * (X1F(b1, m, fk, fj + 1, fi) + X1F(b1, m, fk, fj + 1, fi + 2)) + tmp1;
        X2F(b2, m, fk, fj + 1, fi) = 0.5 * (X2F(b2, m, fk, fj, fi) + X2F(b2, m, fk, fj + 2, fi)) + tmp2;
        X2F(b2, m, fk, fj + 1, fi + 1) = 0.5 * (X2F(b2, m, fk, fj, fi + 1) + X2F(b2, m, fk, fj + 2, fi + 1)) + tmp2;
      }
    }
  }
}
