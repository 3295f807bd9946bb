/* smr_fc_oracle.c -- TEST INFRASTRUCTURE ONLY (#included by mhd_oracle.c after smr_oracle.c).
 * CPU restatement of the reference's fine/coarse boundary path for FACE-CENTRED fields on a statically
 * refined mesh (MHD): RestrictB -> SendB/RecvB across levels -> FillCoarseInBndryFC ->
 * ProlongateFC (shared faces, then the divergence-preserving interior faces of Toth & Roe).
 *   index tables   src/bvals/buffs_fc.cpp:29-326 (send), :388-790 (recv, prolongation)
 *   RestrictFC     src/mesh/mesh_refinement.cpp:1290-1383
 *   pack / unpack  src/bvals/bvals_fc.cpp:63-283, :289-436 (active faces are never overwritten)
 *   prolongation   src/bvals/prolongation.cpp:28-357 (helpers, ownership rule), :556-785
 * Loops in Kokkos::Serial order.  Parity: see tests/test_smr_cpu.py. */

typedef struct { int n1, n2, n3; size_t f1, f2, f3; } cfdims;
static cfdims smr_cfdims(const sidx *I) {
  cdims c = smr_cdims(I);
  cfdims d;
  d.n1 = c.n1; d.n2 = c.n2; d.n3 = c.n3;
  d.f1 = (size_t)c.n3 * c.n2 * (c.n1 + 1); d.f2 = (size_t)c.n3 * (c.n2 + 1) * c.n1; d.f3 = (size_t)(c.n3 + 1) * c.n2 * c.n1;
  return d;
}
size_t oracle_smr_coarse_faces(const ocfg *c, int v) {
  sidx I = smr_indcs(c);
  cfdims d = smr_cfdims(&I);
  return v == 0 ? d.f1 : v == 1 ? d.f2 : d.f3;
}
/* coarse face arrays (same shapes as the fine ones with the coarse extents) */
#define CX1F(a, m, k, j, i) (a)[(((size_t)(m) * CF.n3 + (k)) * CF.n2 + (j)) * (CF.n1 + 1) + (i)]
#define CX2F(a, m, k, j, i) (a)[(((size_t)(m) * CF.n3 + (k)) * (CF.n2 + 1) + (j)) * CF.n1 + (i)]
#define CX3F(a, m, k, j, i) (a)[(((size_t)(m) * (CF.n3 + 1) + (k)) * CF.n2 + (j)) * CF.n1 + (i)]

/* element of component v, fine or coarse */
static double *fc_at(const odims *Dp, const cfdims *CFp, double *const *b, double *const *cb, int coarse, int v, int m, int k, int j, int i) {
  odims D = *Dp;
  cfdims CF = *CFp;
  if (!coarse) return v == 0 ? &X1F(b[0], m, k, j, i) : v == 1 ? &X2F(b[1], m, k, j, i) : &X3F(b[2], m, k, j, i);
  return v == 0 ? &CX1F(cb[0], m, k, j, i) : v == 1 ? &CX2F(cb[1], m, k, j, i) : &CX3F(cb[2], m, k, j, i);
}

/* buffs_fc.cpp: one function for the send (29-326) and receive (388-790) tables.  v = component, kinds as in
 * smr_oracle.c; d-th direction staggered when v == d. */
static sbox smr_fc_box(const sidx *I, const int o[3], int f1, int f2, int kind, int v, int send) {
  const int ng = I->ng, ng1 = ng - 1;
  const int s_[3] = {I->is, I->js, I->ks}, e_[3] = {I->ie, I->je, I->ke};
  const int cs_[3] = {I->cis, I->cjs, I->cks}, ce_[3] = {I->cie, I->cje, I->cke};
  const int cnx[3] = {I->cnx1, I->cnx2, I->cnx3};
  const int act[3] = {1, I->nx2 > 1, I->nx3 > 1};
  int lo[3], hi[3];
  for (int d = 0; d < 3; ++d) {
    const int st = (v == d);
    const int sub = (d == 0) ? f1 : (d == 1) ? ((o[0] != 0) ? f1 : f2) : ((o[0] != 0 && o[1] != 0) ? f1 : f2);
    const int others = (o[(d + 1) % 3] != 0) || (o[(d + 2) % 3] != 0);
    if (send) {
      if (kind == SMR_SAME || kind == SMR_COAR) {
        const int s = (kind == SMR_COAR) ? cs_[d] : s_[d], e = (kind == SMR_COAR) ? ce_[d] : e_[d];
        if (o[d] == 0) { lo[d] = s; hi[d] = e + st; }
        else if (o[d] > 0) { lo[d] = e - ng1; hi[d] = e; }
        else { lo[d] = s + st; hi[d] = s + ng1 + st; }
        if (st && others) { if (o[d] > 0) hi[d]++; if (o[d] < 0) lo[d]--; }      /* multilevel: overlapping faces on edges / corners */
      } else {   /* SMR_FINE */
        if (o[d] == 0) {
          lo[d] = s_[d]; hi[d] = e_[d] + st;
          if (d == 0 || act[d]) { if (sub == 1) lo[d] += cnx[d] - ng; else hi[d] -= cnx[d] - ng; }
        } else if (o[d] > 0) { lo[d] = e_[d] - ng1; hi[d] = e_[d] + st; }
        else { lo[d] = s_[d]; hi[d] = s_[d] + ng1 + st; }
      }
    } else {
      if (kind == SMR_SAME || kind == SMR_FINE) {
        if (o[d] == 0) {
          lo[d] = s_[d]; hi[d] = e_[d] + st;
          if (kind == SMR_FINE && (d == 0 || act[d])) { if (sub == 1) lo[d] += cnx[d]; else hi[d] -= cnx[d]; }
        } else if (o[d] > 0) { lo[d] = e_[d] + 1 + st; hi[d] = e_[d] + ng + st; }
        else { lo[d] = s_[d] - ng; hi[d] = s_[d] - 1; }
        if (st && others) { if (o[d] > 0) lo[d]--; if (o[d] < 0) hi[d]++; }
      } else if (kind == SMR_COAR) {
        if (o[d] == 0) {
          lo[d] = cs_[d]; hi[d] = ce_[d] + st;
          if (d == 0 || act[d]) { if (sub == 0) hi[d] += ng; else lo[d] -= ng; }
        } else if (o[d] > 0) { lo[d] = ce_[d] + 1; hi[d] = ce_[d] + ng + st; }
        else { lo[d] = cs_[d] - ng; hi[d] = cs_[d] - 1 + st; }
      } else {   /* SMR_PROL */
        const int cn = ng / 2;
        if (o[d] == 0) {
          lo[d] = cs_[d]; hi[d] = ce_[d] + st;
          if (d == 0 || act[d]) { if (sub == 0) hi[d] += cn; else lo[d] -= cn; }
        } else if (o[d] > 0) { lo[d] = ce_[d] + 1 + st; hi[d] = ce_[d] + cn + st; }
        else { lo[d] = cs_[d] - cn; hi[d] = cs_[d] - 1; }
      }
    }
  }
  sbox b = {lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]};
  return b;
}

/* mesh_refinement.cpp:1290-1383 */
void oracle_smr_restrict_fc(const ocfg *c, const double *b1, const double *b2, const double *b3, double *cb1, double *cb2, double *cb3) {
  odims D = dims(c);
  sidx I = smr_indcs(c);
  cfdims CF = smr_cfdims(&I);
  const int cis = I.cis, cie = I.cie, cjs = I.cjs, cje = I.cje, cks = I.cks, cke = I.cke;
  for (int m = 0; m < c->nmb; ++m) {
    if (!I.multi_d) {
      for (int i = cis; i <= cie; ++i) {
        int fi = 2 * i - cis;
        CX1F(cb1, m, cks, cjs, i) = X1F(b1, m, cks, cjs, fi);
        if (i == cie) CX1F(cb1, m, cks, cjs, i + 1) = X1F(b1, m, cks, cjs, fi + 2);
        double b2c = 0.5 * (X2F(b2, m, cks, cjs, fi) + X2F(b2, m, cks, cjs, fi + 1));
        CX2F(cb2, m, cks, cjs, i) = b2c; CX2F(cb2, m, cks, cjs + 1, i) = b2c;
        double b3c = 0.5 * (X3F(b3, m, cks, cjs, fi) + X3F(b3, m, cks, cjs, fi + 1));
        CX3F(cb3, m, cks, cjs, i) = b3c; CX3F(cb3, m, cks + 1, cjs, i) = b3c;
      }
    } else if (!I.three_d) {
      for (int j = cjs; j <= cje; ++j)
        for (int i = cis; i <= cie; ++i) {
          int fi = 2 * i - cis, fj = 2 * j - cjs;
          CX1F(cb1, m, cks, j, i) = 0.5 * (X1F(b1, m, cks, fj, fi) + X1F(b1, m, cks, fj + 1, fi));
          if (i == cie) CX1F(cb1, m, cks, j, i + 1) = 0.5 * (X1F(b1, m, cks, fj, fi + 2) + X1F(b1, m, cks, fj + 1, fi + 2));
          CX2F(cb2, m, cks, j, i) = 0.5 * (X2F(b2, m, cks, fj, fi) + X2F(b2, m, cks, fj, fi + 1));
          if (j == cje) CX2F(cb2, m, cks, j + 1, i) = 0.5 * (X2F(b2, m, cks, fj + 2, fi) + X2F(b2, m, cks, fj + 2, fi + 1));
          double b3c = 0.25 * (X3F(b3, m, cks, fj, fi) + X3F(b3, m, cks, fj, fi + 1) + X3F(b3, m, cks, fj + 1, fi) + X3F(b3, m, cks, fj + 1, fi + 1));
          CX3F(cb3, m, cks, j, i) = b3c; CX3F(cb3, m, cks + 1, j, i) = b3c;
        }
    } else {
      for (int k = cks; k <= cke; ++k)
        for (int j = cjs; j <= cje; ++j)
          for (int i = cis; i <= cie; ++i) {
            int fi = 2 * i - cis, fj = 2 * j - cjs, fk = 2 * k - cks;
            CX1F(cb1, m, k, j, i) = 0.25 * (X1F(b1, m, fk, fj, fi) + X1F(b1, m, fk, fj + 1, fi) + X1F(b1, m, fk + 1, fj, fi) + X1F(b1, m, fk + 1, fj + 1, fi));
            if (i == cie)
              CX1F(cb1, m, k, j, i + 1) = 0.25 * (X1F(b1, m, fk, fj, fi + 2) + X1F(b1, m, fk, fj + 1, fi + 2) + X1F(b1, m, fk + 1, fj, fi + 2) + X1F(b1, m, fk + 1, fj + 1, fi + 2));
            CX2F(cb2, m, k, j, i) = 0.25 * (X2F(b2, m, fk, fj, fi) + X2F(b2, m, fk, fj, fi + 1) + X2F(b2, m, fk + 1, fj, fi) + X2F(b2, m, fk + 1, fj, fi + 1));
            if (j == cje)
              CX2F(cb2, m, k, j + 1, i) = 0.25 * (X2F(b2, m, fk, fj + 2, fi) + X2F(b2, m, fk, fj + 2, fi + 1) + X2F(b2, m, fk + 1, fj + 2, fi) + X2F(b2, m, fk + 1, fj + 2, fi + 1));
            CX3F(cb3, m, k, j, i) = 0.25 * (X3F(b3, m, fk, fj, fi) + X3F(b3, m, fk, fj, fi + 1) + X3F(b3, m, fk, fj + 1, fi) + X3F(b3, m, fk, fj + 1, fi + 1));
            if (k == cke)
              CX3F(cb3, m, k + 1, j, i) = 0.25 * (X3F(b3, m, fk + 2, fj, fi) + X3F(b3, m, fk + 2, fj, fi + 1) + X3F(b3, m, fk + 2, fj + 1, fi) + X3F(b3, m, fk + 2, fj + 1, fi + 1));
          }
    }
  }
}

/* prolongation.cpp:75-90 */
static int fc_is_active(const sidx *I, int v, int k, int j, int i) {
  if (v == 0) return i >= I->is && i <= I->ie + 1 && j >= I->js && j <= I->je && k >= I->ks && k <= I->ke;
  if (v == 1) return i >= I->is && i <= I->ie && j >= I->js && j <= I->je + 1 && k >= I->ks && k <= I->ke;
  return i >= I->is && i <= I->ie && j >= I->js && j <= I->je && k >= I->ks && k <= I->ke + 1;
}

/* bvals_fc.cpp:63-283 + :289-436 */
void oracle_smr_halo_fc(const ocfg *c, double *b1, double *b2, double *b3, double *cb1, double *cb2, double *cb3,
                        const int *nghbr, const int *mblev) {
  odims D = dims(c);
  sidx I = smr_indcs(c);
  cfdims CF = smr_cfdims(&I);
  double *b[3] = {b1, b2, b3}, *cb[3] = {cb1, cb2, cb3};
  const int nmb = c->nmb;
  double **rbuf = calloc((size_t)nmb * 56 * 3, sizeof(double *));
  for (int m = 0; m < nmb; ++m)
    for (int n = 0; n < 56; ++n) {
      const int *e = nghbr + ((size_t)m * 56 + n) * 4;
      if (e[0] < 0) continue;
      int o[3], f1, f2;
      smr_slot_decode(n, &o[0], &o[1], &o[2], &f1, &f2);
      const int kind = (e[1] < mblev[m]) ? SMR_COAR : (e[1] == mblev[m]) ? SMR_SAME : SMR_FINE;
      for (int v = 0; v < 3; ++v) {
        sbox x = smr_fc_box(&I, o, f1, f2, kind, v, 1);
        const int ni = x.ie - x.is + 1, nj = x.je - x.js + 1, nk = x.ke - x.ks + 1;
        double *buf = malloc(sizeof(double) * (size_t)(ni > 0 ? ni : 0) * (nj > 0 ? nj : 0) * (nk > 0 ? nk : 0) + 8);
        for (int k = x.ks; k <= x.ke; ++k)
          for (int j = x.js; j <= x.je; ++j)
            for (int i = x.is; i <= x.ie; ++i)
              buf[(i - x.is) + ni * ((j - x.js) + nj * (k - x.ks))] = *fc_at(&D, &CF, b, cb, kind == SMR_COAR, v, m, k, j, i);
        rbuf[((size_t)e[0] * 56 + e[3]) * 3 + v] = buf;
      }
    }
  for (int m = 0; m < nmb; ++m)
    for (int v = 0; v < 3; ++v)
      for (int n = 0; n < 56; ++n) {
        const int *e = nghbr + ((size_t)m * 56 + n) * 4;
        if (e[0] < 0) continue;
        int o[3], f1, f2;
        smr_slot_decode(n, &o[0], &o[1], &o[2], &f1, &f2);
        const int kind = (e[1] < mblev[m]) ? SMR_COAR : (e[1] == mblev[m]) ? SMR_SAME : SMR_FINE;
        sbox x = smr_fc_box(&I, o, f1, f2, kind, v, 0);
        const int ni = x.ie - x.is + 1, nj = x.je - x.js + 1;
        const double *buf = rbuf[((size_t)m * 56 + n) * 3 + v];
        for (int k = x.ks; k <= x.ke; ++k)
          for (int j = x.js; j <= x.je; ++j)
            for (int i = x.is; i <= x.ie; ++i) {
              if (kind != SMR_COAR && fc_is_active(&I, v, k, j, i)) continue;      /* bvals_fc.cpp:372-374 */
              *fc_at(&D, &CF, b, cb, kind == SMR_COAR, v, m, k, j, i) = buf[(i - x.is) + ni * ((j - x.js) + nj * (k - x.ks))];
            }
      }
  for (size_t q = 0; q < (size_t)nmb * 56 * 3; ++q) free(rbuf[q]);
  free(rbuf);
}

/* prolongation.cpp:28-50: the reference's own slot -> offset search, INCLUDING its aliasing of edge slots
 * (it tries every (n1,n2) for every offset, so e.g. slot 34 resolves to (-1,0,-1)); only the ownership test
 * below consumes it, and that test is true for pure face offsets only */
static void fc_ref_offset(int n, int *ox1, int *ox2, int *ox3) {
  *ox1 = *ox2 = *ox3 = 0;
  for (int iz = -1; iz <= 1; ++iz)
    for (int iy = -1; iy <= 1; ++iy)
      for (int ix = -1; ix <= 1; ++ix) {
        if (!ix && !iy && !iz) continue;
        for (int n1 = 0; n1 <= 1; ++n1)
          for (int n2 = 0; n2 <= 1; ++n2)
            if (oracle_neighbor_index(ix, iy, iz, n1, n2) == n) { *ox1 = ix; *ox2 = iy; *ox3 = iz; return; }
      }
}
static int fc_max_lev_at(const int *nghbr, int m, int ox1, int ox2, int ox3) {   /* prolongation.cpp:53-68 */
  int mx = -1;
  for (int n1 = 0; n1 <= 1; ++n1)
    for (int n2 = 0; n2 <= 1; ++n2) {
      const int idx = oracle_neighbor_index(ox1, ox2, ox3, n1, n2);
      if (idx >= 0 && idx < 56 && nghbr[((size_t)m * 56 + idx) * 4] >= 0) {
        const int l = nghbr[((size_t)m * 56 + idx) * 4 + 1];
        if (l > mx) mx = l;
      }
    }
  return mx;
}
/* prolongation.cpp:92-133: ghost faces are always written; an ACTIVE face only on the fine/coarse interface
 * normal to it, by the face slot of that side, when every neighbour across it is coarser */
static int fc_can_prolong(const sidx *I, const int *nghbr, int m, int v, int k, int j, int i, int ox1, int ox2, int ox3, int my_lev) {
  if (!fc_is_active(I, v, k, j, i)) return 1;
  int nox = 0;
  if (v == 0) {
    if (i == I->is) nox = -1; else if (i == I->ie + 1) nox = 1; else return 0;
    return ox1 == nox && ox2 == 0 && ox3 == 0 && fc_max_lev_at(nghbr, m, nox, 0, 0) < my_lev;
  } else if (v == 1) {
    if (j == I->js) nox = -1; else if (j == I->je + 1) nox = 1; else return 0;
    return ox1 == 0 && ox2 == nox && ox3 == 0 && fc_max_lev_at(nghbr, m, 0, nox, 0) < my_lev;
  }
  if (k == I->ks) nox = -1; else if (k == I->ke + 1) nox = 1; else return 0;
  return ox1 == 0 && ox2 == 0 && ox3 == nox && fc_max_lev_at(nghbr, m, 0, 0, nox) < my_lev;
}

static double fc_mm(double dl, double dr) { return 0.125 * (smr_sign(dl) + smr_sign(dr)) * fmin(fabs(dl), fabs(dr)); }

/* prolongation.cpp:556-785.  steps: 1 = FillCoarseInBndryFC, 2 = ProlongateFC, 3 = both */
void oracle_smr_prolongate_fc_steps(const ocfg *c, double *b1, double *b2, double *b3, double *cb1, double *cb2, double *cb3,
                                    const int *nghbr, const int *mblev, int steps) {
  odims D = dims(c);
  sidx I = smr_indcs(c);
  cfdims CF = smr_cfdims(&I);
  const int multi_d = I.multi_d, three_d = I.three_d;
  if (multi_d && (steps & 1))
    for (int m = 0; m < c->nmb; ++m)
      for (int n = 0; n < 56; ++n) {
        const int *e = nghbr + ((size_t)m * 56 + n) * 4;
        if (e[0] < 0 || e[1] != mblev[m]) continue;
        int o[3], f1, f2;
        smr_slot_decode(n, &o[0], &o[1], &o[2], &f1, &f2);
        for (int v = 0; v < 3; ++v) {
          sbox r = smr_fc_box(&I, o, 0, 0, SMR_SAME, v, 0);
          const int il = (r.is + I.cis) / 2, iu = (r.ie + I.cis) / 2, jl = (r.js + I.cjs) / 2, ju = (r.je + I.cjs) / 2;
          const int kl = (r.ks + I.cks) / 2, ku = (r.ke + I.cks) / 2;
          for (int k = kl; k <= ku; ++k)
            for (int j = jl; j <= ju; ++j)
              for (int i = il; i <= iu; ++i) {
                const int fk = (k - I.cks) * 2 + I.ks, fj = (j - I.cjs) * 2 + I.js, fi = (i - I.cis) * 2 + I.is;
                if (!three_d) {
                  if (v == 0) CX1F(cb1, m, kl, j, i) = 0.5 * (X1F(b1, m, kl, fj, fi) + X1F(b1, m, kl, fj + 1, fi));
                  else if (v == 1) CX2F(cb2, m, kl, j, i) = 0.5 * (X2F(b2, m, kl, fj, fi) + X2F(b2, m, kl, fj, fi + 1));
                  else {
                    double b3c = 0.25 * (X3F(b3, m, kl, fj, fi) + X3F(b3, m, kl, fj, fi + 1) + X3F(b3, m, kl, fj + 1, fi) + X3F(b3, m, kl, fj + 1, fi + 1));
                    CX3F(cb3, m, kl, j, i) = b3c; CX3F(cb3, m, kl + 1, j, i) = b3c;
                  }
                } else {
                  if (v == 0) CX1F(cb1, m, k, j, i) = 0.25 * (X1F(b1, m, fk, fj, fi) + X1F(b1, m, fk, fj + 1, fi) + X1F(b1, m, fk + 1, fj, fi) + X1F(b1, m, fk + 1, fj + 1, fi));
                  else if (v == 1) CX2F(cb2, m, k, j, i) = 0.25 * (X2F(b2, m, fk, fj, fi) + X2F(b2, m, fk, fj, fi + 1) + X2F(b2, m, fk + 1, fj, fi) + X2F(b2, m, fk + 1, fj, fi + 1));
                  else CX3F(cb3, m, k, j, i) = 0.25 * (X3F(b3, m, fk, fj, fi) + X3F(b3, m, fk, fj, fi + 1) + X3F(b3, m, fk, fj + 1, fi) + X3F(b3, m, fk, fj + 1, fi + 1));
                }
              }
        }
      }
  if (!(steps & 2)) return;
#define FC_STORE(v_, arr_, K_, J_, I_, val_) do { if (fc_can_prolong(&I, nghbr, m, v_, K_, J_, I_, rx1, rx2, rx3, my_lev)) arr_ = (val_); } while (0)
  /* shared faces */
  for (int m = 0; m < c->nmb; ++m)
    for (int n = 0; n < 56; ++n) {
      const int *e = nghbr + ((size_t)m * 56 + n) * 4;
      if (e[0] < 0 || e[1] >= mblev[m]) continue;
      int o[3], f1, f2, rx1, rx2, rx3;
      smr_slot_decode(n, &o[0], &o[1], &o[2], &f1, &f2);
      fc_ref_offset(n, &rx1, &rx2, &rx3);
      const int my_lev = mblev[m];
      for (int v = 0; v < 3; ++v) {
        sbox x = smr_fc_box(&I, o, f1, f2, SMR_PROL, v, 0);
        for (int k = x.ks; k <= x.ke; ++k)
          for (int j = x.js; j <= x.je; ++j)
            for (int i = x.is; i <= x.ie; ++i) {
              const int fi = (i - I.cis) * 2 + I.is, fj = multi_d ? ((j - I.cjs) * 2 + I.js) : j, fk = three_d ? ((k - I.cks) * 2 + I.ks) : k;
              if (v == 0) {
                const double q = CX1F(cb1, m, k, j, i);
                double dv2 = 0.0, dv3 = 0.0;
                if (multi_d) dv2 = fc_mm(q - CX1F(cb1, m, k, j - 1, i), CX1F(cb1, m, k, j + 1, i) - q);
                if (three_d) dv3 = fc_mm(q - CX1F(cb1, m, k - 1, j, i), CX1F(cb1, m, k + 1, j, i) - q);
                FC_STORE(0, X1F(b1, m, fk, fj, fi), fk, fj, fi, q - dv2 - dv3);
                if (multi_d) FC_STORE(0, X1F(b1, m, fk, fj + 1, fi), fk, fj + 1, fi, q + dv2 - dv3);
                if (three_d) {
                  FC_STORE(0, X1F(b1, m, fk + 1, fj, fi), fk + 1, fj, fi, q - dv2 + dv3);
                  FC_STORE(0, X1F(b1, m, fk + 1, fj + 1, fi), fk + 1, fj + 1, fi, q + dv2 + dv3);
                }
              } else if (v == 1) {
                const double q = CX2F(cb2, m, k, j, i);
                const double dv1 = fc_mm(q - CX2F(cb2, m, k, j, i - 1), CX2F(cb2, m, k, j, i + 1) - q);
                double dv3 = 0.0;
                if (three_d) dv3 = fc_mm(q - CX2F(cb2, m, k - 1, j, i), CX2F(cb2, m, k + 1, j, i) - q);
                FC_STORE(1, X2F(b2, m, fk, fj, fi), fk, fj, fi, q - dv1 - dv3);
                FC_STORE(1, X2F(b2, m, fk, fj, fi + 1), fk, fj, fi + 1, q + dv1 - dv3);
                if (three_d) {
                  FC_STORE(1, X2F(b2, m, fk + 1, fj, fi), fk + 1, fj, fi, q - dv1 + dv3);
                  FC_STORE(1, X2F(b2, m, fk + 1, fj, fi + 1), fk + 1, fj, fi + 1, q + dv1 + dv3);
                }
              } else {
                const double q = CX3F(cb3, m, k, j, i);
                const double dv1 = fc_mm(q - CX3F(cb3, m, k, j, i - 1), CX3F(cb3, m, k, j, i + 1) - q);
                double dv2 = 0.0;
                if (multi_d) dv2 = fc_mm(q - CX3F(cb3, m, k, j - 1, i), CX3F(cb3, m, k, j + 1, i) - q);
                FC_STORE(2, X3F(b3, m, fk, fj, fi), fk, fj, fi, q - dv1 - dv2);
                FC_STORE(2, X3F(b3, m, fk, fj, fi + 1), fk, fj, fi + 1, q + dv1 - dv2);
                if (multi_d) {
                  FC_STORE(2, X3F(b3, m, fk, fj + 1, fi), fk, fj + 1, fi, q - dv1 + dv2);
                  FC_STORE(2, X3F(b3, m, fk, fj + 1, fi + 1), fk, fj + 1, fi + 1, q + dv1 + dv2);
                }
              }
            }
      }
    }
  /* interior faces: Toth & Roe divergence-preserving interpolation (prolongation.cpp:253-357) */
  for (int m = 0; m < c->nmb; ++m)
    for (int n = 0; n < 56; ++n) {
      const int *e = nghbr + ((size_t)m * 56 + n) * 4;
      if (e[0] < 0 || e[1] >= mblev[m]) continue;
      int o[3], f1, f2, rx1, rx2, rx3;
      smr_slot_decode(n, &o[0], &o[1], &o[2], &f1, &f2);
      fc_ref_offset(n, &rx1, &rx2, &rx3);
      const int my_lev = mblev[m];
      sbox p0 = smr_fc_box(&I, o, f1, f2, SMR_PROL, 0, 0), p1 = smr_fc_box(&I, o, f1, f2, SMR_PROL, 1, 0), p2 = smr_fc_box(&I, o, f1, f2, SMR_PROL, 2, 0);
      for (int k = p1.ks; k <= p1.ke; ++k)
        for (int j = p0.js; j <= p0.je; ++j)
          for (int i = p2.is; i <= p2.ie; ++i) {
            const int fi = (i - I.cis) * 2 + I.is, fj = (j - I.cjs) * 2 + I.js, fk = (k - I.cks) * 2 + I.ks;
            if (!multi_d) {
              FC_STORE(0, X1F(b1, m, fk, fj, fi + 1), fk, fj, fi + 1, 0.5 * (X1F(b1, m, fk, fj, fi) + X1F(b1, m, fk, fj, fi + 2)));
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
              FC_STORE(0, X1F(b1, m, fk, fj, fi + 1), fk, fj, fi + 1, 0.5 * (X1F(b1, m, fk, fj, fi) + X1F(b1, m, fk, fj, fi + 2)) + Uxx - Vxyz - Wxyz);
              FC_STORE(0, X1F(b1, m, fk, fj + 1, fi + 1), fk, fj + 1, fi + 1, 0.5 * (X1F(b1, m, fk, fj + 1, fi) + X1F(b1, m, fk, fj + 1, fi + 2)) + Uxx - Vxyz + Wxyz);
              FC_STORE(0, X1F(b1, m, fk + 1, fj, fi + 1), fk + 1, fj, fi + 1, 0.5 * (X1F(b1, m, fk + 1, fj, fi) + X1F(b1, m, fk + 1, fj, fi + 2)) + Uxx + Vxyz - Wxyz);
              FC_STORE(0, X1F(b1, m, fk + 1, fj + 1, fi + 1), fk + 1, fj + 1, fi + 1, 0.5 * (X1F(b1, m, fk + 1, fj + 1, fi) + X1F(b1, m, fk + 1, fj + 1, fi + 2)) + Uxx + Vxyz + Wxyz);
              FC_STORE(1, X2F(b2, m, fk, fj + 1, fi), fk, fj + 1, fi, 0.5 * (X2F(b2, m, fk, fj, fi) + X2F(b2, m, fk, fj + 2, fi)) + Vyy - Uxyz - Wxyz);
              FC_STORE(1, X2F(b2, m, fk, fj + 1, fi + 1), fk, fj + 1, fi + 1, 0.5 * (X2F(b2, m, fk, fj, fi + 1) + X2F(b2, m, fk, fj + 2, fi + 1)) + Vyy - Uxyz + Wxyz);
              FC_STORE(1, X2F(b2, m, fk + 1, fj + 1, fi), fk + 1, fj + 1, fi, 0.5 * (X2F(b2, m, fk + 1, fj, fi) + X2F(b2, m, fk + 1, fj + 2, fi)) + Vyy + Uxyz - Wxyz);
              FC_STORE(1, X2F(b2, m, fk + 1, fj + 1, fi + 1), fk + 1, fj + 1, fi + 1, 0.5 * (X2F(b2, m, fk + 1, fj, fi + 1) + X2F(b2, m, fk + 1, fj + 2, fi + 1)) + Vyy + Uxyz + Wxyz);
              FC_STORE(2, X3F(b3, m, fk + 1, fj, fi), fk + 1, fj, fi, 0.5 * (X3F(b3, m, fk + 2, fj, fi) + X3F(b3, m, fk, fj, fi)) + Wzz - Uxyz - Vxyz);
              FC_STORE(2, X3F(b3, m, fk + 1, fj, fi + 1), fk + 1, fj, fi + 1, 0.5 * (X3F(b3, m, fk + 2, fj, fi + 1) + X3F(b3, m, fk, fj, fi + 1)) + Wzz - Uxyz + Vxyz);
              FC_STORE(2, X3F(b3, m, fk + 1, fj + 1, fi), fk + 1, fj + 1, fi, 0.5 * (X3F(b3, m, fk + 2, fj + 1, fi) + X3F(b3, m, fk, fj + 1, fi)) + Wzz + Uxyz - Vxyz);
              FC_STORE(2, X3F(b3, m, fk + 1, fj + 1, fi + 1), fk + 1, fj + 1, fi + 1, 0.5 * (X3F(b3, m, fk + 2, fj + 1, fi + 1) + X3F(b3, m, fk, fj + 1, fi + 1)) + Wzz + Uxyz + Vxyz);
            } else {
              double tmp1 = 0.25 * (X2F(b2, m, fk, fj + 2, fi + 1) - X2F(b2, m, fk, fj, fi + 1) - X2F(b2, m, fk, fj + 2, fi) + X2F(b2, m, fk, fj, fi));
              double tmp2 = 0.25 * (X1F(b1, m, fk, fj, fi) - X1F(b1, m, fk, fj, fi + 2) - X1F(b1, m, fk, fj + 1, fi) + X1F(b1, m, fk, fj + 1, fi + 2));
              FC_STORE(0, X1F(b1, m, fk, fj, fi + 1), fk, fj, fi + 1, 0.5 * (X1F(b1, m, fk, fj, fi) + X1F(b1, m, fk, fj, fi + 2)) + tmp1);
              FC_STORE(0, X1F(b1, m, fk, fj + 1, fi + 1), fk, fj + 1, fi + 1, 0.5 * (X1F(b1, m, fk, fj + 1, fi) + X1F(b1, m, fk, fj + 1, fi + 2)) + tmp1);
              FC_STORE(1, X2F(b2, m, fk, fj + 1, fi), fk, fj + 1, fi, 0.5 * (X2F(b2, m, fk, fj, fi) + X2F(b2, m, fk, fj + 2, fi)) + tmp2);
              FC_STORE(1, X2F(b2, m, fk, fj + 1, fi + 1), fk, fj + 1, fi + 1, 0.5 * (X2F(b2, m, fk, fj, fi + 1) + X2F(b2, m, fk, fj + 2, fi + 1)) + tmp2);
            }
          }
    }
#undef FC_STORE
}

/* MHD::Prolongate (mhd_tasks.cpp:527-551) + MHD::ApplyPhysicalBCs (:504-520): coarse fills of U and B, physical BCs on
 * the COARSE arrays (HydroBCsCoarse / BFieldBCsCoarse, bvals/physics/bfield_bcs.cpp:44-63: the same fill on the coarse
 * index space), prolongation of U and B, then physical BCs on the fine arrays */
static void mhd_smr_prolongate_and_bcs(const ocfg *c, ostate *s, double *cu0, double *cb1, double *cb2, double *cb3,
                                       const int *nghbr, const int *mblev) {
  odims D = dims(c);
  oracle_smr_prolongate_cc_steps(c, s->u0, cu0, D.nv, nghbr, mblev, 1);
  oracle_smr_prolongate_fc_steps(c, s->b0x1f, s->b0x2f, s->b0x3f, cb1, cb2, cb3, nghbr, mblev, 1);
  if (g_mesh_bcs) {
    ocfg cc = *c;
    sidx I = smr_indcs(c);
    cc.nx1 = I.cnx1; cc.nx2 = I.multi_d ? I.cnx2 : 1; cc.nx3 = I.three_d ? I.cnx3 : 1;
    oracle_apply_bcs(&cc, g_mesh_bcs, g_mb_bcs, cu0, cb1, cb2, cb3);
  }
  oracle_smr_prolongate_cc_steps(c, s->u0, cu0, D.nv, nghbr, mblev, 2);
  oracle_smr_prolongate_fc_steps(c, s->b0x1f, s->b0x2f, s->b0x3f, cb1, cb2, cb3, nghbr, mblev, 2);
  if (g_mesh_bcs) oracle_apply_bcs(c, g_mesh_bcs, g_mb_bcs, s->u0, s->b0x1f, s->b0x2f, s->b0x3f);
}

/* Driver::InitBoundaryValuesAndPrimitives for MHD on a multilevel mesh (driver.cpp:607-653):
 * RestrictU/B, SendU/B, RecvU/B, Prolongate (mhd_tasks.cpp:527-551), ConToPrim */
void oracle_mhd_smr_ghosts_and_prims(const ocfg *c, ostate *s, double *cu0, double *cb1, double *cb2, double *cb3,
                                     const int *nghbr, const int *mblev) {
  odims D = dims(c);
  oracle_smr_restrict_cc(c, s->u0, cu0, D.nv);
  oracle_smr_restrict_fc(c, s->b0x1f, s->b0x2f, s->b0x3f, cb1, cb2, cb3);
  oracle_smr_halo_cc(c, s->u0, cu0, D.nv, nghbr, mblev);
  oracle_smr_halo_fc(c, s->b0x1f, s->b0x2f, s->b0x3f, cb1, cb2, cb3, nghbr, mblev);
  mhd_smr_prolongate_and_bcs(c, s, cu0, cb1, cb2, cb3, nghbr, mblev);
  int nf[3];
  oracle_cons2prim(c, s->u0, s->b0x1f, s->b0x2f, s->b0x3f, s->w0, s->bcc0, 0, D.n1 - 1, 0, D.n2 - 1, 0, D.n3 - 1, nf);
}

/* ---------------------------------------------------------------------------------------------------------
 * SendE / RecvE on a multilevel mesh: flux_correct_fc.cpp:29-366 (pack: same level, or restricted onto the
 * coarser neighbour), :374-437 (RecvAndUnpackFluxFC), :445-644 (SumBoundaryFluxes), :655-796
 * (ZeroFluxesAtBoundaryWithFiner), :801-1034 (AverageBoundaryFluxes); index tables buffs_fc.cpp:282-385 (send)
 * and :785-935 (recv).  No shear-periodic boundaries. */
enum { EMF_SAME = 0, EMF_COAR_SEND = 1, EMF_COAR_RECV = 2 };
static sbox emf_flux_box(const sidx *I, const int o[3], int f1, int f2, int kind, int v) {
  const int s_[3] = {I->is, I->js, I->ks}, e_[3] = {I->ie, I->je, I->ke};
  const int cs_[3] = {I->cis, I->cjs, I->cks}, ce_[3] = {I->cie, I->cje, I->cke};
  const int cnx[3] = {I->cnx1, I->cnx2, I->cnx3};
  const int act[3] = {1, I->nx2 > 1, I->nx3 > 1};
  int lo[3], hi[3];
  for (int d = 0; d < 3; ++d) {
    const int node = (v != d);      /* an edge field is cell-like along its own direction, node-like across */
    const int s = (kind == EMF_COAR_SEND) ? cs_[d] : s_[d], e = (kind == EMF_COAR_SEND) ? ce_[d] : e_[d];
    if (o[d] == 0) {
      lo[d] = s; hi[d] = e + node;
      if (kind == EMF_COAR_RECV && (d == 0 || act[d])) {
        const int sub = (d == 0) ? f1 : (d == 1) ? ((o[0] != 0) ? f1 : f2) : ((o[0] != 0 && o[1] != 0) ? f1 : f2);
        if (sub == 1) lo[d] += cnx[d]; else hi[d] -= cnx[d];
      }
    } else if (o[d] > 0) { lo[d] = hi[d] = e + 1; }
    else { lo[d] = hi[d] = s; }
  }
  sbox b = {lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]};
  return b;
}
/* which edge-field component(s) a slot carries: faces the two tangential ones, edges the one along the edge */
static int emf_carries(int n, int v) {
  if (n < 8) return v != 0;
  if (n < 16) return v != 1;
  if (n < 24) return v == 2;
  if (n < 32) return v != 2;
  if (n < 40) return v == 1;
  if (n < 48) return v == 0;
  return 0;
}
static void emf_nflx_faces(int *nf, int n, int add) {   /* the four edge counters of a face slot with sub-block (0,0) */
  static const int tab[6][5] = {{0, 16, 20, 32, 36}, {4, 18, 22, 34, 38}, {8, 16, 18, 40, 44}, {12, 20, 22, 42, 46}, {24, 32, 34, 40, 42}, {28, 36, 38, 44, 46}};
  for (int q = 0; q < 6; ++q)
    if (tab[q][0] == n) for (int r = 1; r <= 4; ++r) { if (add) nf[tab[q][r]] += 1; else nf[tab[q][r]] = 0; }
}

void oracle_smr_emf_correct(const ocfg *c, double *e1, double *e2, double *e3, const int *nghbr, const int *mblev) {
  odims D = dims(c);
  sidx I = smr_indcs(c);
  const int nmb = c->nmb, multi_d = I.multi_d, three_d = I.three_d;
  const int actv[3] = {1, multi_d, three_d};
  double **rbuf = calloc((size_t)nmb * 48 * 3, sizeof(double *));
  /* ---- pack */
  for (int m = 0; m < nmb; ++m)
    for (int n = 0; n < 48; ++n) {
      const int *e = nghbr + ((size_t)m * 56 + n) * 4;
      if (e[0] < 0 || e[1] > mblev[m]) continue;
      int o[3], f1, f2;
      smr_slot_decode(n, &o[0], &o[1], &o[2], &f1, &f2);
      const int coarser = e[1] < mblev[m];
      for (int v = 0; v < 3; ++v) {
        if (!emf_carries(n, v)) continue;
        sbox x = emf_flux_box(&I, o, f1, f2, coarser ? EMF_COAR_SEND : EMF_SAME, v);
        const int ni = x.ie - x.is + 1, nj = x.je - x.js + 1, nk = x.ke - x.ks + 1;
        double *buf = malloc(sizeof(double) * (size_t)ni * nj * nk);
        for (int k = x.ks; k <= x.ke; ++k)
          for (int j = x.js; j <= x.je; ++j)
            for (int i = x.is; i <= x.ie; ++i) {
              double r;
              if (!coarser) r = *emf_at(&D, e1, e2, e3, v, m, k, j, i);
              else {
                /* flux_correct_fc.cpp:108-113,131-136: inactive directions always read index 0 of the fine array */
                const int fi = 2 * i - I.cis, fj = multi_d ? 2 * j - I.cjs : 0, fk = three_d ? 2 * k - I.cks : 0;
                const int d1 = (v == 0), d2 = (v == 1), d3 = (v == 2);
                if (actv[v]) r = 0.5 * (*emf_at(&D, e1, e2, e3, v, m, fk, fj, fi) + *emf_at(&D, e1, e2, e3, v, m, fk + d3, fj + d2, fi + d1));
                else r = *emf_at(&D, e1, e2, e3, v, m, fk, fj, fi);
              }
              buf[(i - x.is) + ni * ((j - x.js) + nj * (k - x.ks))] = r;
            }
        rbuf[((size_t)e[0] * 48 + e[3]) * 3 + v] = buf;
      }
    }
  int *nflx = malloc(sizeof(int) * (size_t)nmb * 48);
  for (int q = 0; q < nmb * 48; ++q) nflx[q] = 1;
  const int multilevel = 1;
  for (int pass = 0; pass < (multilevel ? 2 : 1); ++pass) {
    const int same_level = (pass == 0);
    if (!same_level) {   /* ZeroFluxesAtBoundaryWithFiner */
      for (int m = 0; m < nmb; ++m)
        for (int n = 0; n < 48; ++n) {
          const int *e = nghbr + ((size_t)m * 56 + n) * 4;
          if (e[0] < 0 || e[1] <= mblev[m]) continue;
          int o[3], f1, f2;
          smr_slot_decode(n, &o[0], &o[1], &o[2], &f1, &f2);
          if (n < 16 || (n >= 24 && n < 32)) emf_nflx_faces(nflx + (size_t)m * 48, n, 0);
          else nflx[(size_t)m * 48 + n] = 0;
          for (int v = 0; v < 3; ++v) {
            if (!emf_carries(n, v)) continue;
            sbox x = emf_flux_box(&I, o, f1, f2, EMF_COAR_RECV, v);
            for (int k = x.ks; k <= x.ke; ++k) for (int j = x.js; j <= x.je; ++j) for (int i = x.is; i <= x.ie; ++i)
              *emf_at(&D, e1, e2, e3, v, m, k, j, i) = 0.0;
          }
        }
    }
    /* SumBoundaryFluxes: per (block, component) in increasing slot order */
    for (int m = 0; m < nmb; ++m)
      for (int n = 0; n < 48; ++n) {
        const int *e = nghbr + ((size_t)m * 56 + n) * 4;
        if (e[0] < 0) continue;
        if (same_level ? (e[1] != mblev[m]) : (e[1] <= mblev[m])) continue;
        int o[3], f1, f2;
        smr_slot_decode(n, &o[0], &o[1], &o[2], &f1, &f2);
        if (n < 16 || (n >= 24 && n < 32)) emf_nflx_faces(nflx + (size_t)m * 48, n, 1);     /* only n == 0,4,8,12,24,28 are in the table */
        else nflx[(size_t)m * 48 + n] += 1;
        for (int v = 0; v < 3; ++v) {
          if (!emf_carries(n, v)) continue;
          sbox x = emf_flux_box(&I, o, f1, f2, same_level ? EMF_SAME : EMF_COAR_RECV, v);
          const int ni = x.ie - x.is + 1, nj = x.je - x.js + 1;
          const double *buf = rbuf[((size_t)m * 48 + n) * 3 + v];
          for (int k = x.ks; k <= x.ke; ++k) for (int j = x.js; j <= x.je; ++j) for (int i = x.is; i <= x.ie; ++i)
            *emf_at(&D, e1, e2, e3, v, m, k, j, i) += buf[(i - x.is) + ni * ((j - x.js) + nj * (k - x.ks))];
        }
      }
  }
  /* AverageBoundaryFluxes: slots with sub-block (0,0) only; boxes = the same-level receive boxes */
  for (int m = 0; m < nmb; ++m)
    for (int n = 0; n < 48; ++n) {
      int o[3], f1, f2;
      smr_slot_decode(n, &o[0], &o[1], &o[2], &f1, &f2);
      if (f1 || f2) continue;
      if ((o[1] && !multi_d) || (o[2] && !three_d)) continue;
      const int *e = nghbr + ((size_t)m * 56 + n) * 4;
      const int nz = (o[0] != 0) + (o[1] != 0) + (o[2] != 0);
      for (int v = 0; v < 3; ++v) {
        if (!emf_carries(n, v)) continue;
        sbox x = emf_flux_box(&I, o, 0, 0, EMF_SAME, v);
        int lo[3] = {x.is, x.js, x.ks}, hi[3] = {x.ie, x.je, x.ke};
        if (nz == 2) {     /* block edge: divide by the number of contributions */
          for (int k = lo[2]; k <= hi[2]; ++k) for (int j = lo[1]; j <= hi[1]; ++j) for (int i = lo[0]; i <= hi[0]; ++i)
            *emf_at(&D, e1, e2, e3, v, m, k, j, i) /= (double)nflx[(size_t)m * 48 + n];
          continue;
        }
        /* face: t = the tangential direction ACROSS the edge field (the node-like one) */
        int t = -1;
        for (int d = 0; d < 3; ++d) if (o[d] == 0 && d != v) t = d;
        if (e[1] == mblev[m]) {              /* same level: halve the face interior, excluding the edge lines */
          if (actv[t] || t == 0) { if (t == 0 || (t == 1 ? multi_d : three_d)) { lo[t] += 1; hi[t] -= 1; } }
          for (int k = lo[2]; k <= hi[2]; ++k) for (int j = lo[1]; j <= hi[1]; ++j) for (int i = lo[0]; i <= hi[0]; ++i)
            *emf_at(&D, e1, e2, e3, v, m, k, j, i) *= 0.5;
        } else if (e[1] >= mblev[m]) {       /* finer: halve the line where two fine faces meet */
          if (!(t == 0 || (t == 1 ? multi_d : three_d))) continue;
          const int mid = lo[t] + (hi[t] - lo[t] + 1) / 2;
          lo[t] = hi[t] = mid;
          for (int k = lo[2]; k <= hi[2]; ++k) for (int j = lo[1]; j <= hi[1]; ++j) for (int i = lo[0]; i <= hi[0]; ++i)
            *emf_at(&D, e1, e2, e3, v, m, k, j, i) *= 0.5;
        }
      }
    }
  for (size_t q = 0; q < (size_t)nmb * 48 * 3; ++q) free(rbuf[q]);
  free(rbuf);
  free(nflx);
}

/* MHD stage on a multilevel mesh: mhd_tasks.cpp:48-75 (physical BCs: oracle_set_bcs) */
int oracle_mhd_smr_run(const ocfg *c, ostate *s, double *cu0, double *cb1, double *cb2, double *cb3, const double *mb_size,
                       const int *nghbr, const int *mblev, int integrator, double cfl, double tlim, int ncycles,
                       double *time, double *dt) {
  odims D = dims(c);
  size_t nm = (size_t)c->nmb;
  double g0[3] = {0.0, 0.0, 0.0}, g1[3] = {1.0, 0.0, 0.0}, be[3] = {1.0, 0.0, 0.0};
  if (integrator == 2) { g0[1] = 0.5; g1[1] = 0.5; be[1] = 0.5; }
  if (integrator == 3) { g0[1] = 0.25; g1[1] = 0.75; be[1] = 0.25; g0[2] = 2.0 / 3.0; g1[2] = 1.0 / 3.0; be[2] = 2.0 / 3.0; }
  double *flx1 = calloc(D.nv * nm * D.f1, 8), *flx2 = calloc(D.nv * nm * D.f2, 8), *flx3 = calloc(D.nv * nm * D.f3, 8);
  double *fe[6];
  for (int q = 0; q < 6; ++q) fe[q] = calloc(nm * D.cc, 8);
  double *e1 = calloc(nm * D.e1, 8), *e2 = calloc(nm * D.e2, 8), *e3 = calloc(nm * D.e3, 8);
  int n = 0;
  while (*time < tlim && (ncycles < 0 || n < ncycles)) {
    for (int st = 1; st <= integrator; ++st) {
      const double bdt = be[st - 1] * (*dt);
      if (st == 1) {
        memcpy(s->u1, s->u0, 8 * D.nv * nm * D.cc);
        memcpy(s->b1x1f, s->b0x1f, 8 * nm * D.f1); memcpy(s->b1x2f, s->b0x2f, 8 * nm * D.f2); memcpy(s->b1x3f, s->b0x3f, 8 * nm * D.f3);
      }
      oracle_fluxes(c, s->w0, s->bcc0, s->b0x1f, s->b0x2f, s->b0x3f, flx1, flx2, flx3, fe[0], fe[1], fe[2], fe[3], fe[4], fe[5]);
      if (c->fofc)   /* mhd_tasks.cpp:208-211: first-order flux correction inside the Fluxes task, before SendFlux */
        oracle_fofc(c, s->u0, s->u1, s->w0, s->bcc0, s->b0x1f, s->b0x2f, s->b0x3f, s->b1x1f, s->b1x2f, s->b1x3f, flx1, flx2, flx3,
                    fe[0], fe[1], fe[2], fe[3], fe[4], fe[5], mb_size, g0[st - 1], g1[st - 1], bdt);
      oracle_smr_flux_correct_cc(c, flx1, flx2, flx3, D.nv, nghbr, mblev);                      /* SendFlux / RecvFlux */
      oracle_rk_update(c, s->u0, s->u1, flx1, flx2, flx3, mb_size, g0[st - 1], g1[st - 1], bdt);
      oracle_smr_restrict_cc(c, s->u0, cu0, D.nv);                                                /* RestrictU, SendU, RecvU */
      oracle_smr_halo_cc(c, s->u0, cu0, D.nv, nghbr, mblev);
      oracle_corner_e(c, s->w0, s->bcc0, flx1, flx2, flx3, fe[0], fe[1], fe[2], fe[3], fe[4], fe[5], e1, e2, e3);
      oracle_smr_emf_correct(c, e1, e2, e3, nghbr, mblev);                                        /* SendE / RecvE */
      oracle_ct(c, s->b0x1f, s->b0x2f, s->b0x3f, s->b1x1f, s->b1x2f, s->b1x3f, e1, e2, e3, mb_size, g0[st - 1], g1[st - 1], bdt);
      oracle_smr_restrict_fc(c, s->b0x1f, s->b0x2f, s->b0x3f, cb1, cb2, cb3);                     /* RestrictB, SendB, RecvB */
      oracle_smr_halo_fc(c, s->b0x1f, s->b0x2f, s->b0x3f, cb1, cb2, cb3, nghbr, mblev);
      mhd_smr_prolongate_and_bcs(c, s, cu0, cb1, cb2, cb3, nghbr, mblev);                         /* Prolongate, ApplyPhysicalBCs */
      int nf[3];
      oracle_cons2prim(c, s->u0, s->b0x1f, s->b0x2f, s->b0x3f, s->w0, s->bcc0, 0, D.n1 - 1, 0, D.n2 - 1, 0, D.n3 - 1, nf);
    }
    double dtnew = oracle_new_dt(c, s->w0, s->bcc0, mb_size, NULL);
    *time = *time + *dt;
    n++;
    double d = 2.0 * (*dt);
    double lim = cfl * dtnew;
    if (lim < d) d = lim;
    if ((*time < tlim) && ((*time + d) > tlim)) d = tlim - *time;
    *dt = d;
  }
  free(flx1); free(flx2); free(flx3);
  for (int q = 0; q < 6; ++q) free(fe[q]);
  free(e1); free(e2); free(e3);
  return n;
}
