/* smr_oracle.c -- TEST INFRASTRUCTURE ONLY (#included by mhd_oracle.c).
 * CPU restatement of the reference's fine/coarse boundary path for CELL-CENTRED data on a statically
 * refined mesh, hydro: RestrictU -> SendU/RecvU across levels -> Prolongate, and the
 * flux correction at fine/coarse faces.  Single rank, all MeshBlocks; loops in the reference's
 * Kokkos::Serial order (block, neighbour slot, variable).
 *   index tables      src/bvals/buffs_cc.cpp:28-152 (send), :164-484 (recv, prolongation, flux)
 *   RestrictCC        src/mesh/mesh_refinement.cpp:1226-1284
 *   pack / unpack     src/bvals/bvals_cc.cpp:42-167, :273-402
 *   FillCoarseInBndry src/bvals/prolongation.cpp:366-464;  ProlongateCC :471-547;  ProlongCC mesh/prolongation.hpp:21-60
 *   flux correction   src/bvals/flux_correct_cc.cpp:28-167 (pack), :173-290 (unpack)
 *   task order        src/hydro/hydro_tasks.cpp:47-79, Driver::InitBoundaryValuesAndPrimitives driver.cpp:587-604
 * Parity PINNED: tests/test_smr_cpu.py compares whole runs with restart dumps of the unmodified reference. */

typedef struct { int is, ie, js, je, ks, ke; } sbox;

/* NeighborIndex inverse: slot -> (ox1,ox2,ox3,f1,f2) */
static int smr_slot_decode(int n, int *ox, int *oy, int *oz, int *f1, int *f2) {
  for (int iz = -1; iz <= 1; ++iz)
    for (int iy = -1; iy <= 1; ++iy)
      for (int ix = -1; ix <= 1; ++ix) {
        if (!ix && !iy && !iz) continue;
        /* faces carry (f1,f2), edges f1 only, corners neither: other combinations alias other slots */
        const int nzero = (ix == 0) + (iy == 0) + (iz == 0);
        for (int a = 0; a <= (nzero >= 1 ? 1 : 0); ++a)
          for (int b = 0; b <= (nzero == 2 ? 1 : 0); ++b)
            if (oracle_neighbor_index(ix, iy, iz, a, b) == n) {
              *ox = ix; *oy = iy; *oz = iz; *f1 = a; *f2 = b;
              return 1;
            }
      }
  return 0;
}

typedef struct {
  int ng, nx1, nx2, nx3, cnx1, cnx2, cnx3;
  int is, ie, js, je, ks, ke, cis, cie, cjs, cje, cks, cke;
  int multi_d, three_d;
} sidx;

static sidx smr_indcs(const ocfg *c) {
  sidx x;
  odims D = dims(c);
  x.ng = c->ng; x.nx1 = c->nx1; x.nx2 = c->nx2; x.nx3 = c->nx3;
  x.multi_d = c->nx2 > 1; x.three_d = c->nx3 > 1;
  x.cnx1 = c->nx1 / 2; x.cnx2 = (c->nx2 / 2 > 1) ? c->nx2 / 2 : 1; x.cnx3 = (c->nx3 / 2 > 1) ? c->nx3 / 2 : 1;   /* mesh.cpp:286-289 */
  x.is = D.is; x.ie = D.ie; x.js = D.js; x.je = D.je; x.ks = D.ks; x.ke = D.ke;
  x.cis = c->ng; x.cie = x.cis + x.cnx1 - 1;
  x.cjs = x.multi_d ? c->ng : 0; x.cje = x.cjs + (x.multi_d ? x.cnx2 : 1) - 1;
  x.cks = x.three_d ? c->ng : 0; x.cke = x.cks + (x.three_d ? x.cnx3 : 1) - 1;
  return x;
}

enum { SMR_SAME = 0, SMR_COAR = 1, SMR_FINE = 2, SMR_FLUX = 3, SMR_PROL = 4 };

/* buffs_cc.cpp:28-152 */
static sbox smr_send_box(const sidx *I, int ox1, int ox2, int ox3, int f1, int f2, int kind) {
  sbox b;
  const int ng = I->ng, ng1 = ng - 1;
  if (kind == SMR_SAME) {
    b.is = (ox1 > 0) ? (I->ie - ng1) : I->is; b.ie = (ox1 < 0) ? (I->is + ng1) : I->ie;
    b.js = (ox2 > 0) ? (I->je - ng1) : I->js; b.je = (ox2 < 0) ? (I->js + ng1) : I->je;
    b.ks = (ox3 > 0) ? (I->ke - ng1) : I->ks; b.ke = (ox3 < 0) ? (I->ks + ng1) : I->ke;
  } else if (kind == SMR_COAR) {
    b.is = (ox1 > 0) ? (I->cie - ng1) : I->cis; b.ie = (ox1 < 0) ? (I->cis + ng1) : I->cie;
    b.js = (ox2 > 0) ? (I->cje - ng1) : I->cjs; b.je = (ox2 < 0) ? (I->cjs + ng1) : I->cje;
    b.ks = (ox3 > 0) ? (I->cke - ng1) : I->cks; b.ke = (ox3 < 0) ? (I->cks + ng1) : I->cke;
  } else if (kind == SMR_FINE) {
    b.is = (ox1 > 0) ? (I->ie - ng1) : I->is; b.ie = (ox1 < 0) ? (I->is + ng1) : I->ie;
    b.js = (ox2 > 0) ? (I->je - ng1) : I->js; b.je = (ox2 < 0) ? (I->js + ng1) : I->je;
    b.ks = (ox3 > 0) ? (I->ke - ng1) : I->ks; b.ke = (ox3 < 0) ? (I->ks + ng1) : I->ke;
    if (ox1 == 0) { if (f1 == 1) b.is += I->cnx1 - ng; else b.ie -= I->cnx1 - ng; }
    if (ox2 == 0 && I->nx2 > 1) {
      if (ox1 != 0) { if (f1 == 1) b.js += I->cnx2 - ng; else b.je -= I->cnx2 - ng; }
      else { if (f2 == 1) b.js += I->cnx2 - ng; else b.je -= I->cnx2 - ng; }
    }
    if (ox3 == 0 && I->nx3 > 1) {
      if (ox1 != 0 && ox2 != 0) { if (f1 == 1) b.ks += I->cnx3 - ng; else b.ke -= I->cnx3 - ng; }
      else { if (f2 == 1) b.ks += I->cnx3 - ng; else b.ke -= I->cnx3 - ng; }
    }
  } else {   /* SMR_FLUX: coarse indices */
    if (ox1 == 0) { b.is = I->cis; b.ie = I->cie; } else if (ox1 > 0) { b.is = b.ie = I->cie + 1; } else { b.is = b.ie = I->cis; }
    if (ox2 == 0) { b.js = I->cjs; b.je = I->cje; } else if (ox2 > 0) { b.js = b.je = I->cje + 1; } else { b.js = b.je = I->cjs; }
    if (ox3 == 0) { b.ks = I->cks; b.ke = I->cke; } else if (ox3 > 0) { b.ks = b.ke = I->cke + 1; } else { b.ks = b.ke = I->cks; }
  }
  return b;
}

/* buffs_cc.cpp:164-484 */
static sbox smr_recv_box(const sidx *I, int ox1, int ox2, int ox3, int f1, int f2, int kind) {
  sbox b;
  const int ng = I->ng;
  if (kind == SMR_SAME) {
    if (ox1 == 0) { b.is = I->is; b.ie = I->ie; } else if (ox1 > 0) { b.is = I->ie + 1; b.ie = I->ie + ng; } else { b.is = I->is - ng; b.ie = I->is - 1; }
    if (ox2 == 0) { b.js = I->js; b.je = I->je; } else if (ox2 > 0) { b.js = I->je + 1; b.je = I->je + ng; } else { b.js = I->js - ng; b.je = I->js - 1; }
    if (ox3 == 0) { b.ks = I->ks; b.ke = I->ke; } else if (ox3 > 0) { b.ks = I->ke + 1; b.ke = I->ke + ng; } else { b.ks = I->ks - ng; b.ke = I->ks - 1; }
  } else if (kind == SMR_COAR || kind == SMR_PROL) {
    const int cn = (kind == SMR_COAR) ? ng : ng / 2;
    if (ox1 == 0) { b.is = I->cis; b.ie = I->cie; if (f1 == 0) b.ie += cn; else b.is -= cn; }
    else if (ox1 > 0) { b.is = I->cie + 1; b.ie = I->cie + cn; } else { b.is = I->cis - cn; b.ie = I->cis - 1; }
    if (ox2 == 0) {
      b.js = I->cjs; b.je = I->cje;
      if (I->nx2 > 1) {
        if (ox1 != 0) { if (f1 == 0) b.je += cn; else b.js -= cn; }
        else { if (f2 == 0) b.je += cn; else b.js -= cn; }
      }
    } else if (ox2 > 0) { b.js = I->cje + 1; b.je = I->cje + cn; } else { b.js = I->cjs - cn; b.je = I->cjs - 1; }
    if (ox3 == 0) {
      b.ks = I->cks; b.ke = I->cke;
      if (I->nx3 > 1) {
        if (ox1 != 0 && ox2 != 0) { if (f1 == 0) b.ke += cn; else b.ks -= cn; }
        else { if (f2 == 0) b.ke += cn; else b.ks -= cn; }
      }
    } else if (ox3 > 0) { b.ks = I->cke + 1; b.ke = I->cke + cn; } else { b.ks = I->cks - cn; b.ke = I->cks - 1; }
  } else {   /* SMR_FINE and SMR_FLUX share the sub-block logic; FLUX collapses the normal direction to one face */
    const int w = (kind == SMR_FINE) ? ng : 1;      /* ghost width / single face */
    if (ox1 == 0) { b.is = I->is; b.ie = I->ie; if (f1 == 1) b.is += I->cnx1; else b.ie -= I->cnx1; }
    else if (ox1 > 0) { b.is = I->ie + 1; b.ie = I->ie + w; }
    else { if (kind == SMR_FINE) { b.is = I->is - ng; b.ie = I->is - 1; } else { b.is = b.ie = I->is; } }
    if (ox2 == 0) {
      b.js = I->js; b.je = I->je;
      if (I->nx2 > 1) {
        if (ox1 != 0) { if (f1 == 1) b.js += I->cnx2; else b.je -= I->cnx2; }
        else { if (f2 == 1) b.js += I->cnx2; else b.je -= I->cnx2; }
      }
    } else if (ox2 > 0) { b.js = I->je + 1; b.je = I->je + w; }
    else { if (kind == SMR_FINE) { b.js = I->js - ng; b.je = I->js - 1; } else { b.js = b.je = I->js; } }
    if (ox3 == 0) {
      b.ks = I->ks; b.ke = I->ke;
      if (I->nx3 > 1) {
        if (ox1 != 0 && ox2 != 0) { if (f1 == 1) b.ks += I->cnx3; else b.ke -= I->cnx3; }
        else { if (f2 == 1) b.ks += I->cnx3; else b.ke -= I->cnx3; }
      }
    } else if (ox3 > 0) { b.ks = I->ke + 1; b.ke = I->ke + w; }
    else { if (kind == SMR_FINE) { b.ks = I->ks - ng; b.ke = I->ks - 1; } else { b.ks = b.ke = I->ks; } }
  }
  return b;
}

/* coarse arrays have (cnx + 2ng) cells per active dimension (hydro.cpp:121-128) */
typedef struct { int n1, n2, n3; size_t cc; } cdims;
static cdims smr_cdims(const sidx *I) {
  cdims d;
  d.n1 = I->cnx1 + 2 * I->ng; d.n2 = I->multi_d ? I->cnx2 + 2 * I->ng : 1; d.n3 = I->three_d ? I->cnx3 + 2 * I->ng : 1;
  d.cc = (size_t)d.n3 * d.n2 * d.n1;
  return d;
}
#define CCC(a, nv, m, n, k, j, i) (a)[(((size_t)(m) * (nv) + (n)) * CD.n3 + (k)) * CD.n2 * CD.n1 + (size_t)(j) * CD.n1 + (i)]

size_t oracle_smr_coarse_cells(const ocfg *c) { sidx I = smr_indcs(c); return smr_cdims(&I).cc; }

/* mesh_refinement.cpp:1226-1284 */
void oracle_smr_restrict_cc(const ocfg *c, const double *u, double *cu, int nv) {
  odims D = dims(c);
  sidx I = smr_indcs(c);
  cdims CD = smr_cdims(&I);
  for (int m = 0; m < c->nmb; ++m)
    for (int n = 0; n < nv; ++n) {
      if (!I.multi_d) {
        for (int i = I.cis; i <= I.cie; ++i) {
          int fi = 2 * i - I.cis;
          CCC(cu, nv, m, n, I.cks, I.cjs, i) = 0.5 * (CC(u, nv, m, n, I.cks, I.cjs, fi) + CC(u, nv, m, n, I.cks, I.cjs, fi + 1));
        }
      } else if (!I.three_d) {
        for (int j = I.cjs; j <= I.cje; ++j)
          for (int i = I.cis; i <= I.cie; ++i) {
            int fi = 2 * i - I.cis, fj = 2 * j - I.cjs;
            CCC(cu, nv, m, n, I.cks, j, i) = 0.25 * (CC(u, nv, m, n, I.cks, fj, fi) + CC(u, nv, m, n, I.cks, fj, fi + 1)
                                                    + CC(u, nv, m, n, I.cks, fj + 1, fi) + CC(u, nv, m, n, I.cks, fj + 1, fi + 1));
          }
      } else {
        for (int k = I.cks; k <= I.cke; ++k)
          for (int j = I.cjs; j <= I.cje; ++j)
            for (int i = I.cis; i <= I.cie; ++i) {
              int fi = 2 * i - I.cis, fj = 2 * j - I.cjs, fk = 2 * k - I.cks;
              CCC(cu, nv, m, n, k, j, i) =
                  0.125 * (CC(u, nv, m, n, fk, fj, fi) + CC(u, nv, m, n, fk, fj, fi + 1)
                         + CC(u, nv, m, n, fk, fj + 1, fi) + CC(u, nv, m, n, fk, fj + 1, fi + 1)
                         + CC(u, nv, m, n, fk + 1, fj, fi) + CC(u, nv, m, n, fk + 1, fj, fi + 1)
                         + CC(u, nv, m, n, fk + 1, fj + 1, fi) + CC(u, nv, m, n, fk + 1, fj + 1, fi + 1));
            }
      }
    }
}

/* bvals_cc.cpp:42-167 + :273-402.  nghbr: [nmb][56][4] = gid, lev, rank, dest; mblev: [nmb].
 * Every message is packed first (senders read active cells only), then unpacked in (block, slot) order. */
void oracle_smr_halo_cc(const ocfg *c, double *u, double *cu, int nv, const int *nghbr, const int *mblev) {
  odims D = dims(c);
  sidx I = smr_indcs(c);
  cdims CD = smr_cdims(&I);
  const int nmb = c->nmb;
  double **rbuf = calloc((size_t)nmb * 56, sizeof(double *));
  for (int m = 0; m < nmb; ++m)
    for (int n = 0; n < 56; ++n) {
      const int *e = nghbr + ((size_t)m * 56 + n) * 4;
      if (e[0] < 0) continue;
      int ox, oy, oz, f1, f2;
      smr_slot_decode(n, &ox, &oy, &oz, &f1, &f2);
      const int kind = (e[1] < mblev[m]) ? SMR_COAR : (e[1] == mblev[m]) ? SMR_SAME : SMR_FINE;
      sbox b = smr_send_box(&I, ox, oy, oz, f1, f2, kind);
      const int ni = b.ie - b.is + 1, nj = b.je - b.js + 1, nk = b.ke - b.ks + 1;
      double *buf = malloc(sizeof(double) * (size_t)ni * nj * nk * nv);
      for (int v = 0; v < nv; ++v)
        for (int k = b.ks; k <= b.ke; ++k)
          for (int j = b.js; j <= b.je; ++j)
            for (int i = b.is; i <= b.ie; ++i)
              buf[(i - b.is) + ni * ((j - b.js) + nj * ((k - b.ks) + nk * v))] =
                  (kind == SMR_COAR) ? CCC(cu, nv, m, v, k, j, i) : CC(u, nv, m, v, k, j, i);
      rbuf[(size_t)e[0] * 56 + e[3]] = buf;
    }
  for (int m = 0; m < nmb; ++m)
    for (int n = 0; n < 56; ++n) {
      const int *e = nghbr + ((size_t)m * 56 + n) * 4;
      if (e[0] < 0) continue;
      int ox, oy, oz, f1, f2;
      smr_slot_decode(n, &ox, &oy, &oz, &f1, &f2);
      const int kind = (e[1] < mblev[m]) ? SMR_COAR : (e[1] == mblev[m]) ? SMR_SAME : SMR_FINE;
      sbox b = smr_recv_box(&I, ox, oy, oz, f1, f2, kind);
      const int ni = b.ie - b.is + 1, nj = b.je - b.js + 1, nk = b.ke - b.ks + 1;
      const double *buf = rbuf[(size_t)m * 56 + n];
      for (int v = 0; v < nv; ++v)
        for (int k = b.ks; k <= b.ke; ++k)
          for (int j = b.js; j <= b.je; ++j)
            for (int i = b.is; i <= b.ie; ++i) {
              const double val = buf[(i - b.is) + ni * ((j - b.js) + nj * ((k - b.ks) + nk * v))];
              if (kind == SMR_COAR) CCC(cu, nv, m, v, k, j, i) = val; else CC(u, nv, m, v, k, j, i) = val;
            }
    }
  for (size_t q = 0; q < (size_t)nmb * 56; ++q) free(rbuf[q]);
  free(rbuf);
}

static double smr_sign(double x) { return (x < 0.0) ? -1.0 : 1.0; }   /* athena.hpp SIGN(): ((x) < 0.0 ? -1.0 : 1.0) */

/* prolongation.cpp:366-547.  steps: 1 = FillCoarseInBndryCC, 2 = ProlongateCC, 3 = both (physical BCs on the
 * coarse array go between the two, hydro_tasks.cpp:380-394) */
void oracle_smr_prolongate_cc_steps(const ocfg *c, double *u, double *cu, int nv, const int *nghbr, const int *mblev, int steps) {
  odims D = dims(c);
  sidx I = smr_indcs(c);
  cdims CD = smr_cdims(&I);
  if (I.multi_d && (steps & 1))    /* FillCoarseInBndryCC: restrict the same-level ghost data into the coarse array */
    for (int m = 0; m < c->nmb; ++m)
      for (int n = 0; n < 56; ++n) {
        const int *e = nghbr + ((size_t)m * 56 + n) * 4;
        if (e[0] < 0 || e[1] != mblev[m]) continue;
        int ox, oy, oz, f1, f2;
        smr_slot_decode(n, &ox, &oy, &oz, &f1, &f2);
        sbox r = smr_recv_box(&I, ox, oy, oz, 0, 0, SMR_SAME);
        const int il = (r.is + I.cis) / 2, iu = (r.ie + I.cis) / 2, jl = (r.js + I.cjs) / 2, ju = (r.je + I.cjs) / 2;
        const int kl = (r.ks + I.cks) / 2, ku = (r.ke + I.cks) / 2;
        for (int v = 0; v < nv; ++v)
          for (int k = kl; k <= ku; ++k)
            for (int j = jl; j <= ju; ++j)
              for (int i = il; i <= iu; ++i) {
                const int fi = (i - I.cis) * 2 + I.is, fj = (j - I.cjs) * 2 + I.js, fk = (k - I.cks) * 2 + I.ks;
                if (!I.three_d)
                  CCC(cu, nv, m, v, kl, j, i) = 0.25 * (CC(u, nv, m, v, kl, fj, fi) + CC(u, nv, m, v, kl, fj, fi + 1)
                                                       + CC(u, nv, m, v, kl, fj + 1, fi) + CC(u, nv, m, v, kl, fj + 1, fi + 1));
                else
                  CCC(cu, nv, m, v, k, j, i) =
                      0.125 * (CC(u, nv, m, v, fk, fj, fi) + CC(u, nv, m, v, fk, fj, fi + 1)
                             + CC(u, nv, m, v, fk, fj + 1, fi) + CC(u, nv, m, v, fk, fj + 1, fi + 1)
                             + CC(u, nv, m, v, fk + 1, fj, fi) + CC(u, nv, m, v, fk + 1, fj, fi + 1)
                             + CC(u, nv, m, v, fk + 1, fj + 1, fi) + CC(u, nv, m, v, fk + 1, fj + 1, fi + 1));
              }
      }
  if (steps & 2)
  for (int m = 0; m < c->nmb; ++m)   /* ProlongateCC: slots whose neighbour is coarser */
    for (int n = 0; n < 56; ++n) {
      const int *e = nghbr + ((size_t)m * 56 + n) * 4;
      if (e[0] < 0 || e[1] >= mblev[m]) continue;
      int ox, oy, oz, f1, f2;
      smr_slot_decode(n, &ox, &oy, &oz, &f1, &f2);
      sbox b = smr_recv_box(&I, ox, oy, oz, f1, f2, SMR_PROL);
      for (int v = 0; v < nv; ++v)
        for (int k = b.ks; k <= b.ke; ++k)
          for (int j = b.js; j <= b.je; ++j)
            for (int i = b.is; i <= b.ie; ++i) {
              const int fi = (i - I.cis) * 2 + I.is, fj = (j - I.cjs) * 2 + I.js, fk = (k - I.cks) * 2 + I.ks;
              const double q = CCC(cu, nv, m, v, k, j, i);
              double dl = q - CCC(cu, nv, m, v, k, j, i - 1), dr = CCC(cu, nv, m, v, k, j, i + 1) - q;
              const double dvar1 = 0.125 * (smr_sign(dl) + smr_sign(dr)) * fmin(fabs(dl), fabs(dr));
              double dvar2 = 0.0, dvar3 = 0.0;
              if (I.multi_d) {
                dl = q - CCC(cu, nv, m, v, k, j - 1, i); dr = CCC(cu, nv, m, v, k, j + 1, i) - q;
                dvar2 = 0.125 * (smr_sign(dl) + smr_sign(dr)) * fmin(fabs(dl), fabs(dr));
              }
              if (I.three_d) {
                dl = q - CCC(cu, nv, m, v, k - 1, j, i); dr = CCC(cu, nv, m, v, k + 1, j, i) - q;
                dvar3 = 0.125 * (smr_sign(dl) + smr_sign(dr)) * fmin(fabs(dl), fabs(dr));
              }
              CC(u, nv, m, v, fk, fj, fi) = q - dvar1 - dvar2 - dvar3;
              CC(u, nv, m, v, fk, fj, fi + 1) = q + dvar1 - dvar2 - dvar3;
              if (I.multi_d) {
                CC(u, nv, m, v, fk, fj + 1, fi) = q - dvar1 + dvar2 - dvar3;
                CC(u, nv, m, v, fk, fj + 1, fi + 1) = q + dvar1 + dvar2 - dvar3;
              }
              if (I.three_d) {
                CC(u, nv, m, v, fk + 1, fj, fi) = q - dvar1 - dvar2 + dvar3;
                CC(u, nv, m, v, fk + 1, fj, fi + 1) = q + dvar1 - dvar2 + dvar3;
                CC(u, nv, m, v, fk + 1, fj + 1, fi) = q - dvar1 + dvar2 + dvar3;
                CC(u, nv, m, v, fk + 1, fj + 1, fi + 1) = q + dvar1 + dvar2 + dvar3;
              }
            }
    }
}

void oracle_smr_prolongate_cc(const ocfg *c, double *u, double *cu, int nv, const int *nghbr, const int *mblev) {
  oracle_smr_prolongate_cc_steps(c, u, cu, nv, nghbr, mblev, 3);
}

/* flux_correct_cc.cpp:28-167 + :173-290: fine blocks restrict their boundary-face fluxes onto the coarse
 * neighbour's face, which replaces its own flux there.  Face slots only (0-15, 24-31). */
void oracle_smr_flux_correct_cc(const ocfg *c, double *flx1, double *flx2, double *flx3, int nv, const int *nghbr,
                                const int *mblev) {
  odims D = dims(c);
  sidx I = smr_indcs(c);
  const int nmb = c->nmb;
  const int one_d = !I.multi_d, two_d = I.multi_d && !I.three_d;
  double **rbuf = calloc((size_t)nmb * 56, sizeof(double *));
  for (int m = 0; m < nmb; ++m)
    for (int n = 0; n < 56; ++n) {
      const int *e = nghbr + ((size_t)m * 56 + n) * 4;
      if (e[0] < 0 || e[1] >= mblev[m]) continue;
      if (!((n < 16) || (n >= 24 && n < 32))) continue;
      int ox, oy, oz, f1, f2;
      smr_slot_decode(n, &ox, &oy, &oz, &f1, &f2);
      sbox b = smr_send_box(&I, ox, oy, oz, f1, f2, SMR_FLUX);
      const int ni = b.ie - b.is + 1, nj = b.je - b.js + 1, nk = b.ke - b.ks + 1;
      double *buf = malloc(sizeof(double) * (size_t)ni * nj * nk * nv);
      for (int v = 0; v < nv; ++v) {
        if (n < 8) {
          const int fi = 2 * b.is - I.cis;
          for (int k = b.ks; k <= b.ke; ++k)
            for (int j = b.js; j <= b.je; ++j) {
              const int fj = 2 * j - I.cjs, fk = 2 * k - I.cks;
              double r;
              if (one_d) r = F1(flx1, m, v, 0, 0, fi);
              else if (two_d) r = 0.5 * (F1(flx1, m, v, 0, fj, fi) + F1(flx1, m, v, 0, fj + 1, fi));
              else r = 0.25 * (F1(flx1, m, v, fk, fj, fi) + F1(flx1, m, v, fk, fj + 1, fi) + F1(flx1, m, v, fk + 1, fj, fi) + F1(flx1, m, v, fk + 1, fj + 1, fi));
              buf[(j - b.js) + nj * ((k - b.ks) + nk * v)] = r;
            }
        } else if (n < 16) {
          const int fj = 2 * b.js - I.cjs;
          for (int k = b.ks; k <= b.ke; ++k)
            for (int i = b.is; i <= b.ie; ++i) {
              const int fi = 2 * i - I.cis, fk = 2 * k - I.cks;
              double r;
              if (two_d) r = 0.5 * (F2(flx2, m, v, 0, fj, fi) + F2(flx2, m, v, 0, fj, fi + 1));
              else r = 0.25 * (F2(flx2, m, v, fk, fj, fi) + F2(flx2, m, v, fk, fj, fi + 1) + F2(flx2, m, v, fk + 1, fj, fi) + F2(flx2, m, v, fk + 1, fj, fi + 1));
              buf[(i - b.is) + ni * ((k - b.ks) + nk * v)] = r;
            }
        } else {
          const int fk = 2 * b.ks - I.cks;
          for (int j = b.js; j <= b.je; ++j)
            for (int i = b.is; i <= b.ie; ++i) {
              const int fi = 2 * i - I.cis, fj = 2 * j - I.cjs;
              buf[(i - b.is) + ni * ((j - b.js) + nj * v)] =
                  0.25 * (F3(flx3, m, v, fk, fj, fi) + F3(flx3, m, v, fk, fj, fi + 1) + F3(flx3, m, v, fk, fj + 1, fi) + F3(flx3, m, v, fk, fj + 1, fi + 1));
            }
        }
      }
      rbuf[(size_t)e[0] * 56 + e[3]] = buf;
    }
  for (int m = 0; m < nmb; ++m)
    for (int n = 0; n < 56; ++n) {
      const int *e = nghbr + ((size_t)m * 56 + n) * 4;
      if (e[0] < 0 || e[1] <= mblev[m]) continue;
      if (!((n < 16) || (n >= 24 && n < 32))) continue;
      int ox, oy, oz, f1, f2;
      smr_slot_decode(n, &ox, &oy, &oz, &f1, &f2);
      sbox b = smr_recv_box(&I, ox, oy, oz, f1, f2, SMR_FLUX);
      const int ni = b.ie - b.is + 1, nj = b.je - b.js + 1, nk = b.ke - b.ks + 1;
      const double *buf = rbuf[(size_t)m * 56 + n];
      for (int v = 0; v < nv; ++v) {
        if (n < 8) {
          for (int k = b.ks; k <= b.ke; ++k)
            for (int j = b.js; j <= b.je; ++j) F1(flx1, m, v, k, j, b.is) = buf[(j - b.js) + nj * ((k - b.ks) + nk * v)];
        } else if (n < 16) {
          for (int k = b.ks; k <= b.ke; ++k)
            for (int i = b.is; i <= b.ie; ++i) F2(flx2, m, v, k, b.js, i) = buf[(i - b.is) + ni * ((k - b.ks) + nk * v)];
        } else {
          for (int j = b.js; j <= b.je; ++j)
            for (int i = b.is; i <= b.ie; ++i) F3(flx3, m, v, b.ks, j, i) = buf[(i - b.is) + ni * ((j - b.js) + nj * v)];
        }
      }
    }
  for (size_t q = 0; q < (size_t)nmb * 56; ++q) free(rbuf[q]);
  free(rbuf);
}

/* Driver::InitBoundaryValuesAndPrimitives for hydro on a multilevel mesh (driver.cpp:587-604); coarse-array BCs:
 * src/bvals/physics/hydro_bcs.cpp:51-67 (HydroBCsCoarse), between FillCoarseInBndryCC and ProlongateCC */
static void smr_hydro_ghosts_and_prims(const ocfg *c, double *u0, double *cu0, double *w0, const int *nghbr, const int *mblev) {
  odims D = dims(c);
  oracle_smr_restrict_cc(c, u0, cu0, D.nv);
  oracle_smr_halo_cc(c, u0, cu0, D.nv, nghbr, mblev);
  oracle_smr_prolongate_cc_steps(c, u0, cu0, D.nv, nghbr, mblev, 1);
  if (g_mesh_bcs) {   /* HydroBCsCoarse (hydro_bcs.cpp:51-67): the same fill on the coarse array's index space */
    ocfg cc = *c;
    sidx I = smr_indcs(c);
    cc.nx1 = I.cnx1; cc.nx2 = I.multi_d ? I.cnx2 : 1; cc.nx3 = I.three_d ? I.cnx3 : 1;
    oracle_hydro_apply_bcs(&cc, g_mesh_bcs, g_mb_bcs, cu0);
  }
  oracle_smr_prolongate_cc_steps(c, u0, cu0, D.nv, nghbr, mblev, 2);
  if (g_mesh_bcs) oracle_hydro_apply_bcs(c, g_mesh_bcs, g_mb_bcs, u0);     /* hydro_tasks.cpp:357-371, after prolongation */
  oracle_hydro_cons2prim(c, u0, w0, 0, D.n1 - 1, 0, D.n2 - 1, 0, D.n3 - 1, NULL);
}

double oracle_hydro_smr_init(const ocfg *c, double *u0, double *cu0, double *w0, const double *mb_size, const int *nghbr,
                             const int *mblev, double cfl, double tlim, double time) {
  smr_hydro_ghosts_and_prims(c, u0, cu0, w0, nghbr, mblev);
  double dtnew = oracle_hydro_new_dt(c, w0, mb_size, NULL);
  double d = 2.0 * (double)FLT_MAX;
  double lim = cfl * dtnew;
  if (lim < d) d = lim;
  if ((time < tlim) && ((time + d) > tlim)) d = tlim - time;
  return d;
}

/* stagen chain of hydro_tasks.cpp:55-71 on a multilevel, strictly periodic mesh */
int oracle_hydro_smr_run(const ocfg *c, double *u0, double *u1, double *cu0, double *w0, const double *mb_size,
                         const int *nghbr, const int *mblev, int integrator, double cfl, double tlim, int ncycles,
                         double *time, double *dt) {
  odims D = dims(c);
  size_t nm = (size_t)c->nmb;
  double g0[3] = {0.0, 0.0, 0.0}, g1[3] = {1.0, 0.0, 0.0}, be[3] = {1.0, 0.0, 0.0};
  if (integrator == 2) { g0[1] = 0.5; g1[1] = 0.5; be[1] = 0.5; }
  if (integrator == 3) { g0[1] = 0.25; g1[1] = 0.75; be[1] = 0.25; g0[2] = 2.0 / 3.0; g1[2] = 1.0 / 3.0; be[2] = 2.0 / 3.0; }
  double *flx1 = calloc(D.nv * nm * D.f1, 8), *flx2 = calloc(D.nv * nm * D.f2, 8), *flx3 = calloc(D.nv * nm * D.f3, 8);
  int n = 0;
  while (*time < tlim && (ncycles < 0 || n < ncycles)) {
    for (int st = 1; st <= integrator; ++st) {
      if (st == 1) memcpy(u1, u0, 8 * D.nv * nm * D.cc);
      oracle_hydro_fluxes(c, w0, flx1, flx2, flx3);
      oracle_smr_flux_correct_cc(c, flx1, flx2, flx3, D.nv, nghbr, mblev);
      oracle_hydro_rk_update(c, u0, u1, flx1, flx2, flx3, mb_size, g0[st - 1], g1[st - 1], be[st - 1] * (*dt));
      smr_hydro_ghosts_and_prims(c, u0, cu0, w0, nghbr, mblev);
    }
    double dtnew = oracle_hydro_new_dt(c, w0, mb_size, NULL);
    *time = *time + *dt;
    n++;
    double d = 2.0 * (*dt);
    double lim = cfl * dtnew;
    if (lim < d) d = lim;
    if ((*time < tlim) && ((*time + d) > tlim)) d = tlim - *time;
    *dt = d;
  }
  free(flx1); free(flx2); free(flx3);
  return n;
}
