/* oracle/hydro_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product).
 *
 * Plain-C, single-thread CPU restatement of the reference's Newtonian ideal-gas HYDRO stage
 * update (AthenaK /root/reference/src/hydro), the hydro twin of mhd_oracle.c.  #included at the
 * end of mhd_oracle.c (it reuses that file's reconstruction, ghost-exchange and BC routines).
 * 1-D, 2-D and 3-D MeshBlocks.  Every function cites the reference file:line it follows.
 *
 * Parity status: PINNED -- tests/test_oracle_cpu.py compares it bitwise with restart dumps of
 * the unmodified reference (oracle/_ref/athena_builtin: Sod shock tube, linear waves) and with
 * the committed fixtures tests/golden/hydro_*.npz generated from it.
 */

/* rsolver codes: 10 advect, 11 llf, 12 hlle, 13 hllc, 14 roe (hydro.hpp:32) */

static double sound_speed(double gamma, double d, double p) { return sqrt(gamma * p / d); } /* eos.hpp:41-45 */

typedef struct { double d, vx, vy, vz, e; } hprim1d;
typedef struct { double d, mx, my, mz, e; } hcons1d;

/* advect_hyd.hpp:18-52 */
static void hyd_advect(const hprim1d *l, const hprim1d *r, hcons1d *f) {
  if (l->vx >= 0.0) {
    f->d = l->d * l->vx; f->mx = l->d * l->vx * l->vx; f->my = l->vy * l->vx; f->mz = l->vz * l->vx;
    f->e = l->e * l->vx;
  } else {
    f->d = r->d * r->vx; f->mx = r->d * r->vx * r->vx; f->my = r->vy * r->vx; f->mz = r->vz * r->vx;
    f->e = r->e * r->vx;
  }
}

/* llf_hyd_singlestate.hpp:27-82 */
static void hyd_llf(const ocfg *c, const hprim1d *wl, const hprim1d *wr, hcons1d *flux) {
  double qa = wl->d * wl->vx;
  double qb = wr->d * wr->vx;
  hcons1d fsum, du;
  fsum.d = qa + qb;
  fsum.mx = qa * wl->vx + qb * wr->vx;
  fsum.my = qa * wl->vy + qb * wr->vy;
  fsum.mz = qa * wl->vz + qb * wr->vz;
  double el, er, pl, pr;
  pl = (c->gamma - 1.0) * wl->e;
  pr = (c->gamma - 1.0) * wr->e;
  el = wl->e + 0.5 * wl->d * (SQR(wl->vx) + SQR(wl->vy) + SQR(wl->vz));
  er = wr->e + 0.5 * wr->d * (SQR(wr->vx) + SQR(wr->vy) + SQR(wr->vz));
  fsum.mx += (pl + pr);
  fsum.e = (el + pl) * wl->vx + (er + pr) * wr->vx;
  qa = sound_speed(c->gamma, wl->d, pl);
  qb = sound_speed(c->gamma, wr->d, pr);
  double a = fmax((fabs(wl->vx) + qa), (fabs(wr->vx) + qb));
  du.d = a * (wr->d - wl->d);
  du.mx = a * (wr->d * wr->vx - wl->d * wl->vx);
  du.my = a * (wr->d * wr->vy - wl->d * wl->vy);
  du.mz = a * (wr->d * wr->vz - wl->d * wl->vz);
  du.e = a * (er - el);
  flux->d = 0.5 * (fsum.d - du.d);
  flux->mx = 0.5 * (fsum.mx - du.mx);
  flux->my = 0.5 * (fsum.my - du.my);
  flux->mz = 0.5 * (fsum.mz - du.mz);
  flux->e = 0.5 * (fsum.e - du.e);
}

/* hlle_hyd.hpp:26-129 */
static void hyd_hlle(const ocfg *c, const hprim1d *wl, const hprim1d *wr, hcons1d *flux) {
  const double gm1 = c->gamma - 1.0;
  const double igm1 = 1.0 / gm1;
  const double wl_ipr = gm1 * wl->e, wr_ipr = gm1 * wr->e;
  double sqrtdl = sqrt(wl->d);
  double sqrtdr = sqrt(wr->d);
  double isdlpdr = 1.0 / (sqrtdl + sqrtdr);
  double wroe_ivx = (sqrtdl * wl->vx + sqrtdr * wr->vx) * isdlpdr;
  double wroe_ivy = (sqrtdl * wl->vy + sqrtdr * wr->vy) * isdlpdr;
  double wroe_ivz = (sqrtdl * wl->vz + sqrtdr * wr->vz) * isdlpdr;
  double el = wl_ipr * igm1 + 0.5 * wl->d * (SQR(wl->vx) + SQR(wl->vy) + SQR(wl->vz));
  double er = wr_ipr * igm1 + 0.5 * wr->d * (SQR(wr->vx) + SQR(wr->vy) + SQR(wr->vz));
  double hroe = ((el + wl_ipr) / sqrtdl + (er + wr_ipr) / sqrtdr) * isdlpdr;
  double qa = sound_speed(c->gamma, wl->d, wl_ipr);
  double qb = sound_speed(c->gamma, wr->d, wr_ipr);
  double a = hroe - 0.5 * (SQR(wroe_ivx) + SQR(wroe_ivy) + SQR(wroe_ivz));
  a = (a < 0.0) ? 0.0 : sqrt(gm1 * a);
  double al = fmin((wroe_ivx - a), (wl->vx - qa));
  double ar = fmax((wroe_ivx + a), (wr->vx + qb));
  double bp = (ar > 0.0) ? ar : 1.0e-20;
  double bm = (al < 0.0) ? al : -1.0e-20;
  qa = wl->vx - bm;
  qb = wr->vx - bp;
  hcons1d fl, fr;
  fl.d = wl->d * qa;              fr.d = wr->d * qb;
  fl.mx = wl->d * wl->vx * qa;    fr.mx = wr->d * wr->vx * qb;
  fl.my = wl->d * wl->vy * qa;    fr.my = wr->d * wr->vy * qb;
  fl.mz = wl->d * wl->vz * qa;    fr.mz = wr->d * wr->vz * qb;
  fl.mx += wl_ipr;                fr.mx += wr_ipr;
  fl.e = el * qa + wl_ipr * wl->vx;
  fr.e = er * qb + wr_ipr * wr->vx;
  qa = 0.0;
  if (bp != bm) { qa = 0.5 * (bp + bm) / (bp - bm); }
  flux->d = 0.5 * (fl.d + fr.d) + qa * (fl.d - fr.d);
  flux->mx = 0.5 * (fl.mx + fr.mx) + qa * (fl.mx - fr.mx);
  flux->my = 0.5 * (fl.my + fr.my) + qa * (fl.my - fr.my);
  flux->mz = 0.5 * (fl.mz + fr.mz) + qa * (fl.mz - fr.mz);
  flux->e = 0.5 * (fl.e + fr.e) + qa * (fl.e - fr.e);
}

/* hllc_hyd.hpp:20-115 */
static void hyd_hllc(const ocfg *c, const hprim1d *wl, const hprim1d *wr, hcons1d *flux) {
  const double gm1 = c->gamma - 1.0;
  const double igm1 = 1.0 / gm1;
  const double alpha = ((c->gamma) + 1.0) / (2.0 * (c->gamma));
  double wl_idn = wl->d, wl_ivx = wl->vx, wl_ivy = wl->vy, wl_ivz = wl->vz;
  double wr_idn = wr->d, wr_ivx = wr->vx, wr_ivy = wr->vy, wr_ivz = wr->vz;
  double wl_ipr = gm1 * wl->e;
  double wr_ipr = gm1 * wr->e;
  double qa, qb, qc, qd, qe, qf;
  qa = sound_speed(c->gamma, wl_idn, wl_ipr);
  qb = sound_speed(c->gamma, wr_idn, wr_ipr);
  double el = wl_ipr * igm1 + 0.5 * wl_idn * (SQR(wl_ivx) + SQR(wl_ivy) + SQR(wl_ivz));
  double er = wr_ipr * igm1 + 0.5 * wr_idn * (SQR(wr_ivx) + SQR(wr_ivy) + SQR(wr_ivz));
  qc = 0.25 * (wl_idn + wr_idn) * (qa + qb);
  qd = 0.5 * (wl_ipr + wr_ipr + (wl_ivx - wr_ivx) * qc);
  qe = (qd <= wl_ipr) ? 1.0 : sqrt(1.0 + alpha * ((qd / wl_ipr) - 1.0));
  qf = (qd <= wr_ipr) ? 1.0 : sqrt(1.0 + alpha * ((qd / wr_ipr) - 1.0));
  qc = wl_ivx - qa * qe;
  qd = wr_ivx + qb * qf;
  qa = qd > 0.0 ? qd : 1.0e-20;
  qb = qc < 0.0 ? qc : -1.0e-20;
  qe = wl_ivx - qc;
  qf = wr_ivx - qd;
  qc = wl_ipr + qe * wl_idn * wl_ivx;
  qd = wr_ipr + qf * wr_idn * wr_ivx;
  double ml = wl_idn * qe;
  double mr = -(wr_idn * qf);
  double am = (qc - qd) / (ml + mr);
  double cp = (ml * qd + mr * qc) / (ml + mr);
  cp = cp > 0.0 ? cp : 0.0;
  qe = wl_idn * (wl_ivx - qb);
  qf = wr_idn * (wr_ivx - qa);
  double fl_d = qe, fr_d = qf;
  double fl_mx = qe * wl_ivx + wl_ipr, fr_mx = qf * wr_ivx + wr_ipr;
  double fl_my = qe * wl_ivy, fr_my = qf * wr_ivy;
  double fl_mz = qe * wl_ivz, fr_mz = qf * wr_ivz;
  double fl_e = el * (wl_ivx - qb) + wl_ipr * wl_ivx;
  double fr_e = er * (wr_ivx - qa) + wr_ipr * wr_ivx;
  if (am >= 0.0) {
    qc = am / (am - qb);
    qd = 0.0;
    qe = -qb / (am - qb);
  } else {
    qc = 0.0;
    qd = -am / (qa - am);
    qe = qa / (qa - am);
  }
  flux->d = qc * fl_d + qd * fr_d;
  flux->mx = qc * fl_mx + qd * fr_mx + qe * cp;
  flux->my = qc * fl_my + qd * fr_my;
  flux->mz = qc * fl_mz + qd * fr_mz;
  flux->e = qc * fl_e + qd * fr_e + qe * cp * am;
}

/* roe_hyd.hpp:187-283 (RoeFluxAdb) */
static void roe_flux_adb(const double wroe[], const double du[], const double wli[], double gm1,
                         double flx[], double ev[], int *llf_flag) {
  double v1 = wroe[IVX], v2 = wroe[IVY], v3 = wroe[IVZ];
  double h = wroe[IEN];
  double vsq = v1 * v1 + v2 * v2 + v3 * v3;
  double q = h - 0.5 * vsq;
  double cs_sq = (q < 0.0) ? (FLT_MIN) : gm1 * q;
  double cs = sqrt(cs_sq);
  ev[0] = v1 - cs; ev[1] = v1; ev[2] = v1; ev[3] = v1; ev[4] = v1 + cs;
  double a[5];
  double na = 0.5 / cs_sq;
  a[0] = du[0] * (0.5 * gm1 * vsq + v1 * cs);
  a[0] -= du[1] * (gm1 * v1 + cs);
  a[0] -= du[2] * gm1 * v2;
  a[0] -= du[3] * gm1 * v3;
  a[0] += du[4] * gm1;
  a[0] *= na;
  a[1] = du[0] * (-v2);
  a[1] += du[2];
  a[2] = du[0] * (-v3);
  a[2] += du[3];
  double qa = gm1 / cs_sq;
  a[3] = du[0] * (1.0 - na * gm1 * vsq);
  a[3] += du[1] * qa * v1;
  a[3] += du[2] * qa * v2;
  a[3] += du[3] * qa * v3;
  a[3] -= du[4] * qa;
  a[4] = du[0] * (0.5 * gm1 * vsq - v1 * cs);
  a[4] -= du[1] * (gm1 * v1 - cs);
  a[4] -= du[2] * gm1 * v2;
  a[4] -= du[3] * gm1 * v3;
  a[4] += du[4] * gm1;
  a[4] *= na;
  double coeff[5];
  for (int n = 0; n < 5; ++n) coeff[n] = -0.5 * fabs(ev[n]) * a[n];
  double dens = wli[IDN] + a[0];
  if (dens < 0.0) *llf_flag = 1;
  dens += a[3];
  if (dens < 0.0) *llf_flag = 1;
  flx[0] += coeff[0];
  flx[0] += coeff[3];
  flx[0] += coeff[4];
  flx[1] += coeff[0] * (v1 - cs);
  flx[1] += coeff[3] * v1;
  flx[1] += coeff[4] * (v1 + cs);
  flx[2] += coeff[0] * v2;
  flx[2] += coeff[1];
  flx[2] += coeff[3] * v2;
  flx[2] += coeff[4] * v2;
  flx[3] += coeff[0] * v3;
  flx[3] += coeff[2];
  flx[3] += coeff[3] * v3;
  flx[3] += coeff[4] * v3;
  flx[4] += coeff[0] * (h - v1 * cs);
  flx[4] += coeff[1] * v2;
  flx[4] += coeff[2] * v3;
  flx[4] += coeff[3] * 0.5 * vsq;
  flx[4] += coeff[4] * (h + v1 * cs);
}

/* roe_hyd.hpp:38-176 */
static void hyd_roe(const ocfg *c, const hprim1d *l, const hprim1d *r, hcons1d *flux) {
  const double gm1 = c->gamma - 1.0;
  double wli[5], wri[5], wroe[5], fl[5], fr[5], flxi[5], ev[5], du[5];
  wli[IDN] = l->d; wli[IVX] = l->vx; wli[IVY] = l->vy; wli[IVZ] = l->vz;
  wri[IDN] = r->d; wri[IVX] = r->vx; wri[IVY] = r->vy; wri[IVZ] = r->vz;
  wli[IEN] = gm1 * l->e;
  wri[IEN] = gm1 * r->e;
  double sqrtdl = sqrt(wli[IDN]);
  double sqrtdr = sqrt(wri[IDN]);
  double isdlpdr = 1.0 / (sqrtdl + sqrtdr);
  wroe[IDN] = sqrtdl * sqrtdr;
  wroe[IVX] = (sqrtdl * wli[IVX] + sqrtdr * wri[IVX]) * isdlpdr;
  wroe[IVY] = (sqrtdl * wli[IVY] + sqrtdr * wri[IVY]) * isdlpdr;
  wroe[IVZ] = (sqrtdl * wli[IVZ] + sqrtdr * wri[IVZ]) * isdlpdr;
  double el = wli[IEN] / gm1 + 0.5 * wli[IDN] * (SQR(wli[IVX]) + SQR(wli[IVY]) + SQR(wli[IVZ]));
  double er = wri[IEN] / gm1 + 0.5 * wri[IDN] * (SQR(wri[IVX]) + SQR(wri[IVY]) + SQR(wri[IVZ]));
  wroe[IEN] = ((el + wli[IEN]) / sqrtdl + (er + wri[IEN]) / sqrtdr) * isdlpdr;
  double mxl = wli[IDN] * wli[IVX];
  double mxr = wri[IDN] * wri[IVX];
  fl[IDN] = mxl;             fr[IDN] = mxr;
  fl[IVX] = mxl * wli[IVX];  fr[IVX] = mxr * wri[IVX];
  fl[IVY] = mxl * wli[IVY];  fr[IVY] = mxr * wri[IVY];
  fl[IVZ] = mxl * wli[IVZ];  fr[IVZ] = mxr * wri[IVZ];
  fl[IVX] += wli[IEN];       fr[IVX] += wri[IEN];
  fl[IEN] = (el + wli[IEN]) * wli[IVX];
  fr[IEN] = (er + wri[IEN]) * wri[IVX];
  du[IDN] = wri[IDN] - wli[IDN];
  du[IVX] = wri[IDN] * wri[IVX] - wli[IDN] * wli[IVX];
  du[IVY] = wri[IDN] * wri[IVY] - wli[IDN] * wli[IVY];
  du[IVZ] = wri[IDN] * wri[IVZ] - wli[IDN] * wli[IVZ];
  du[IEN] = er - el;
  for (int n = 0; n < 5; ++n) flxi[n] = 0.5 * (fl[n] + fr[n]);
  int llf_flag = 0;
  roe_flux_adb(wroe, du, wli, gm1, flxi, ev, &llf_flag);
  if (ev[0] >= 0.0) for (int n = 0; n < 5; ++n) flxi[n] = fl[n];
  if (ev[4] <= 0.0) for (int n = 0; n < 5; ++n) flxi[n] = fr[n];
  if (llf_flag != 0) {
    double cl = sound_speed(c->gamma, wli[IDN], wli[IEN]);
    double cr = sound_speed(c->gamma, wri[IDN], wri[IEN]);
    double a = 0.5 * fmax((fabs(wli[IVX]) + cl), (fabs(wri[IVX]) + cr));
    for (int n = 0; n < 5; ++n) flxi[n] = 0.5 * (fl[n] + fr[n]) - a * du[n];
  }
  flux->d = flxi[IDN]; flux->mx = flxi[IVX]; flux->my = flxi[IVY]; flux->mz = flxi[IVZ]; flux->e = flxi[IEN];
}

/* isothermal branches of llf_hyd_singlestate.hpp:27-82 */
static void hyd_llf_iso(const ocfg *c, const hprim1d *wl, const hprim1d *wr, hcons1d *flux) {
  double qa = wl->d * wl->vx;
  double qb = wr->d * wr->vx;
  hcons1d fsum, du;
  fsum.d = qa + qb;
  fsum.mx = qa * wl->vx + qb * wr->vx;
  fsum.my = qa * wl->vy + qb * wr->vy;
  fsum.mz = qa * wl->vz + qb * wr->vz;
  fsum.mx += SQR(c->iso_cs) * (wl->d + wr->d);
  qa = c->iso_cs;
  qb = c->iso_cs;
  double a = fmax((fabs(wl->vx) + qa), (fabs(wr->vx) + qb));
  du.d = a * (wr->d - wl->d);
  du.mx = a * (wr->d * wr->vx - wl->d * wl->vx);
  du.my = a * (wr->d * wr->vy - wl->d * wl->vy);
  du.mz = a * (wr->d * wr->vz - wl->d * wl->vz);
  flux->d = 0.5 * (fsum.d - du.d);
  flux->mx = 0.5 * (fsum.mx - du.mx);
  flux->my = 0.5 * (fsum.my - du.my);
  flux->mz = 0.5 * (fsum.mz - du.mz);
  flux->e = 0.0;
}

/* isothermal branches of hlle_hyd.hpp:26-129 */
static void hyd_hlle_iso(const ocfg *c, const hprim1d *wl, const hprim1d *wr, hcons1d *flux) {
  const double iso_cs = c->iso_cs;
  double sqrtdl = sqrt(wl->d);
  double sqrtdr = sqrt(wr->d);
  double isdlpdr = 1.0 / (sqrtdl + sqrtdr);
  double wroe_ivx = (sqrtdl * wl->vx + sqrtdr * wr->vx) * isdlpdr;
  double qa = iso_cs, qb = iso_cs, a = iso_cs;
  double al = fmin((wroe_ivx - a), (wl->vx - qa));
  double ar = fmax((wroe_ivx + a), (wr->vx + qb));
  double bp = (ar > 0.0) ? ar : 1.0e-20;
  double bm = (al < 0.0) ? al : -1.0e-20;
  qa = wl->vx - bm;
  qb = wr->vx - bp;
  hcons1d fl, fr;
  fl.d = wl->d * qa;              fr.d = wr->d * qb;
  fl.mx = wl->d * wl->vx * qa;    fr.mx = wr->d * wr->vx * qb;
  fl.my = wl->d * wl->vy * qa;    fr.my = wr->d * wr->vy * qb;
  fl.mz = wl->d * wl->vz * qa;    fr.mz = wr->d * wr->vz * qb;
  fl.mx += (iso_cs * iso_cs) * wl->d;
  fr.mx += (iso_cs * iso_cs) * wr->d;
  qa = 0.0;
  if (bp != bm) { qa = 0.5 * (bp + bm) / (bp - bm); }
  flux->d = 0.5 * (fl.d + fr.d) + qa * (fl.d - fr.d);
  flux->mx = 0.5 * (fl.mx + fr.mx) + qa * (fl.mx - fr.mx);
  flux->my = 0.5 * (fl.my + fr.my) + qa * (fl.my - fr.my);
  flux->mz = 0.5 * (fl.mz + fr.mz) + qa * (fl.mz - fr.mz);
  flux->e = 0.0;
}

/* roe_hyd.hpp:38-176 isothermal branches + RoeFluxIso :290-345 */
static void hyd_roe_iso(const ocfg *c, const hprim1d *l, const hprim1d *r, hcons1d *flux) {
  const double iso_cs = c->iso_cs;
  double wli[4], wri[4], fl[4], fr[4], flxi[4], ev[4], du[4];
  wli[IDN] = l->d; wli[IVX] = l->vx; wli[IVY] = l->vy; wli[IVZ] = l->vz;
  wri[IDN] = r->d; wri[IVX] = r->vx; wri[IVY] = r->vy; wri[IVZ] = r->vz;
  double sqrtdl = sqrt(wli[IDN]);
  double sqrtdr = sqrt(wri[IDN]);
  double isdlpdr = 1.0 / (sqrtdl + sqrtdr);
  double v1 = (sqrtdl * wli[IVX] + sqrtdr * wri[IVX]) * isdlpdr;
  double v2 = (sqrtdl * wli[IVY] + sqrtdr * wri[IVY]) * isdlpdr;
  double v3 = (sqrtdl * wli[IVZ] + sqrtdr * wri[IVZ]) * isdlpdr;
  double mxl = wli[IDN] * wli[IVX];
  double mxr = wri[IDN] * wri[IVX];
  fl[IDN] = mxl;             fr[IDN] = mxr;
  fl[IVX] = mxl * wli[IVX];  fr[IVX] = mxr * wri[IVX];
  fl[IVY] = mxl * wli[IVY];  fr[IVY] = mxr * wri[IVY];
  fl[IVZ] = mxl * wli[IVZ];  fr[IVZ] = mxr * wri[IVZ];
  fl[IVX] += (iso_cs * iso_cs) * wli[IDN];
  fr[IVX] += (iso_cs * iso_cs) * wri[IDN];
  du[IDN] = wri[IDN] - wli[IDN];
  du[IVX] = wri[IDN] * wri[IVX] - wli[IDN] * wli[IVX];
  du[IVY] = wri[IDN] * wri[IVY] - wli[IDN] * wli[IVY];
  du[IVZ] = wri[IDN] * wri[IVZ] - wli[IDN] * wli[IVZ];
  for (int n = 0; n < 4; ++n) flxi[n] = 0.5 * (fl[n] + fr[n]);
  int llf_flag = 0;
  {
    ev[0] = v1 - iso_cs; ev[1] = v1; ev[2] = v1; ev[3] = v1 + iso_cs;
    double a[4];
    a[0] = du[0] * (0.5 + 0.5 * v1 / iso_cs);
    a[0] -= du[1] * 0.5 / iso_cs;
    a[1] = du[0] * (-v2);
    a[1] += du[2];
    a[2] = du[0] * (-v3);
    a[2] += du[3];
    a[3] = du[0] * (0.5 - 0.5 * v1 / iso_cs);
    a[3] += du[1] * 0.5 / iso_cs;
    double coeff[4];
    for (int n = 0; n < 4; ++n) coeff[n] = -0.5 * fabs(ev[n]) * a[n];
    double dens = wli[IDN] + a[0];
    if (dens < 0.0) llf_flag = 1;
    dens += a[3];
    if (dens < 0.0) llf_flag = 1;
    flxi[0] += coeff[0];
    flxi[0] += coeff[3];
    flxi[1] += coeff[0] * (v1 - iso_cs);
    flxi[1] += coeff[3] * (v1 + iso_cs);
    flxi[2] += coeff[0] * v2;
    flxi[2] += coeff[1];
    flxi[2] += coeff[3] * v2;
    flxi[3] += coeff[0] * v3;
    flxi[3] += coeff[2];
    flxi[3] += coeff[3] * v3;
  }
  if (ev[0] >= 0.0) for (int n = 0; n < 4; ++n) flxi[n] = fl[n];
  if (ev[3] <= 0.0) for (int n = 0; n < 4; ++n) flxi[n] = fr[n];
  if (llf_flag != 0) {
    double a = 0.5 * fmax((fabs(wli[IVX]) + iso_cs), (fabs(wri[IVX]) + iso_cs));
    for (int n = 0; n < 4; ++n) flxi[n] = 0.5 * (fl[n] + fr[n]) - a * du[n];
  }
  flux->d = flxi[IDN]; flux->mx = flxi[IVX]; flux->my = flxi[IVY]; flux->mz = flxi[IVZ]; flux->e = 0.0;
}

/* Hydro::CalculateFluxes, hydro_fluxes.cpp:78-229 (no FOFC, no scalars): reconstruction over
 * [il1-1,iu1] then a Riemann solve on faces [il1,iu1]; no transverse extension; x2 only if
 * multi_d, x3 only if three_d. */
void oracle_hydro_fluxes(const ocfg *c, const double *w0, double *flx1, double *flx2, double *flx3) {
  odims D = dims(c);
  size_t ncc = (size_t)c->nmb * D.cc;
  double *wl = calloc(D.nv * ncc, 8), *wr = calloc(D.nv * ncc, 8);
  const int is = D.is, ie = D.ie, js = D.js, je = D.je, ks = D.ks, ke = D.ke;
  const int ndir = 1 + (c->nx2 > 1) + (c->nx3 > 1);
  for (int dir = 0; dir < ndir; ++dir) {
    int kl = ks, ku = ke, jl = js, ju = je, il = is, iu = ie;          /* recon cell box */
    int fkl = ks, fku = ke, fjl = js, fju = je, fil = is, fiu = ie;    /* face box       */
    if (dir == 0) { il = is - 1; iu = ie + 1; fiu = ie + 1; }
    else if (dir == 1) { jl = js - 1; ju = je + 1; fju = je + 1; }
    else { kl = ks - 1; ku = ke + 1; fku = ke + 1; }
    recon_pass(c, dir, D.nv, 1, w0, wl, wr, kl, ku, jl, ju, il, iu);
    const int ivx = IVX + dir, ivy = IVX + (dir + 1) % 3, ivz = IVX + (dir + 2) % 3;
    for (int m = 0; m < c->nmb; ++m)
      for (int k = fkl; k <= fku; ++k)
        for (int j = fjl; j <= fju; ++j)
          for (int i = fil; i <= fiu; ++i) {
            hprim1d l, r;
            hcons1d f;
            l.d = CC(wl, D.nv, m, IDN, k, j, i); l.vx = CC(wl, D.nv, m, ivx, k, j, i); l.vy = CC(wl, D.nv, m, ivy, k, j, i);
            l.vz = CC(wl, D.nv, m, ivz, k, j, i); l.e = c->is_iso ? 0.0 : CC(wl, D.nv, m, IEN, k, j, i);
            r.d = CC(wr, D.nv, m, IDN, k, j, i); r.vx = CC(wr, D.nv, m, ivx, k, j, i); r.vy = CC(wr, D.nv, m, ivy, k, j, i);
            r.vz = CC(wr, D.nv, m, ivz, k, j, i); r.e = c->is_iso ? 0.0 : CC(wr, D.nv, m, IEN, k, j, i);
            if (c->is_iso) {
              if (c->rsolver == 11) hyd_llf_iso(c, &l, &r, &f);
              else if (c->rsolver == 12) hyd_hlle_iso(c, &l, &r, &f);
              else hyd_roe_iso(c, &l, &r, &f);
            } else
            switch (c->rsolver) {
              case 10: hyd_advect(&l, &r, &f); break;
              case 11: hyd_llf(c, &l, &r, &f); break;
              case 12: hyd_hlle(c, &l, &r, &f); break;
              case 13: hyd_hllc(c, &l, &r, &f); break;
              default: hyd_roe(c, &l, &r, &f); break;
            }
            double f5[5];
            f5[IDN] = f.d; f5[ivx] = f.mx; f5[ivy] = f.my; f5[ivz] = f.mz; f5[IEN] = f.e;
            for (int n = 0; n < D.nfl; ++n) {
              if (dir == 0) F1(flx1, m, n, k, j, i) = f5[n];
              else if (dir == 1) F2(flx2, m, n, k, j, i) = f5[n];
              else F3(flx3, m, n, k, j, i) = f5[n];
            }
            for (int n = D.nfl; n < D.nv; ++n) {   /* scalar fluxes, hydro_fluxes.cpp:126-138,165-177,204-216 */
              double fd = f5[IDN];
              double v = fd * ((fd >= 0.0) ? CC(wl, D.nv, m, n, k, j, i) : CC(wr, D.nv, m, n, k, j, i));
              if (dir == 0) F1(flx1, m, n, k, j, i) = v;
              else if (dir == 1) F2(flx2, m, n, k, j, i) = v;
              else F3(flx3, m, n, k, j, i) = v;
            }
          }
  }
  free(wl); free(wr);
}

/* Hydro::RKUpdate, hydro_update.cpp:23-82 */
void oracle_hydro_rk_update(const ocfg *c, double *u0, const double *u1, const double *flx1,
                            const double *flx2, const double *flx3, const double *mb_size, double gam0,
                            double gam1, double beta_dt) {
  odims D = dims(c);
  const int multi_d = c->nx2 > 1, three_d = c->nx3 > 1;
  for (int m = 0; m < c->nmb; ++m) {
    const double dx1 = mb_size[m * 9 + 6], dx2 = mb_size[m * 9 + 7], dx3 = mb_size[m * 9 + 8];
    for (int n = 0; n < D.nv; ++n)
      for (int k = D.ks; k <= D.ke; ++k)
        for (int j = D.js; j <= D.je; ++j)
          for (int i = D.is; i <= D.ie; ++i) {
            double divf = (F1(flx1, m, n, k, j, i + 1) - F1(flx1, m, n, k, j, i)) / dx1;
            if (multi_d) divf += (F2(flx2, m, n, k, j + 1, i) - F2(flx2, m, n, k, j, i)) / dx2;
            if (three_d) divf += (F3(flx3, m, n, k + 1, j, i) - F3(flx3, m, n, k, j, i)) / dx3;
            CC(u0, D.nv, m, n, k, j, i) = gam0 * CC(u0, D.nv, m, n, k, j, i) + gam1 * CC(u1, D.nv, m, n, k, j, i) - beta_dt * divf;
          }
  }
}

/* IdealHydro::ConsToPrim, ideal_hyd.cpp:28-113 + SingleC2P_IdealHyd, ideal_c2p_hyd.hpp:20-67 */
void oracle_hydro_cons2prim(const ocfg *c, double *u0, double *w0, int il, int iu, int jl, int ju, int kl,
                            int ku, int *nfloor) {
  odims D = dims(c);
  const double gm1 = c->gamma - 1.0;
  int nd = 0, ne = 0, nt = 0;
  for (int m = 0; m < c->nmb; ++m)
    for (int k = kl; k <= ku; ++k)
      for (int j = jl; j <= ju; ++j)
        for (int i = il; i <= iu; ++i) {
          double ud = CC(u0, D.nv, m, IDN, k, j, i);
          double umx = CC(u0, D.nv, m, IM1, k, j, i), umy = CC(u0, D.nv, m, IM2, k, j, i), umz = CC(u0, D.nv, m, IM3, k, j, i);
          if (c->is_iso) {   /* isothermal_hyd.cpp:25-45,60-130 */
            if (ud < c->dfloor) { ud = c->dfloor; CC(u0, D.nv, m, IDN, k, j, i) = ud; nd++; }
            double dii = 1.0 / ud;
            CC(w0, D.nv, m, IDN, k, j, i) = ud; CC(w0, D.nv, m, IVX, k, j, i) = dii * umx;
            CC(w0, D.nv, m, IVY, k, j, i) = dii * umy; CC(w0, D.nv, m, IVZ, k, j, i) = dii * umz;
            for (int n = D.nfl; n < D.nv; ++n)   /* no scalar floor in the isothermal c2p */
              CC(w0, D.nv, m, n, k, j, i) = CC(u0, D.nv, m, n, k, j, i) / ud;
            continue;
          }
          double ue = CC(u0, D.nv, m, IEN, k, j, i);
          int dfl = 0, efl = 0, tfl = 0;
          double efloor = c->pfloor / (c->gamma - 1.0);
          if (ud < c->dfloor) { ud = c->dfloor; dfl = 1; }
          double wd = ud;
          double di = 1.0 / ud;
          double wvx = di * umx, wvy = di * umy, wvz = di * umz;
          double e_k = 0.5 * di * (SQR(umx) + SQR(umy) + SQR(umz));
          double we = (ue - e_k);
          if (we < efloor) { we = efloor; ue = efloor + e_k; efl = 1; }
          if (gm1 * we * di < c->tfloor) { we = wd * c->tfloor / gm1; ue = we + e_k; tfl = 1; }
          double spe_over_eps = gm1 / pow(wd, gm1);
          double spe = spe_over_eps * we * di;
          if (spe <= c->sfloor) { we = wd * c->sfloor / spe_over_eps; efl = 1; }
          if (dfl) { CC(u0, D.nv, m, IDN, k, j, i) = ud; nd++; }
          if (efl) { CC(u0, D.nv, m, IEN, k, j, i) = ue; ne++; }
          if (tfl) { CC(u0, D.nv, m, IEN, k, j, i) = ue; nt++; }
          CC(w0, D.nv, m, IDN, k, j, i) = wd; CC(w0, D.nv, m, IVX, k, j, i) = wvx; CC(w0, D.nv, m, IVY, k, j, i) = wvy;
          CC(w0, D.nv, m, IVZ, k, j, i) = wvz; CC(w0, D.nv, m, IEN, k, j, i) = we;
          for (int n = D.nfl; n < D.nv; ++n) {   /* ideal_hyd.cpp:95-101 */
            if (CC(u0, D.nv, m, n, k, j, i) < 0.0) CC(u0, D.nv, m, n, k, j, i) = 0.0;
            CC(w0, D.nv, m, n, k, j, i) = CC(u0, D.nv, m, n, k, j, i) / ud;
          }
        }
  if (nfloor) { nfloor[0] = nd; nfloor[1] = ne; nfloor[2] = nt; }
}

/* Hydro::NewTimeStep, hydro_newdt.cpp:31-138 (Newtonian dynamic branch): returns dtnew, dt3
 * (may be NULL) receives the three directional minima */
double oracle_hydro_new_dt(const ocfg *c, const double *w0, const double *mb_size, double *dt3) {
  odims D = dims(c);
  double dt1 = (double)FLT_MAX, dt2 = (double)FLT_MAX, dt3_ = (double)FLT_MAX;
  for (int m = 0; m < c->nmb; ++m)
    for (int k = D.ks; k <= D.ke; ++k)
      for (int j = D.js; j <= D.je; ++j)
        for (int i = D.is; i <= D.ie; ++i) {
          double cs = c->iso_cs;                                    /* hydro_newdt.cpp:107-112 */
          if (!c->is_iso) {
            double p = (c->gamma - 1.0) * CC(w0, D.nv, m, IEN, k, j, i);
            cs = sound_speed(c->gamma, CC(w0, D.nv, m, IDN, k, j, i), p);
          }
          double max_dv1 = fabs(CC(w0, D.nv, m, IVX, k, j, i)) + cs;
          double max_dv2 = fabs(CC(w0, D.nv, m, IVY, k, j, i)) + cs;
          double max_dv3 = fabs(CC(w0, D.nv, m, IVZ, k, j, i)) + cs;
          dt1 = fmin((mb_size[m * 9 + 6] / max_dv1), dt1);
          dt2 = fmin((mb_size[m * 9 + 7] / max_dv2), dt2);
          dt3_ = fmin((mb_size[m * 9 + 8] / max_dv3), dt3_);
        }
  if (dt3) { dt3[0] = dt1; dt3[1] = dt2; dt3[2] = dt3_; }
  double dtnew = dt1;
  if (c->nx2 > 1) dtnew = fmin(dtnew, dt2);
  if (c->nx3 > 1) dtnew = fmin(dtnew, dt3_);
  return dtnew;
}

/* HydroBCs only (hydro_tasks.cpp:357-371) */
int oracle_hydro_apply_bcs(const ocfg *c, const int *mesh_bcs, const int *mb_bcs, double *u0) {
  for (int q = 0; q < 6 * c->nmb; ++q)
    if (mb_bcs[q] == BC_INFLOW || mb_bcs[q] == BC_USER || mb_bcs[q] == BC_SHEAR) return 1;
  for (int d = 0; d < 3; ++d) {
    if ((d == 1 && c->nx2 <= 1) || (d == 2 && c->nx3 <= 1)) continue;
    if (mesh_bcs[2 * d] == BC_PERIODIC || (d == 0 && mesh_bcs[0] == BC_SHEAR)) continue;
    bc_array(c, u0, (c->is_iso ? 4 : 5) + c->nscalars, -1, d, mb_bcs);
  }
  return 0;
}

/* One RK stage = the "stagen" task chain of hydro_tasks.cpp:55-71 on a uniform mesh */
void oracle_hydro_stage(const ocfg *c, double *u0, double *u1, double *w0, const double *mb_size,
                        const int *nghbr_gid, int stage, double gam0, double gam1, double beta_dt) {
  odims D = dims(c);
  size_t nm = (size_t)c->nmb;
  if (stage == 1) memcpy(u1, u0, 8 * D.nv * nm * D.cc);   /* CopyCons hydro_tasks.cpp:137-141 */
  double *flx1 = calloc(D.nv * nm * D.f1, 8), *flx2 = calloc(D.nv * nm * D.f2, 8), *flx3 = calloc(D.nv * nm * D.f3, 8);
  oracle_hydro_fluxes(c, w0, flx1, flx2, flx3);
  oracle_hydro_rk_update(c, u0, u1, flx1, flx2, flx3, mb_size, gam0, gam1, beta_dt);
  oracle_halo_cc(c, u0, D.nv, nghbr_gid);
  if (g_mesh_bcs) oracle_hydro_apply_bcs(c, g_mesh_bcs, g_mb_bcs, u0);
  oracle_hydro_cons2prim(c, u0, w0, 0, D.n1 - 1, 0, D.n2 - 1, 0, D.n3 - 1, NULL);
  free(flx1); free(flx2); free(flx3);
}

/* Driver::Initialize for hydro: ghosts, BCs, c2p; returns the first dt (mesh.cpp:573-640) */
double oracle_hydro_init(const ocfg *c, double *u0, double *w0, const double *mb_size, const int *nghbr_gid,
                         double cfl, double tlim, double time) {
  odims D = dims(c);
  oracle_halo_cc(c, u0, D.nv, nghbr_gid);
  if (g_mesh_bcs) oracle_hydro_apply_bcs(c, g_mesh_bcs, g_mb_bcs, u0);
  oracle_hydro_cons2prim(c, u0, w0, 0, D.n1 - 1, 0, D.n2 - 1, 0, D.n3 - 1, NULL);
  double dtnew = oracle_hydro_new_dt(c, w0, mb_size, NULL);
  double d = 2.0 * (double)FLT_MAX;
  double lim = cfl * dtnew;
  if (lim < d) d = lim;
  if ((time < tlim) && ((time + d) > tlim)) d = tlim - time;
  return d;
}

/* Driver::Execute + Mesh::NewTimeStep for hydro (see oracle_run) */
int oracle_hydro_run(const ocfg *c, double *u0, double *u1, double *w0, const double *mb_size,
                     const int *nghbr_gid, int integrator, double cfl, double tlim, int ncycles, double *time,
                     double *dt, double *dt_hist) {
  double g0[3] = {0.0, 0.0, 0.0}, g1[3] = {1.0, 0.0, 0.0}, be[3] = {1.0, 0.0, 0.0};
  int nst = integrator;
  if (integrator == 2) { g0[1] = 0.5; g1[1] = 0.5; be[1] = 0.5; }
  if (integrator == 3) { g0[1] = 0.25; g1[1] = 0.75; be[1] = 0.25; g0[2] = 2.0 / 3.0; g1[2] = 1.0 / 3.0; be[2] = 2.0 / 3.0; }
  int n = 0;
  while (*time < tlim && (ncycles < 0 || n < ncycles)) {
    if (dt_hist) dt_hist[n] = *dt;
    for (int st = 1; st <= nst; ++st)
      oracle_hydro_stage(c, u0, u1, w0, mb_size, nghbr_gid, st, g0[st - 1], g1[st - 1], be[st - 1] * (*dt));
    double dtnew = oracle_hydro_new_dt(c, w0, mb_size, NULL);
    *time = *time + *dt;
    n++;
    double d = 2.0 * (*dt);
    double lim = cfl * dtnew;
    if (lim < d) d = lim;
    if ((*time < tlim) && ((*time + d) > tlim)) d = tlim - *time;
    *dt = d;
  }
  return n;
}
