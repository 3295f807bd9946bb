/* oracle/mhd_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product).
 *
 * Plain-C, single-thread CPU restatement of the reference's Newtonian ideal-MHD stage update
 * (AthenaK, /root/reference/src).  One function per reference task, same split-kernel
 * structure, same loop ranges and the same FP64 expression order, so that a build with
 * contraction disabled (-ffp-contract=off, no -mfma) is bitwise comparable with the reference
 * Kokkos::Serial build.  Every function cites the reference file:line it follows.
 *
 * Parity status: PINNED -- tests/test_oracle_vs_reference.py compares this file against
 * restart dumps of the unmodified reference (oracle/_ref, built by oracle/build_ref.sh) and
 * against the committed fixtures in tests/golden/ generated from it.
 *
 * Layout (athena.hpp:111-130,178-231): cell arrays (m,n,k,j,i) row-major, i fastest;
 * faces x1f:(m,k,j,i+1) x2f:(m,k,j+1,i) x3f:(m,k+1,j,i); edges x1e:(m,k+1,j+1,i)
 * x2e:(m,k+1,j,i+1) x3e:(m,k,j+1,i+1).  3-D only.
 */
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int nmb, ng, nx1, nx2, nx3;
  int recon;   /* 0 dc, 1 plm, 2 ppm4, 3 ppmx, 4 wenoz, 5 teno   (athena.hpp:75) */
  int rsolver; /* 0 llf, 1 hlle, 2 hlld                   (mhd.hpp:39)    */
  double gamma, dfloor, pfloor, tfloor, sfloor, sigma_max; /* eos.hpp:27-33 */
  int nscalars; /* passive scalars stored after the fluid variables (mhd.cpp / hydro.cpp:75) */
  int is_iso;   /* 1: isothermal EOS (eos/isothermal_mhd.cpp, isothermal_hyd.cpp): 4 fluid variables, no IEN */
  double iso_cs;
  int fofc;     /* 1: first-order flux correction (<mhd>/<hydro> fofc, mhd_fofc.cpp, hydro_fofc.cpp) */
} ocfg;

enum { IDN = 0, IM1 = 1, IM2 = 2, IM3 = 3, IEN = 4, IVX = 1, IVY = 2, IVZ = 3 };
#define SQR(x) ((x) * (x))

typedef struct {
  int n1, n2, n3, is, ie, js, je, ks, ke;
  int nfl;                   /* fluid variables: 5 (ideal gas) or 4 (isothermal) */
  int nv;                    /* nfl + nscalars: variables per cell-centred / flux array */
  size_t cc;                 /* cells per variable per block  */
  size_t f1, f2, f3;         /* faces per block               */
  size_t e1, e2, e3;         /* edges per block               */
} odims;

static odims dims(const ocfg *c) {
  odims d;
  /* mesh.cpp:285-333: inactive dimensions (nx = 1) have extent 1 and no ghost cells */
  const int multi_d = c->nx2 > 1, three_d = c->nx3 > 1;
  d.n1 = c->nx1 + 2 * c->ng; d.n2 = multi_d ? c->nx2 + 2 * c->ng : 1; d.n3 = three_d ? c->nx3 + 2 * c->ng : 1;
  d.is = c->ng; d.ie = d.is + c->nx1 - 1;
  d.js = multi_d ? c->ng : 0; d.je = d.js + c->nx2 - 1;
  d.ks = three_d ? c->ng : 0; d.ke = d.ks + c->nx3 - 1;
  d.nfl = c->is_iso ? 4 : 5;
  d.nv = d.nfl + c->nscalars;
  d.cc = (size_t)d.n3 * d.n2 * d.n1;
  d.f1 = (size_t)d.n3 * d.n2 * (d.n1 + 1);
  d.f2 = (size_t)d.n3 * (d.n2 + 1) * d.n1;
  d.f3 = (size_t)(d.n3 + 1) * d.n2 * d.n1;
  d.e1 = (size_t)(d.n3 + 1) * (d.n2 + 1) * d.n1;
  d.e2 = (size_t)(d.n3 + 1) * d.n2 * (d.n1 + 1);
  d.e3 = (size_t)d.n3 * (d.n2 + 1) * (d.n1 + 1);
  return d;
}

/* index helpers */
#define CC(a, nv, m, n, k, j, i) (a)[(((size_t)(m) * (nv) + (n)) * D.n3 + (k)) * D.n2 * D.n1 + (size_t)(j) * D.n1 + (i)]
#define S4(a, m, k, j, i) (a)[((size_t)(m) * D.n3 + (k)) * D.n2 * D.n1 + (size_t)(j) * D.n1 + (i)]
#define X1F(a, m, k, j, i) (a)[(((size_t)(m) * D.n3 + (k)) * D.n2 + (j)) * (D.n1 + 1) + (i)]
#define X2F(a, m, k, j, i) (a)[(((size_t)(m) * D.n3 + (k)) * (D.n2 + 1) + (j)) * D.n1 + (i)]
#define X3F(a, m, k, j, i) (a)[(((size_t)(m) * (D.n3 + 1) + (k)) * D.n2 + (j)) * D.n1 + (i)]
#define F1(a, m, n, k, j, i) (a)[((((size_t)(m) * D.nv + (n)) * D.n3 + (k)) * D.n2 + (j)) * (D.n1 + 1) + (i)]
#define F2(a, m, n, k, j, i) (a)[((((size_t)(m) * D.nv + (n)) * D.n3 + (k)) * (D.n2 + 1) + (j)) * D.n1 + (i)]
#define F3(a, m, n, k, j, i) (a)[((((size_t)(m) * D.nv + (n)) * (D.n3 + 1) + (k)) * D.n2 + (j)) * D.n1 + (i)]
#define X1E(a, m, k, j, i) (a)[(((size_t)(m) * (D.n3 + 1) + (k)) * (D.n2 + 1) + (j)) * D.n1 + (i)]
#define X2E(a, m, k, j, i) (a)[(((size_t)(m) * (D.n3 + 1) + (k)) * D.n2 + (j)) * (D.n1 + 1) + (i)]
#define X3E(a, m, k, j, i) (a)[(((size_t)(m) * D.n3 + (k)) * (D.n2 + 1) + (j)) * (D.n1 + 1) + (i)]

/* ------------------------------------------------------------------------------------ */
/* reconstruction: plm.hpp:21-37, ppm.hpp:44-77, ppm.hpp:87-183, wenoz.hpp:29-83          */

static void PLM(double qm1, double q, double qp1, double *ql_ip1, double *qr_i) {
  double dql = (q - qm1);
  double dqr = (qp1 - q);
  double dq2 = dql * dqr;
  double dqm = dq2 / (dql + dqr);
  if (dq2 <= 0.0) dqm = 0.0;
  *ql_ip1 = q + dqm;
  *qr_i = q - dqm;
}

static void PPM4(double qm2, double qm1, double q, double qp1, double qp2, double *ql_ip1,
                 double *qr_i) {
  double qlv = (7. * (q + qm1) - (qm2 + qp1)) / 12.0;
  double qrv = (7. * (q + qp1) - (qm1 + qp2)) / 12.0;
  qlv = fmax(qlv, fmin(q, qm1));
  qlv = fmin(qlv, fmax(q, qm1));
  qrv = fmax(qrv, fmin(q, qp1));
  qrv = fmin(qrv, fmax(q, qp1));
  double qc = qrv - q;
  double qd = qlv - q;
  if ((qc * qd) >= 0.0) {
    qlv = q;
    qrv = q;
  } else {
    if (fabs(qc) >= 2.0 * fabs(qd)) qrv = q - 2.0 * qd;
    if (fabs(qd) >= 2.0 * fabs(qc)) qlv = q - 2.0 * qc;
  }
  *ql_ip1 = qrv;
  *qr_i = qlv;
}

#define SIGN(x) (((x) < 0.0) ? -1.0 : 1.0) /* athena.hpp SIGN macro */

static void PPMX(double qm2, double qm1, double q, double qp1, double qp2, double *ql_ip1,
                 double *qr_i) {
  double qlv = (7. * (q + qm1) - (qm2 + qp1)) / 12.0;
  double qrv = (7. * (q + qp1) - (qm1 + qp2)) / 12.0;
  double d2qc = 3.0 * ((qm1 + q) - 2.0 * qlv);
  double d2ql = (qm2 + q) - 2.0 * qm1;
  double d2qr = (qm1 + qp1) - 2.0 * q;
  double d2qlim = 0.0;
  double lim_slope = fmin(fabs(d2ql), fabs(d2qr));
  if (d2qc > 0.0 && d2ql > 0.0 && d2qr > 0.0) d2qlim = SIGN(d2qc) * fmin(1.25 * lim_slope, fabs(d2qc));
  if (d2qc < 0.0 && d2ql < 0.0 && d2qr < 0.0) d2qlim = SIGN(d2qc) * fmin(1.25 * lim_slope, fabs(d2qc));
  if (((qm1 - qlv) * (q - qlv)) > 0.0) qlv = 0.5 * (q + qm1) - d2qlim / 6.0;
  d2qc = 3.0 * ((q + qp1) - 2.0 * qrv);
  d2ql = d2qr;
  d2qr = (q + qp2) - 2.0 * qp1;
  d2qlim = 0.0;
  lim_slope = fmin(fabs(d2ql), fabs(d2qr));
  if (d2qc > 0.0 && d2ql > 0.0 && d2qr > 0.0) d2qlim = SIGN(d2qc) * fmin(1.25 * lim_slope, fabs(d2qc));
  if (d2qc < 0.0 && d2ql < 0.0 && d2qr < 0.0) d2qlim = SIGN(d2qc) * fmin(1.25 * lim_slope, fabs(d2qc));
  if (((q - qrv) * (qp1 - qrv)) > 0.0) qrv = 0.5 * (q + qp1) - d2qlim / 6.0;
  double qa = (qrv - q) * (q - qlv);
  double qb = (qm1 - q) * (q - qp1);
  if (qa <= 0.0 || qb <= 0.0) {
    double d2q = 6.0 * (qlv + qrv - 2.0 * q);
    double e2qc = (qm1 + qp1) - 2.0 * q;
    double e2ql = (qm2 + q) - 2.0 * qm1;
    double e2qr = (q + qp2) - 2.0 * qp1;
    d2qlim = 0.0;
    lim_slope = fmin(fabs(e2ql), fabs(e2qr));
    lim_slope = fmin(fabs(e2qc), lim_slope);
    if (e2qc > 0.0 && e2ql > 0.0 && e2qr > 0.0 && d2q > 0.0) d2qlim = SIGN(d2q) * fmin(1.25 * lim_slope, fabs(d2q));
    if (e2qc < 0.0 && e2ql < 0.0 && e2qr < 0.0 && d2q < 0.0) d2qlim = SIGN(d2q) * fmin(1.25 * lim_slope, fabs(d2q));
    double rho = 0.0;
    if (fabs(d2q) > (1.0e-12) * fmax(fabs(qm1), fmax(fabs(q), fabs(qp1)))) rho = d2qlim / d2q;
    qlv = q + (qlv - q) * rho;
    qrv = q + (qrv - q) * rho;
  } else {
    double qc = qrv - q;
    double qd = qlv - q;
    if (fabs(qc) >= 2.0 * fabs(qd)) qrv = q - 2.0 * qd;
    if (fabs(qd) >= 2.0 * fabs(qc)) qlv = q - 2.0 * qc;
  }
  *ql_ip1 = qrv;
  *qr_i = qlv;
}

static void WENOZ(double qm2, double qm1, double q, double qp1, double qp2, double *ql_ip1,
                  double *qr_i) {
  const double bc0 = 13. / 12., bc1 = 0.25;
  double beta[3];
  beta[0] = bc0 * SQR(qm2 + q - 2.0 * qm1) + bc1 * SQR(qm2 + 3.0 * q - 4.0 * qm1);
  beta[1] = bc0 * SQR(qm1 + qp1 - 2.0 * q) + bc1 * SQR(qm1 - qp1);
  beta[2] = bc0 * SQR(qp2 + q - 2.0 * qp1) + bc1 * SQR(qp2 + 3.0 * q - 4.0 * qp1);
  const double epsL = 1.0e-42;
  const double tau_5 = fabs(beta[0] - beta[2]);
  double ind[3];
  ind[0] = SQR(tau_5 / (beta[0] + epsL));
  ind[1] = SQR(tau_5 / (beta[1] + epsL));
  ind[2] = SQR(tau_5 / (beta[2] + epsL));
  double f[3];
  f[0] = (2.0 * qm2 - 7.0 * qm1 + 11.0 * q);
  f[1] = (-1.0 * qm1 + 5.0 * q + 2.0 * qp1);
  f[2] = (2.0 * q + 5.0 * qp1 - qp2);
  double alpha[3];
  alpha[0] = 0.1 * (1.0 + ind[0]);
  alpha[1] = 0.6 * (1.0 + ind[1]);
  alpha[2] = 0.3 * (1.0 + ind[2]);
  double alpha_sum = 6.0 * (alpha[0] + alpha[1] + alpha[2]);
  *ql_ip1 = (f[0] * alpha[0] + f[1] * alpha[1] + f[2] * alpha[2]) / alpha_sum;
  f[0] = (2.0 * qp2 - 7.0 * qp1 + 11.0 * q);
  f[1] = (-1.0 * qp1 + 5.0 * q + 2.0 * qm1);
  f[2] = (2.0 * q + 5.0 * qm1 - qm2);
  alpha[0] = 0.1 * (1.0 + ind[2]);
  alpha[2] = 0.3 * (1.0 + ind[0]);
  alpha_sum = 6.0 * (alpha[0] + alpha[1] + alpha[2]);
  *qr_i = (f[0] * alpha[0] + f[1] * alpha[1] + f[2] * alpha[2]) / alpha_sum;
}

/* teno.hpp:29-91 */
#define CUB(x) ((x) * (x) * (x))
static void TENO(double qm2, double qm1, double q, double qp1, double qp2, double *ql_ip1,
                 double *qr_i) {
  const double bc0 = 13. / 12., bc1 = 0.25;
  double beta[3];
  beta[0] = bc0 * SQR(qm2 + q - 2.0 * qm1) + bc1 * SQR(qm2 + 3.0 * q - 4.0 * qm1);
  beta[1] = bc0 * SQR(qm1 + qp1 - 2.0 * q) + bc1 * SQR(qm1 - qp1);
  beta[2] = bc0 * SQR(qp2 + q - 2.0 * qp1) + bc1 * SQR(qp2 + 3.0 * q - 4.0 * qp1);
  const double epsT = 1.0e-40, cT = 1.0e-6;
  double alpha[3];
  alpha[0] = 1.0 / SQR(CUB(beta[0] + epsT));
  alpha[1] = 1.0 / SQR(CUB(beta[1] + epsT));
  alpha[2] = 1.0 / SQR(CUB(beta[2] + epsT));
  double alpha_sum = alpha[0] + alpha[1] + alpha[2];
  double ind[3];
  ind[0] = (alpha[0] < cT * alpha_sum ? 0.0 : 1.0);
  ind[1] = (alpha[1] < cT * alpha_sum ? 0.0 : 1.0);
  ind[2] = (alpha[2] < cT * alpha_sum ? 0.0 : 1.0);
  double f[3];
  f[0] = (2.0 * qm2 - 7.0 * qm1 + 11.0 * q);
  f[1] = (-1.0 * qm1 + 5.0 * q + 2.0 * qp1);
  f[2] = (2.0 * q + 5.0 * qp1 - qp2);
  alpha[0] = 0.1 * ind[0];
  alpha[1] = 0.6 * ind[1];
  alpha[2] = 0.3 * ind[2];
  alpha_sum = 6.0 * (alpha[0] + alpha[1] + alpha[2]);
  *ql_ip1 = (f[0] * alpha[0] + f[1] * alpha[1] + f[2] * alpha[2]) / alpha_sum;
  f[0] = (2.0 * qp2 - 7.0 * qp1 + 11.0 * q);
  f[1] = (-1.0 * qp1 + 5.0 * q + 2.0 * qm1);
  f[2] = (2.0 * q + 5.0 * qm1 - qm2);
  alpha[0] = 0.1 * ind[2];
  alpha[2] = 0.3 * ind[0];
  alpha_sum = 6.0 * (alpha[0] + alpha[1] + alpha[2]);
  *qr_i = (f[0] * alpha[0] + f[1] * alpha[1] + f[2] * alpha[2]) / alpha_sum;
}

/* recon.hpp:40-118 (ReconCellT) + :134-185 (ReconDispatch): one variable of q over a cell
 * box; ql -> wl at the upper face index, qr -> wr at the cell's own index. */
static void recon_pass(const ocfg *c, int dir, int nv, int apply_floors, const double *q,
                       double *wl, double *wr, int kl, int ku, int jl, int ju, int il, int iu) {
  odims D = dims(c);
  const int di = (dir == 0), dj = (dir == 1), dk = (dir == 2);
  const double dfloor = c->dfloor;
  const double efloor = c->pfloor / (c->gamma - 1.0);
  for (int m = 0; m < c->nmb; ++m)
    for (int n = 0; n < nv; ++n)
      for (int k = kl; k <= ku; ++k)
        for (int j = jl; j <= ju; ++j)
          for (int i = il; i <= iu; ++i) {
            double qm2 = 0, qp2 = 0;
            double qm1 = CC(q, nv, m, n, k - dk, j - dj, i - di);
            double q0 = CC(q, nv, m, n, k, j, i);
            double qp1 = CC(q, nv, m, n, k + dk, j + dj, i + di);
            if (c->recon >= 2) {
              qm2 = CC(q, nv, m, n, k - 2 * dk, j - 2 * dj, i - 2 * di);
              qp2 = CC(q, nv, m, n, k + 2 * dk, j + 2 * dj, i + 2 * di);
            }
            double ql, qr;
            switch (c->recon) {
              case 1: PLM(qm1, q0, qp1, &ql, &qr); break;
              case 2: PPM4(qm2, qm1, q0, qp1, qp2, &ql, &qr); break;
              case 3: PPMX(qm2, qm1, q0, qp1, qp2, &ql, &qr); break;
              case 4: WENOZ(qm2, qm1, q0, qp1, qp2, &ql, &qr); break;
              case 5: TENO(qm2, qm1, q0, qp1, qp2, &ql, &qr); break;
              default: ql = q0; qr = q0; break;
            }
            if (apply_floors && c->recon >= 3) { /* recon.hpp:78-108: ppmx/wenoz/teno only */
              if (n == IDN) { ql = fmax(ql, dfloor); qr = fmax(qr, dfloor); }
              if (n == IEN && !c->is_iso) { ql = fmax(ql, efloor); qr = fmax(qr, efloor); }   /* recon.hpp:80: is_ideal */
            }
            CC(wl, nv, m, n, k + dk, j + dj, i + di) = ql;
            CC(wr, nv, m, n, k, j, i) = qr;
          }
}

/* ------------------------------------------------------------------------------------ */
/* eos.hpp:49-57 */
static double fast_speed(double gamma, double d, double p, double bx, double by, double bz) {
  double asq = gamma * p;
  double ct2 = by * by + bz * bz;
  double qsq = bx * bx + ct2 + asq;
  double tmp = bx * bx + ct2 - asq;
  return sqrt(0.5 * (qsq + sqrt(tmp * tmp + 4.0 * asq * ct2)) / d);
}

typedef struct { double d, mx, my, mz, e, by, bz; } cons1d;
typedef struct { double d, vx, vy, vz, e, by, bz; } prim1d;

/* llf_mhd_singlestate.hpp:27-89 (ideal branch).  NB flux.by is returned already negated. */
static void llf_face(const ocfg *c, const prim1d *wl, const prim1d *wr, double bxi, cons1d *flux) {
  const double gm1 = c->gamma - 1.0;
  double qa = wl->d * wl->vx;
  double qb = wr->d * wr->vx;
  double qc = 0.5 * (SQR(wl->by) + SQR(wl->bz) - SQR(bxi));
  double qd = 0.5 * (SQR(wr->by) + SQR(wr->bz) - SQR(bxi));
  cons1d fsum;
  fsum.d = qa + qb;
  fsum.mx = qa * wl->vx + qb * wr->vx + qc + qd;
  fsum.my = qa * wl->vy + qb * wr->vy - bxi * (wl->by + wr->by);
  fsum.mz = qa * wl->vz + qb * wr->vz - bxi * (wl->bz + wr->bz);
  fsum.by = wl->by * wl->vx + wr->by * wr->vx - bxi * (wl->vy + wr->vy);
  fsum.bz = wl->bz * wl->vx + wr->bz * wr->vx - bxi * (wl->vz + wr->vz);
  double pl = gm1 * wl->e;
  double pr = gm1 * wr->e;
  double el = wl->e + 0.5 * wl->d * (SQR(wl->vx) + SQR(wl->vy) + SQR(wl->vz)) + qc + SQR(bxi);
  double er = wr->e + 0.5 * wr->d * (SQR(wr->vx) + SQR(wr->vy) + SQR(wr->vz)) + qd + SQR(bxi);
  fsum.mx += (pl + pr);
  fsum.e = (el + pl + qc) * wl->vx + (er + pr + qd) * wr->vx;
  fsum.e -= bxi * (wl->by * wl->vy + wl->bz * wl->vz);
  fsum.e -= bxi * (wr->by * wr->vy + wr->bz * wr->vz);
  qa = fast_speed(c->gamma, wl->d, pl, bxi, wl->by, wl->bz);
  qb = fast_speed(c->gamma, wr->d, pr, bxi, wr->by, wr->bz);
  double a = fmax((fabs(wl->vx) + qa), (fabs(wr->vx) + qb));
  cons1d du;
  du.d = a * (wr->d - wl->d);
  du.mx = a * (wr->d * wr->vx - wl->d * wl->vx);
  du.my = a * (wr->d * wr->vy - wl->d * wl->vy);
  du.mz = a * (wr->d * wr->vz - wl->d * wl->vz);
  du.e = a * (er - el);
  du.by = a * (wr->by - wl->by);
  du.bz = a * (wr->bz - wl->bz);
  flux->d = 0.5 * (fsum.d - du.d);
  flux->mx = 0.5 * (fsum.mx - du.mx);
  flux->my = 0.5 * (fsum.my - du.my);
  flux->mz = 0.5 * (fsum.mz - du.mz);
  flux->e = 0.5 * (fsum.e - du.e);
  flux->by = -0.5 * (fsum.by - du.by);
  flux->bz = 0.5 * (fsum.bz - du.bz);
}

/* hlle_mhd.hpp:24-179 (ideal branch).  Returns F(By),F(Bz) un-negated in flux->by/bz via
 * the caller convention below (ey = -..., ez = +...) exactly as lines 176-177. */
static void hlle_face(const ocfg *c, const prim1d *wl, const prim1d *wr, double bxi, cons1d *flux,
                      double *ey, double *ez) {
  const double gm1 = c->gamma - 1.0;
  const double igm1 = 1.0 / gm1;
  double wl_ipr = gm1 * wl->e, wr_ipr = gm1 * wr->e;
  double sqrtdl = sqrt(wl->d);
  double sqrtdr = sqrt(wr->d);
  double isdlpdr = 1.0 / (sqrtdl + sqrtdr);
  double wroe_idn = sqrtdl * sqrtdr;
  double wroe_ivx = (sqrtdl * wl->vx + sqrtdr * wr->vx) * isdlpdr;
  double wroe_ivy = (sqrtdl * wl->vy + sqrtdr * wr->vy) * isdlpdr;
  double wroe_ivz = (sqrtdl * wl->vz + sqrtdr * wr->vz) * isdlpdr;
  double wroe_iby = (sqrtdr * wl->by + sqrtdl * wr->by) * isdlpdr;
  double wroe_ibz = (sqrtdr * wl->bz + sqrtdl * wr->bz) * isdlpdr;
  double x = 0.5 * (SQR(wl->by - wr->by) + SQR(wl->bz - wr->bz)) / (SQR(sqrtdl + sqrtdr));
  double y = 0.5 * (wl->d + wr->d) / wroe_idn;
  double pbl = 0.5 * (bxi * bxi + SQR(wl->by) + SQR(wl->bz));
  double pbr = 0.5 * (bxi * bxi + SQR(wr->by) + SQR(wr->bz));
  double el = wl_ipr * igm1 + 0.5 * wl->d * (SQR(wl->vx) + SQR(wl->vy) + SQR(wl->vz)) + pbl;
  double er = wr_ipr * igm1 + 0.5 * wr->d * (SQR(wr->vx) + SQR(wr->vy) + SQR(wr->vz)) + pbr;
  double hroe = ((el + wl_ipr + pbl) / sqrtdl + (er + wr_ipr + pbr) / sqrtdr) * isdlpdr;
  double cl = fast_speed(c->gamma, wl->d, wl_ipr, bxi, wl->by, wl->bz);
  double cr = fast_speed(c->gamma, wr->d, wr_ipr, bxi, wr->by, wr->bz);
  double btsq = SQR(wroe_iby) + SQR(wroe_ibz);
  double vaxsq = bxi * bxi / wroe_idn;
  double bt_starsq = (gm1 - (gm1 - 1.0) * y) * btsq;
  double hp = hroe - (vaxsq + btsq / wroe_idn);
  double vsq = SQR(wroe_ivx) + SQR(wroe_ivy) + SQR(wroe_ivz);
  double twid_asq = fmax((gm1 * (hp - 0.5 * vsq) - (gm1 - 1.0) * x), 0.0);
  double ct2 = bt_starsq / wroe_idn;
  double tsum = vaxsq + ct2 + twid_asq;
  double tdif = vaxsq + ct2 - twid_asq;
  double cf2_cs2 = sqrt(tdif * tdif + 4.0 * twid_asq * ct2);
  double cfsq = 0.5 * (tsum + cf2_cs2);
  double a = sqrt(cfsq);
  double al = fmin((wroe_ivx - a), (wl->vx - cl));
  double ar = fmax((wroe_ivx + a), (wr->vx + cr));
  double bp = ar > 0.0 ? ar : 1.0e-20;
  double bm = al < 0.0 ? al : -1.0e-20;
  double vxl = wl->vx - bm;
  double vxr = wr->vx - bp;
  cons1d fl, fr;
  fl.d = wl->d * vxl;
  fr.d = wr->d * vxr;
  fl.mx = wl->d * wl->vx * vxl + pbl - SQR(bxi);
  fr.mx = wr->d * wr->vx * vxr + pbr - SQR(bxi);
  fl.my = wl->d * wl->vy * vxl - bxi * wl->by;
  fr.my = wr->d * wr->vy * vxr - bxi * wr->by;
  fl.mz = wl->d * wl->vz * vxl - bxi * wl->bz;
  fr.mz = wr->d * wr->vz * vxr - bxi * wr->bz;
  fl.mx += wl_ipr;
  fr.mx += wr_ipr;
  fl.e = el * vxl + wl->vx * (wl_ipr + pbl - bxi * bxi);
  fr.e = er * vxr + wr->vx * (wr_ipr + pbr - bxi * bxi);
  fl.e -= bxi * (wl->by * wl->vy + wl->bz * wl->vz);
  fr.e -= bxi * (wr->by * wr->vy + wr->bz * wr->vz);
  fl.by = wl->by * vxl - bxi * wl->vy;
  fr.by = wr->by * vxr - bxi * wr->vy;
  fl.bz = wl->bz * vxl - bxi * wl->vz;
  fr.bz = wr->bz * vxr - bxi * wr->vz;
  double tmp = 0.0;
  if (bp != bm) tmp = 0.5 * (bp + bm) / (bp - bm);
  flux->d = 0.5 * (fl.d + fr.d) + (fl.d - fr.d) * tmp;
  flux->mx = 0.5 * (fl.mx + fr.mx) + (fl.mx - fr.mx) * tmp;
  flux->my = 0.5 * (fl.my + fr.my) + (fl.my - fr.my) * tmp;
  flux->mz = 0.5 * (fl.mz + fr.mz) + (fl.mz - fr.mz) * tmp;
  flux->e = 0.5 * (fl.e + fr.e) + (fl.e - fr.e) * tmp;
  *ey = -0.5 * (fl.by + fr.by) - (fl.by - fr.by) * tmp;
  *ez = 0.5 * (fl.bz + fr.bz) + (fl.bz - fr.bz) * tmp;
}

#define HLLD_SMALL_NUMBER 1.0e-4
/* hlld_mhd.hpp:41-348 (ideal branch) */
static void hlld_face(const ocfg *c, const prim1d *wl, const prim1d *wr, double bxi, cons1d *flxi) {
  double spd[5];
  double gm1 = c->gamma - 1.0;
  double igm1 = 1.0 / gm1;
  double wl_idn = wl->d, wl_ivx = wl->vx, wl_ivy = wl->vy, wl_ivz = wl->vz, wl_iby = wl->by, wl_ibz = wl->bz;
  double wr_idn = wr->d, wr_ivx = wr->vx, wr_ivy = wr->vy, wr_ivz = wr->vz, wr_iby = wr->by, wr_ibz = wr->bz;
  double wl_ipr = gm1 * wl->e;
  double wr_ipr = gm1 * wr->e;
  double bxsq = bxi * bxi;
  double pbl = 0.5 * (bxsq + (SQR(wl_iby) + SQR(wl_ibz)));
  double pbr = 0.5 * (bxsq + (SQR(wr_iby) + SQR(wr_ibz)));
  double kel = 0.5 * wl_idn * (SQR(wl_ivx) + (SQR(wl_ivy) + SQR(wl_ivz)));
  double ker = 0.5 * wr_idn * (SQR(wr_ivx) + (SQR(wr_ivy) + SQR(wr_ivz)));
  cons1d ul, ur;
  ul.d = wl_idn; ul.mx = wl_ivx * ul.d; ul.my = wl_ivy * ul.d; ul.mz = wl_ivz * ul.d;
  ul.e = wl_ipr * igm1 + kel + pbl; ul.by = wl_iby; ul.bz = wl_ibz;
  ur.d = wr_idn; ur.mx = wr_ivx * ur.d; ur.my = wr_ivy * ur.d; ur.mz = wr_ivz * ur.d;
  ur.e = wr_ipr * igm1 + ker + pbr; ur.by = wr_iby; ur.bz = wr_ibz;
  double cfl = fast_speed(c->gamma, wl_idn, wl_ipr, bxi, wl_iby, wl_ibz);
  double cfr = fast_speed(c->gamma, wr_idn, wr_ipr, bxi, wr_iby, wr_ibz);
  spd[0] = fmin(wl_ivx - cfl, wr_ivx - cfr);
  spd[4] = fmax(wl_ivx + cfl, wr_ivx + cfr);
  double ptl = wl_ipr + pbl;
  double ptr = wr_ipr + pbr;
  cons1d fl, fr;
  fl.d = ul.mx;
  fl.mx = ul.mx * wl_ivx + ptl - bxsq;
  fl.my = ul.my * wl_ivx - bxi * ul.by;
  fl.mz = ul.mz * wl_ivx - bxi * ul.bz;
  fl.e = wl_ivx * (ul.e + ptl - bxsq) - bxi * (wl_ivy * ul.by + wl_ivz * ul.bz);
  fl.by = ul.by * wl_ivx - bxi * wl_ivy;
  fl.bz = ul.bz * wl_ivx - bxi * wl_ivz;
  fr.d = ur.mx;
  fr.mx = ur.mx * wr_ivx + ptr - bxsq;
  fr.my = ur.my * wr_ivx - bxi * ur.by;
  fr.mz = ur.mz * wr_ivx - bxi * ur.bz;
  fr.e = wr_ivx * (ur.e + ptr - bxsq) - bxi * (wr_ivy * ur.by + wr_ivz * ur.bz);
  fr.by = ur.by * wr_ivx - bxi * wr_ivy;
  fr.bz = ur.bz * wr_ivx - bxi * wr_ivz;
  double sdl = spd[0] - wl_ivx;
  double sdr = spd[4] - wr_ivx;
  spd[2] = (sdr * ur.mx - sdl * ul.mx + (ptl - ptr)) / (sdr * ur.d - sdl * ul.d);
  double sdml = spd[0] - spd[2];
  double sdmr = spd[4] - spd[2];
  double sdml_inv = 1.0 / sdml;
  double sdmr_inv = 1.0 / sdmr;
  cons1d ulst, uldst, urdst, urst;
  ulst.d = ul.d * sdl * sdml_inv;
  urst.d = ur.d * sdr * sdmr_inv;
  double ulst_d_inv = 1.0 / ulst.d;
  double urst_d_inv = 1.0 / urst.d;
  double sqrtdl = sqrt(ulst.d);
  double sqrtdr = sqrt(urst.d);
  spd[1] = spd[2] - fabs(bxi) / sqrtdl;
  spd[3] = spd[2] + fabs(bxi) / sqrtdr;
  double ptstl = ptl + ul.d * sdl * (spd[2] - wl_ivx);
  double ptstr = ptr + ur.d * sdr * (spd[2] - wr_ivx);
  double ptst = 0.5 * (ptstr + ptstl);
  ulst.mx = ulst.d * spd[2];
  if (fabs(ul.d * sdl * sdml - bxsq) < (HLLD_SMALL_NUMBER)*ptst) {
    ulst.my = ulst.d * wl_ivy;
    ulst.mz = ulst.d * wl_ivz;
    ulst.by = ul.by;
    ulst.bz = ul.bz;
  } else {
    double tmp = bxi * (sdl - sdml) / (ul.d * sdl * sdml - bxsq);
    ulst.my = ulst.d * (wl_ivy - ul.by * tmp);
    ulst.mz = ulst.d * (wl_ivz - ul.bz * tmp);
    tmp = (ul.d * SQR(sdl) - bxsq) / (ul.d * sdl * sdml - bxsq);
    ulst.by = ul.by * tmp;
    ulst.bz = ul.bz * tmp;
  }
  double vbstl = (ulst.mx * bxi + (ulst.my * ulst.by + ulst.mz * ulst.bz)) * ulst_d_inv;
  ulst.e = (sdl * ul.e - ptl * wl_ivx + ptst * spd[2] +
            bxi * (wl_ivx * bxi + (wl_ivy * ul.by + wl_ivz * ul.bz) - vbstl)) * sdml_inv;
  urst.mx = urst.d * spd[2];
  if (fabs(ur.d * sdr * sdmr - bxsq) < (HLLD_SMALL_NUMBER)*ptst) {
    urst.my = urst.d * wr_ivy;
    urst.mz = urst.d * wr_ivz;
    urst.by = ur.by;
    urst.bz = ur.bz;
  } else {
    double tmp = bxi * (sdr - sdmr) / (ur.d * sdr * sdmr - bxsq);
    urst.my = urst.d * (wr_ivy - ur.by * tmp);
    urst.mz = urst.d * (wr_ivz - ur.bz * tmp);
    tmp = (ur.d * SQR(sdr) - bxsq) / (ur.d * sdr * sdmr - bxsq);
    urst.by = ur.by * tmp;
    urst.bz = ur.bz * tmp;
  }
  double vbstr = (urst.mx * bxi + (urst.my * urst.by + urst.mz * urst.bz)) * urst_d_inv;
  urst.e = (sdr * ur.e - ptr * wr_ivx + ptst * spd[2] +
            bxi * (wr_ivx * bxi + (wr_ivy * ur.by + wr_ivz * ur.bz) - vbstr)) * sdmr_inv;
  if (0.5 * bxsq < (HLLD_SMALL_NUMBER)*ptst) {
    uldst = ulst;
    urdst = urst;
  } else {
    double invsumd = 1.0 / (sqrtdl + sqrtdr);
    double bxsig = (bxi > 0.0 ? 1.0 : -1.0);
    uldst.d = ulst.d;
    urdst.d = urst.d;
    uldst.mx = ulst.mx;
    urdst.mx = urst.mx;
    double tmp = invsumd * (sqrtdl * (ulst.my * ulst_d_inv) + sqrtdr * (urst.my * urst_d_inv) +
                            bxsig * (urst.by - ulst.by));
    uldst.my = uldst.d * tmp;
    urdst.my = urdst.d * tmp;
    tmp = invsumd * (sqrtdl * (ulst.mz * ulst_d_inv) + sqrtdr * (urst.mz * urst_d_inv) +
                     bxsig * (urst.bz - ulst.bz));
    uldst.mz = uldst.d * tmp;
    urdst.mz = urdst.d * tmp;
    tmp = invsumd * (sqrtdl * urst.by + sqrtdr * ulst.by +
                     bxsig * sqrtdl * sqrtdr * ((urst.my * urst_d_inv) - (ulst.my * ulst_d_inv)));
    uldst.by = urdst.by = tmp;
    tmp = invsumd * (sqrtdl * urst.bz + sqrtdr * ulst.bz +
                     bxsig * sqrtdl * sqrtdr * ((urst.mz * urst_d_inv) - (ulst.mz * ulst_d_inv)));
    uldst.bz = urdst.bz = tmp;
    tmp = spd[2] * bxi + (uldst.my * uldst.by + uldst.mz * uldst.bz) / uldst.d;
    uldst.e = ulst.e - sqrtdl * bxsig * (vbstl - tmp);
    urdst.e = urst.e + sqrtdr * bxsig * (vbstr - tmp);
  }
  uldst.d = spd[1] * (uldst.d - ulst.d);
  uldst.mx = spd[1] * (uldst.mx - ulst.mx);
  uldst.my = spd[1] * (uldst.my - ulst.my);
  uldst.mz = spd[1] * (uldst.mz - ulst.mz);
  uldst.e = spd[1] * (uldst.e - ulst.e);
  uldst.by = spd[1] * (uldst.by - ulst.by);
  uldst.bz = spd[1] * (uldst.bz - ulst.bz);
  ulst.d = spd[0] * (ulst.d - ul.d);
  ulst.mx = spd[0] * (ulst.mx - ul.mx);
  ulst.my = spd[0] * (ulst.my - ul.my);
  ulst.mz = spd[0] * (ulst.mz - ul.mz);
  ulst.e = spd[0] * (ulst.e - ul.e);
  ulst.by = spd[0] * (ulst.by - ul.by);
  ulst.bz = spd[0] * (ulst.bz - ul.bz);
  urdst.d = spd[3] * (urdst.d - urst.d);
  urdst.mx = spd[3] * (urdst.mx - urst.mx);
  urdst.my = spd[3] * (urdst.my - urst.my);
  urdst.mz = spd[3] * (urdst.mz - urst.mz);
  urdst.e = spd[3] * (urdst.e - urst.e);
  urdst.by = spd[3] * (urdst.by - urst.by);
  urdst.bz = spd[3] * (urdst.bz - urst.bz);
  urst.d = spd[4] * (urst.d - ur.d);
  urst.mx = spd[4] * (urst.mx - ur.mx);
  urst.my = spd[4] * (urst.my - ur.my);
  urst.mz = spd[4] * (urst.mz - ur.mz);
  urst.e = spd[4] * (urst.e - ur.e);
  urst.by = spd[4] * (urst.by - ur.by);
  urst.bz = spd[4] * (urst.bz - ur.bz);
  if (spd[0] >= 0.0) {
    *flxi = fl;
  } else if (spd[4] <= 0.0) {
    *flxi = fr;
  } else if (spd[1] >= 0.0) {
    flxi->d = fl.d + ulst.d; flxi->mx = fl.mx + ulst.mx; flxi->my = fl.my + ulst.my;
    flxi->mz = fl.mz + ulst.mz; flxi->e = fl.e + ulst.e; flxi->by = fl.by + ulst.by;
    flxi->bz = fl.bz + ulst.bz;
  } else if (spd[2] >= 0.0) {
    flxi->d = fl.d + ulst.d + uldst.d; flxi->mx = fl.mx + ulst.mx + uldst.mx;
    flxi->my = fl.my + ulst.my + uldst.my; flxi->mz = fl.mz + ulst.mz + uldst.mz;
    flxi->e = fl.e + ulst.e + uldst.e; flxi->by = fl.by + ulst.by + uldst.by;
    flxi->bz = fl.bz + ulst.bz + uldst.bz;
  } else if (spd[3] > 0.0) {
    flxi->d = fr.d + urst.d + urdst.d; flxi->mx = fr.mx + urst.mx + urdst.mx;
    flxi->my = fr.my + urst.my + urdst.my; flxi->mz = fr.mz + urst.mz + urdst.mz;
    flxi->e = fr.e + urst.e + urdst.e; flxi->by = fr.by + urst.by + urdst.by;
    flxi->bz = fr.bz + urst.bz + urdst.bz;
  } else {
    flxi->d = fr.d + urst.d; flxi->mx = fr.mx + urst.mx; flxi->my = fr.my + urst.my;
    flxi->mz = fr.mz + urst.mz; flxi->e = fr.e + urst.e; flxi->by = fr.by + urst.by;
    flxi->bz = fr.bz + urst.bz;
  }
}

/* eos.hpp:60-68: isothermal fast speed */
static double fast_speed_iso(double iso_cs, double d, double bx, double by, double bz) {
  double asq = (iso_cs * iso_cs) * d;
  double ct2 = by * by + bz * bz;
  double qsq = bx * bx + ct2 + asq;
  double tmp = bx * bx + ct2 - asq;
  return sqrt(0.5 * (qsq + sqrt(tmp * tmp + 4.0 * asq * ct2)) / d);
}

/* llf_mhd_singlestate.hpp:27-89, isothermal branches */
static void llf_face_iso(const ocfg *c, const prim1d *wl, const prim1d *wr, double bxi, cons1d *flux) {
  double qa = wl->d * wl->vx;
  double qb = wr->d * wr->vx;
  double qc = 0.5 * (SQR(wl->by) + SQR(wl->bz) - SQR(bxi));
  double qd = 0.5 * (SQR(wr->by) + SQR(wr->bz) - SQR(bxi));
  cons1d fsum, du;
  fsum.d = qa + qb;
  fsum.mx = qa * wl->vx + qb * wr->vx + qc + qd;
  fsum.my = qa * wl->vy + qb * wr->vy - bxi * (wl->by + wr->by);
  fsum.mz = qa * wl->vz + qb * wr->vz - bxi * (wl->bz + wr->bz);
  fsum.by = wl->by * wl->vx + wr->by * wr->vx - bxi * (wl->vy + wr->vy);
  fsum.bz = wl->bz * wl->vx + wr->bz * wr->vx - bxi * (wl->vz + wr->vz);
  fsum.mx += SQR(c->iso_cs) * (wl->d + wr->d);
  qa = fast_speed_iso(c->iso_cs, wl->d, bxi, wl->by, wl->bz);
  qb = fast_speed_iso(c->iso_cs, wr->d, bxi, wr->by, wr->bz);
  double a = fmax((fabs(wl->vx) + qa), (fabs(wr->vx) + qb));
  du.d = a * (wr->d - wl->d);
  du.mx = a * (wr->d * wr->vx - wl->d * wl->vx);
  du.my = a * (wr->d * wr->vy - wl->d * wl->vy);
  du.mz = a * (wr->d * wr->vz - wl->d * wl->vz);
  du.by = a * (wr->by - wl->by);
  du.bz = a * (wr->bz - wl->bz);
  flux->d = 0.5 * (fsum.d - du.d);
  flux->mx = 0.5 * (fsum.mx - du.mx);
  flux->my = 0.5 * (fsum.my - du.my);
  flux->mz = 0.5 * (fsum.mz - du.mz);
  flux->e = 0.0;
  flux->by = -0.5 * (fsum.by - du.by);
  flux->bz = 0.5 * (fsum.bz - du.bz);
}

/* hlle_mhd.hpp:24-179, isothermal branches; ey/ez as stored by the reference */
static void hlle_face_iso(const ocfg *c, const prim1d *wl, const prim1d *wr, double bxi, cons1d *flux, double *ey, double *ez) {
  const double iso_cs = c->iso_cs;
  double sqrtdl = sqrt(wl->d), sqrtdr = sqrt(wr->d);
  double isdlpdr = 1.0 / (sqrtdl + sqrtdr);
  double wroe_idn = sqrtdl * sqrtdr;
  double wroe_ivx = (sqrtdl * wl->vx + sqrtdr * wr->vx) * isdlpdr;
  double wroe_iby = (sqrtdr * wl->by + sqrtdl * wr->by) * isdlpdr;
  double wroe_ibz = (sqrtdr * wl->bz + sqrtdl * wr->bz) * isdlpdr;
  double x = 0.5 * (SQR(wl->by - wr->by) + SQR(wl->bz - wr->bz)) / (SQR(sqrtdl + sqrtdr));
  double y = 0.5 * (wl->d + wr->d) / wroe_idn;
  double pbl = 0.5 * (bxi * bxi + SQR(wl->by) + SQR(wl->bz));
  double pbr = 0.5 * (bxi * bxi + SQR(wr->by) + SQR(wr->bz));
  double cl = fast_speed_iso(iso_cs, wl->d, bxi, wl->by, wl->bz);
  double cr = fast_speed_iso(iso_cs, wr->d, bxi, wr->by, wr->bz);
  double btsq = SQR(wroe_iby) + SQR(wroe_ibz);
  double vaxsq = bxi * bxi / wroe_idn;
  double bt_starsq = btsq * y;
  double twid_asq = iso_cs * iso_cs + x;
  double ct2 = bt_starsq / wroe_idn;
  double tsum = vaxsq + ct2 + twid_asq;
  double tdif = vaxsq + ct2 - twid_asq;
  double cf2_cs2 = sqrt(tdif * tdif + 4.0 * twid_asq * ct2);
  double cfsq = 0.5 * (tsum + cf2_cs2);
  double a = sqrt(cfsq);
  double al = fmin((wroe_ivx - a), (wl->vx - cl));
  double ar = fmax((wroe_ivx + a), (wr->vx + cr));
  double bp = ar > 0.0 ? ar : 1.0e-20;
  double bm = al < 0.0 ? al : -1.0e-20;
  double vxl = wl->vx - bm, vxr = wr->vx - bp;
  cons1d fl, fr;
  fl.d = wl->d * vxl;                                  fr.d = wr->d * vxr;
  fl.mx = wl->d * wl->vx * vxl + pbl - SQR(bxi);       fr.mx = wr->d * wr->vx * vxr + pbr - SQR(bxi);
  fl.my = wl->d * wl->vy * vxl - bxi * wl->by;         fr.my = wr->d * wr->vy * vxr - bxi * wr->by;
  fl.mz = wl->d * wl->vz * vxl - bxi * wl->bz;         fr.mz = wr->d * wr->vz * vxr - bxi * wr->bz;
  fl.mx += (iso_cs * iso_cs) * wl->d;                  fr.mx += (iso_cs * iso_cs) * wr->d;
  fl.by = wl->by * vxl - bxi * wl->vy;                 fr.by = wr->by * vxr - bxi * wr->vy;
  fl.bz = wl->bz * vxl - bxi * wl->vz;                 fr.bz = wr->bz * vxr - bxi * wr->vz;
  double tmp = 0.0;
  if (bp != bm) tmp = 0.5 * (bp + bm) / (bp - bm);
  flux->d = 0.5 * (fl.d + fr.d) + (fl.d - fr.d) * tmp;
  flux->mx = 0.5 * (fl.mx + fr.mx) + (fl.mx - fr.mx) * tmp;
  flux->my = 0.5 * (fl.my + fr.my) + (fl.my - fr.my) * tmp;
  flux->mz = 0.5 * (fl.mz + fr.mz) + (fl.mz - fr.mz) * tmp;
  flux->e = 0.0;
  *ey = -0.5 * (fl.by + fr.by) - (fl.by - fr.by) * tmp;
  *ez = 0.5 * (fl.bz + fr.bz) + (fl.bz - fr.bz) * tmp;
}

/* hlld_mhd.hpp:350-546: isothermal HLLD (Mignone 2007) */
static void hlld_face_iso(const ocfg *c, const prim1d *wl, const prim1d *wr, double bxi, cons1d *flxi) {
  const double dfloor_ = c->dfloor, iso_cs = c->iso_cs;
  double spd[5];
  cons1d ul, ur;
  ul.d = wl->d; ul.mx = wl->vx * ul.d; ul.my = wl->vy * ul.d; ul.mz = wl->vz * ul.d; ul.by = wl->by; ul.bz = wl->bz;
  ur.d = wr->d; ur.mx = wr->vx * ur.d; ur.my = wr->vy * ur.d; ur.mz = wr->vz * ur.d; ur.by = wr->by; ur.bz = wr->bz;
  double cfl = fast_speed_iso(iso_cs, wl->d, bxi, wl->by, wl->bz);
  double cfr = fast_speed_iso(iso_cs, wr->d, bxi, wr->by, wr->bz);
  spd[0] = fmin(wl->vx - cfl, wr->vx - cfr);
  spd[4] = fmax(wl->vx + cfl, wr->vx + cfr);
  double bxsq = bxi * bxi;
  double ptl = SQR(iso_cs) * wl->d + 0.5 * (bxsq + SQR(wl->by) + SQR(wl->bz));
  double ptr = SQR(iso_cs) * wr->d + 0.5 * (bxsq + SQR(wr->by) + SQR(wr->bz));
  cons1d fl, fr;
  fl.d = ul.mx;
  fl.mx = ul.mx * wl->vx + ptl - bxsq;
  fl.my = ul.my * wl->vx - bxi * ul.by;
  fl.mz = ul.mz * wl->vx - bxi * ul.bz;
  fl.by = ul.by * wl->vx - bxi * wl->vy;
  fl.bz = ul.bz * wl->vx - bxi * wl->vz;
  fr.d = ur.mx;
  fr.mx = ur.mx * wr->vx + ptr - bxsq;
  fr.my = ur.my * wr->vx - bxi * ur.by;
  fr.mz = ur.mz * wr->vx - bxi * ur.bz;
  fr.by = ur.by * wr->vx - bxi * wr->vy;
  fr.bz = ur.bz * wr->vx - bxi * wr->vz;
  double idspd = 1.0 / (spd[4] - spd[0]);
  double dhll = (spd[4] * ur.d - spd[0] * ul.d - fr.d + fl.d) * idspd;
  dhll = fmax(dhll, dfloor_);
  double sqrtdhll = sqrt(dhll);
  double fdhll = (spd[4] * fl.d - spd[0] * fr.d + spd[4] * spd[0] * (ur.d - ul.d)) * idspd;
  double fmxhll = (spd[4] * fl.mx - spd[0] * fr.mx + spd[4] * spd[0] * (ur.mx - ul.mx)) * idspd;
  double ustar = fdhll / dhll;
  double mxhll = (spd[4] * ur.mx - spd[0] * ul.mx - fr.mx + fl.mx) * idspd;
  spd[1] = ustar - fabs(bxi) / sqrtdhll;
  spd[3] = ustar + fabs(bxi) / sqrtdhll;
  cons1d ulst, urst, ucst;
  ulst.d = dhll; ulst.mx = mxhll;
  double tmp = (spd[0] - spd[1]) * (spd[0] - spd[3]);
  if (fabs(spd[0] - spd[1]) < (HLLD_SMALL_NUMBER)*iso_cs) {
    ulst.my = ul.my; ulst.mz = ul.mz; ulst.by = ul.by; ulst.bz = ul.bz;
  } else {
    double mfact = bxi * (ustar - wl->vx) / tmp;
    double bfact = (ul.d * SQR(spd[0] - wl->vx) - bxsq) / (dhll * tmp);
    ulst.my = dhll * wl->vy - ul.by * mfact;
    ulst.mz = dhll * wl->vz - ul.bz * mfact;
    ulst.by = ul.by * bfact;
    ulst.bz = ul.bz * bfact;
  }
  urst.d = dhll; urst.mx = mxhll;
  tmp = (spd[4] - spd[1]) * (spd[4] - spd[3]);
  if (fabs(spd[4] - spd[3]) < (HLLD_SMALL_NUMBER)*iso_cs) {
    urst.my = ur.my; urst.mz = ur.mz; urst.by = ur.by; urst.bz = ur.bz;
  } else {
    double mfact = bxi * (ustar - wr->vx) / tmp;
    double bfact = (ur.d * SQR(spd[4] - wr->vx) - bxsq) / (dhll * tmp);
    urst.my = dhll * wr->vy - ur.by * mfact;
    urst.mz = dhll * wr->vz - ur.bz * mfact;
    urst.by = ur.by * bfact;
    urst.bz = ur.bz * bfact;
  }
  double x = sqrtdhll * (bxi > 0.0 ? 1.0 : -1.0);
  ucst.d = dhll; ucst.mx = mxhll;
  ucst.my = 0.5 * (ulst.my + urst.my + (urst.by - ulst.by) * x);
  ucst.mz = 0.5 * (ulst.mz + urst.mz + (urst.bz - ulst.bz) * x);
  ucst.by = 0.5 * (ulst.by + urst.by + (urst.my - ulst.my) / x);
  ucst.bz = 0.5 * (ulst.bz + urst.bz + (urst.mz - ulst.mz) / x);
  if (spd[0] >= 0.0) {
    *flxi = fl;
  } else if (spd[4] <= 0.0) {
    *flxi = fr;
  } else if (spd[1] >= 0.0) {
    flxi->d = fl.d + spd[0] * (ulst.d - ul.d);
    flxi->mx = fl.mx + spd[0] * (ulst.mx - ul.mx);
    flxi->my = fl.my + spd[0] * (ulst.my - ul.my);
    flxi->mz = fl.mz + spd[0] * (ulst.mz - ul.mz);
    flxi->by = fl.by + spd[0] * (ulst.by - ul.by);
    flxi->bz = fl.bz + spd[0] * (ulst.bz - ul.bz);
  } else if (spd[3] <= 0.0) {
    flxi->d = fr.d + spd[4] * (urst.d - ur.d);
    flxi->mx = fr.mx + spd[4] * (urst.mx - ur.mx);
    flxi->my = fr.my + spd[4] * (urst.my - ur.my);
    flxi->mz = fr.mz + spd[4] * (urst.mz - ur.mz);
    flxi->by = fr.by + spd[4] * (urst.by - ur.by);
    flxi->bz = fr.bz + spd[4] * (urst.bz - ur.bz);
  } else {
    flxi->d = dhll * ustar;
    flxi->mx = fmxhll;
    flxi->my = ucst.my * ustar - bxi * ucst.by;
    flxi->mz = ucst.mz * ustar - bxi * ucst.bz;
    flxi->by = ucst.by * ustar - bxi * ucst.my / ucst.d;
    flxi->bz = ucst.bz * ustar - bxi * ucst.mz / ucst.d;
  }
  flxi->e = 0.0;
}


/* mhd_fluxes.cpp:47-77 (SolveFaceMHD) for one face; writes 5 fluxes into f[0..4] in the
 * UNROTATED variable order and the two EMFs (ey,ez). */
static void solve_face(const ocfg *c, int dir, const double *wl5, const double *wr5,
                       const double *bl3, const double *br3, double bxi, double *f5, double *ey,
                       double *ez) {
  const int ivx = IVX + dir, ivy = IVX + (dir + 1) % 3, ivz = IVX + (dir + 2) % 3;
  const int iby = (dir + 1) % 3, ibz = (dir + 2) % 3;
  prim1d l, r;
  l.d = wl5[IDN]; l.vx = wl5[ivx]; l.vy = wl5[ivy]; l.vz = wl5[ivz]; l.e = c->is_iso ? 0.0 : wl5[IEN];
  l.by = bl3[iby]; l.bz = bl3[ibz];
  r.d = wr5[IDN]; r.vx = wr5[ivx]; r.vy = wr5[ivy]; r.vz = wr5[ivz]; r.e = c->is_iso ? 0.0 : wr5[IEN];
  r.by = br3[iby]; r.bz = br3[ibz];
  cons1d fx;
  if (c->is_iso) {
    if (c->rsolver == 2) { hlld_face_iso(c, &l, &r, bxi, &fx); *ey = -fx.by; *ez = fx.bz; }
    else if (c->rsolver == 1) { hlle_face_iso(c, &l, &r, bxi, &fx, ey, ez); }
    else { llf_face_iso(c, &l, &r, bxi, &fx); *ey = fx.by; *ez = fx.bz; }
    f5[IDN] = fx.d; f5[ivx] = fx.mx; f5[ivy] = fx.my; f5[ivz] = fx.mz;
    return;
  }
  if (c->rsolver == 2) {
    hlld_face(c, &l, &r, bxi, &fx);
    *ey = -fx.by; *ez = fx.bz;                     /* hlld_mhd.hpp:346-347 */
  } else if (c->rsolver == 1) {
    hlle_face(c, &l, &r, bxi, &fx, ey, ez);        /* hlle_mhd.hpp:176-177 */
  } else {
    llf_face(c, &l, &r, bxi, &fx);
    *ey = fx.by; *ez = fx.bz;                      /* llf_mhd.hpp:71-72 */
  }
  f5[IDN] = fx.d; f5[ivx] = fx.mx; f5[ivy] = fx.my; f5[ivz] = fx.mz; f5[IEN] = fx.e;
}

/* MHD::CalculateFluxes, mhd_fluxes.cpp:85-266 (3-D, no FOFC, no scalars).
 * scratch: wl,wr (5 cell arrays each), bl,br (3 each). */
void oracle_fluxes(const ocfg *c, const double *w0, const double *bcc0, const double *b1f,
                   const double *b2f, const double *b3f, double *flx1, double *flx2, double *flx3,
                   double *e3x1, double *e2x1, double *e1x2, double *e3x2, double *e2x3,
                   double *e1x3) {
  odims D = dims(c);
  size_t ncc = (size_t)c->nmb * D.cc;
  double *wl = calloc(D.nv * ncc, 8), *wr = calloc(D.nv * ncc, 8);
  double *bl = calloc(3 * ncc, 8), *br = calloc(3 * ncc, 8);
  const int is = D.is, ie = D.ie, js = D.js, je = D.je, ks = D.ks, ke = D.ke;
  const int multi_d = c->nx2 > 1, three_d = c->nx3 > 1;
  const int ndir = 1 + multi_d + three_d;
  for (int dir = 0; dir < ndir; ++dir) {
    /* CT-extended transverse ranges only in ACTIVE dimensions (mhd_fluxes.cpp:126-128,176-178,225-226) */
    int kl = ks - three_d, ku = ke + three_d, jl = js - multi_d, ju = je + multi_d, il = is - 1, iu = ie + 1;   /* recon cell box */
    int fkl = kl, fku = ku, fjl = jl, fju = ju, fil = il, fiu = iu;                                               /* face box */
    const int nx = c->fofc ? 1 : 0;   /* mhd_fluxes.cpp:100-106: face-normal range one cell wider with FOFC */
    if (dir == 0) { fil = is - nx; fiu = ie + 1 + nx; il = fil - 1; iu = fiu; }
    else if (dir == 1) { fjl = js - nx; fju = je + 1 + nx; jl = fjl - 1; ju = fju; }
    else { fkl = ks - nx; fku = ke + 1 + nx; kl = fkl - 1; ku = fku; }
    recon_pass(c, dir, D.nv, 1, w0, wl, wr, kl, ku, jl, ju, il, iu);
    recon_pass(c, dir, 3, 0, bcc0, bl, br, kl, ku, jl, ju, il, iu);
    for (int m = 0; m < c->nmb; ++m)
      for (int k = fkl; k <= fku; ++k)
        for (int j = fjl; j <= fju; ++j)
          for (int i = fil; i <= fiu; ++i) {
            double wl5[5], wr5[5], bl3[3], br3[3], f5[5], ey, ez, bxi;
            for (int n = 0; n < D.nfl; ++n) { wl5[n] = CC(wl, D.nv, m, n, k, j, i); wr5[n] = CC(wr, D.nv, m, n, k, j, i); }
            for (int n = 0; n < 3; ++n) { bl3[n] = CC(bl, 3, m, n, k, j, i); br3[n] = CC(br, 3, m, n, k, j, i); }
            if (dir == 0) bxi = X1F(b1f, m, k, j, i);
            else if (dir == 1) bxi = X2F(b2f, m, k, j, i);
            else bxi = X3F(b3f, m, k, j, i);
            solve_face(c, dir, wl5, wr5, bl3, br3, bxi, f5, &ey, &ez);
            if (dir == 0) {
              for (int n = 0; n < D.nfl; ++n) F1(flx1, m, n, k, j, i) = f5[n];
              S4(e3x1, m, k, j, i) = ey; S4(e2x1, m, k, j, i) = ez;   /* mhd_fluxes.cpp:146 */
            } else if (dir == 1) {
              for (int n = 0; n < D.nfl; ++n) F2(flx2, m, n, k, j, i) = f5[n];
              S4(e1x2, m, k, j, i) = ey; S4(e3x2, m, k, j, i) = ez;   /* :196 */
            } else {
              for (int n = 0; n < D.nfl; ++n) F3(flx3, m, n, k, j, i) = f5[n];
              S4(e2x3, m, k, j, i) = ey; S4(e1x3, m, k, j, i) = ez;   /* :243 */
            }
          }
    /* scalar fluxes, upwinded by the sign of the mass flux, ACTIVE transverse range only
     * (mhd_fluxes.cpp:153-166,203-216,250-263) */
    if (D.nv > D.nfl)
      for (int m = 0; m < c->nmb; ++m)
        for (int k = ks; k <= ke + (dir == 2); ++k)
          for (int j = js; j <= je + (dir == 1); ++j)
            for (int i = is; i <= ie + (dir == 0); ++i)
              for (int n = D.nfl; n < D.nv; ++n) {
                if (dir == 0) F1(flx1, m, n, k, j, i) = F1(flx1, m, IDN, k, j, i) * ((F1(flx1, m, IDN, k, j, i) >= 0.0) ? CC(wl, D.nv, m, n, k, j, i) : CC(wr, D.nv, m, n, k, j, i));
                else if (dir == 1) F2(flx2, m, n, k, j, i) = F2(flx2, m, IDN, k, j, i) * ((F2(flx2, m, IDN, k, j, i) >= 0.0) ? CC(wl, D.nv, m, n, k, j, i) : CC(wr, D.nv, m, n, k, j, i));
                else F3(flx3, m, n, k, j, i) = F3(flx3, m, IDN, k, j, i) * ((F3(flx3, m, IDN, k, j, i) >= 0.0) ? CC(wl, D.nv, m, n, k, j, i) : CC(wr, D.nv, m, n, k, j, i));
              }
  }
  free(wl); free(wr); free(bl); free(br);
}

/* MHD::RKUpdate, mhd_update.cpp:24-84.  mb_size: [nmb][9] = x?min(3),x?max(3),dx?(3). */
void oracle_rk_update(const ocfg *c, double *u0, const double *u1, const double *flx1,
                      const double *flx2, const double *flx3, const double *mb_size, double gam0,
                      double gam1, double beta_dt) {
  odims D = dims(c);
  for (int m = 0; m < c->nmb; ++m) {
    const double dx1 = mb_size[m * 9 + 6], dx2 = mb_size[m * 9 + 7], dx3 = mb_size[m * 9 + 8];
    for (int n = 0; n < D.nv; ++n)
      for (int k = D.ks; k <= D.ke; ++k)
        for (int j = D.js; j <= D.je; ++j)
          for (int i = D.is; i <= D.ie; ++i) {
            double divf = (F1(flx1, m, n, k, j, i + 1) - F1(flx1, m, n, k, j, i)) / dx1;
            if (c->nx2 > 1) divf += (F2(flx2, m, n, k, j + 1, i) - F2(flx2, m, n, k, j, i)) / dx2;   /* mhd_update.cpp:63-69 multi_d */
            if (c->nx3 > 1) divf += (F3(flx3, m, n, k + 1, j, i) - F3(flx3, m, n, k, j, i)) / dx3;   /* :72-78 three_d */
            CC(u0, D.nv, m, n, k, j, i) = gam0 * CC(u0, D.nv, m, n, k, j, i) + gam1 * CC(u1, D.nv, m, n, k, j, i) - beta_dt * divf;
          }
  }
}

/* MHD::CornerE 3-D Newtonian branch, mhd_corner_e.cpp:206-215,309-318,338-414 */
void oracle_corner_e(const ocfg *c, const double *w0, const double *bcc, const double *flx1,
                     const double *flx2, const double *flx3, const double *e3x1, const double *e2x1,
                     const double *e1x2, const double *e3x2, const double *e2x3, const double *e1x3,
                     double *e1, double *e2, double *e3) {
  odims D = dims(c);
  size_t ncc = (size_t)c->nmb * D.cc;
  const int is = D.is, ie = D.ie, js = D.js, je = D.je, ks = D.ks, ke = D.ke;
  if (c->nx2 == 1) {   /* 1-D: copy face-centred E to the edges, mhd_corner_e.cpp:41-54 */
    for (int m = 0; m < c->nmb; ++m)
      for (int i = is; i <= ie + 1; ++i) {
        X2E(e2, m, ks, js, i) = S4(e2x1, m, ks, js, i);
        X2E(e2, m, ke + 1, js, i) = S4(e2x1, m, ks, js, i);
        X3E(e3, m, ks, js, i) = S4(e3x1, m, ks, js, i);
        X3E(e3, m, ks, je + 1, i) = S4(e3x1, m, ks, js, i);
      }
    return;
  }
  if (c->nx3 == 1) {   /* 2-D: E1, E2 copied, E3 by GS07, mhd_corner_e.cpp:59-201 */
    double *e3c = calloc(ncc, 8);
    for (int m = 0; m < c->nmb; ++m)
      for (int j = js - 1; j <= je + 1; ++j)
        for (int i = is - 1; i <= ie + 1; ++i)
          S4(e3c, m, ks, j, i) = CC(w0, D.nv, m, IVY, ks, j, i) * CC(bcc, 3, m, 0, ks, j, i) - CC(w0, D.nv, m, IVX, ks, j, i) * CC(bcc, 3, m, 1, ks, j, i);
    for (int m = 0; m < c->nmb; ++m)
      for (int j = js; j <= je + 1; ++j)
        for (int i = is; i <= ie + 1; ++i) {
          X2E(e2, m, ks, j, i) = S4(e2x1, m, ks, j, i);
          X2E(e2, m, ke + 1, j, i) = S4(e2x1, m, ks, j, i);
          X1E(e1, m, ks, j, i) = S4(e1x2, m, ks, j, i);
          X1E(e1, m, ke + 1, j, i) = S4(e1x2, m, ks, j, i);
          double e3_l2, e3_r2, e3_l1, e3_r1;
          if (F1(flx1, m, IDN, ks, j - 1, i) >= 0.0) e3_l2 = S4(e3x2, m, ks, j, i - 1) - S4(e3c, m, ks, j - 1, i - 1);
          else e3_l2 = S4(e3x2, m, ks, j, i) - S4(e3c, m, ks, j - 1, i);
          if (F1(flx1, m, IDN, ks, j, i) >= 0.0) e3_r2 = S4(e3x2, m, ks, j, i - 1) - S4(e3c, m, ks, j, i - 1);
          else e3_r2 = S4(e3x2, m, ks, j, i) - S4(e3c, m, ks, j, i);
          if (F2(flx2, m, IDN, ks, j, i - 1) >= 0.0) e3_l1 = S4(e3x1, m, ks, j - 1, i) - S4(e3c, m, ks, j - 1, i - 1);
          else e3_l1 = S4(e3x1, m, ks, j, i) - S4(e3c, m, ks, j, i - 1);
          if (F2(flx2, m, IDN, ks, j, i) >= 0.0) e3_r1 = S4(e3x1, m, ks, j - 1, i) - S4(e3c, m, ks, j - 1, i);
          else e3_r1 = S4(e3x1, m, ks, j, i) - S4(e3c, m, ks, j, i);
          X3E(e3, m, ks, j, i) = 0.25 * (e3_l1 + e3_r1 + e3_l2 + e3_r2 + S4(e3x2, m, ks, j, i - 1) + S4(e3x2, m, ks, j, i) + S4(e3x1, m, ks, j - 1, i) + S4(e3x1, m, ks, j, i));
        }
    free(e3c);
    return;
  }
  double *e1cc = calloc(ncc, 8), *e2cc = calloc(ncc, 8), *e3cc = calloc(ncc, 8);
  for (int m = 0; m < c->nmb; ++m)
    for (int k = ks - 1; k <= ke + 1; ++k)
      for (int j = js - 1; j <= je + 1; ++j)
        for (int i = is - 1; i <= ie + 1; ++i) {
          S4(e1cc, m, k, j, i) = CC(w0, D.nv, m, IVZ, k, j, i) * CC(bcc, 3, m, 1, k, j, i) - CC(w0, D.nv, m, IVY, k, j, i) * CC(bcc, 3, m, 2, k, j, i);
          S4(e2cc, m, k, j, i) = CC(w0, D.nv, m, IVX, k, j, i) * CC(bcc, 3, m, 2, k, j, i) - CC(w0, D.nv, m, IVZ, k, j, i) * CC(bcc, 3, m, 0, k, j, i);
          S4(e3cc, m, k, j, i) = CC(w0, D.nv, m, IVY, k, j, i) * CC(bcc, 3, m, 0, k, j, i) - CC(w0, D.nv, m, IVX, k, j, i) * CC(bcc, 3, m, 1, k, j, i);
        }
  for (int m = 0; m < c->nmb; ++m)
    for (int k = ks; k <= ke + 1; ++k)
      for (int j = js; j <= je + 1; ++j)
        for (int i = is; i <= ie + 1; ++i) {
          double e1_l3, e1_r3, e1_l2, e1_r2;
          if (F2(flx2, m, IDN, k - 1, j, i) >= 0.0) e1_l3 = S4(e1x3, m, k, j - 1, i) - S4(e1cc, m, k - 1, j - 1, i);
          else e1_l3 = S4(e1x3, m, k, j, i) - S4(e1cc, m, k - 1, j, i);
          if (F2(flx2, m, IDN, k, j, i) >= 0.0) e1_r3 = S4(e1x3, m, k, j - 1, i) - S4(e1cc, m, k, j - 1, i);
          else e1_r3 = S4(e1x3, m, k, j, i) - S4(e1cc, m, k, j, i);
          if (F3(flx3, m, IDN, k, j - 1, i) >= 0.0) e1_l2 = S4(e1x2, m, k - 1, j, i) - S4(e1cc, m, k - 1, j - 1, i);
          else e1_l2 = S4(e1x2, m, k, j, i) - S4(e1cc, m, k, j - 1, i);
          if (F3(flx3, m, IDN, k, j, i) >= 0.0) e1_r2 = S4(e1x2, m, k - 1, j, i) - S4(e1cc, m, k - 1, j, i);
          else e1_r2 = S4(e1x2, m, k, j, i) - S4(e1cc, m, k, j, i);
          X1E(e1, m, k, j, i) = 0.25 * (e1_l3 + e1_r3 + e1_l2 + e1_r2 + S4(e1x2, m, k - 1, j, i) + S4(e1x2, m, k, j, i) + S4(e1x3, m, k, j - 1, i) + S4(e1x3, m, k, j, i));

          double e2_l3, e2_r3, e2_l1, e2_r1;
          if (F1(flx1, m, IDN, k - 1, j, i) >= 0.0) e2_l3 = S4(e2x3, m, k, j, i - 1) - S4(e2cc, m, k - 1, j, i - 1);
          else e2_l3 = S4(e2x3, m, k, j, i) - S4(e2cc, m, k - 1, j, i);
          if (F1(flx1, m, IDN, k, j, i) >= 0.0) e2_r3 = S4(e2x3, m, k, j, i - 1) - S4(e2cc, m, k, j, i - 1);
          else e2_r3 = S4(e2x3, m, k, j, i) - S4(e2cc, m, k, j, i);
          if (F3(flx3, m, IDN, k, j, i - 1) >= 0.0) e2_l1 = S4(e2x1, m, k - 1, j, i) - S4(e2cc, m, k - 1, j, i - 1);
          else e2_l1 = S4(e2x1, m, k, j, i) - S4(e2cc, m, k, j, i - 1);
          if (F3(flx3, m, IDN, k, j, i) >= 0.0) e2_r1 = S4(e2x1, m, k - 1, j, i) - S4(e2cc, m, k - 1, j, i);
          else e2_r1 = S4(e2x1, m, k, j, i) - S4(e2cc, m, k, j, i);
          X2E(e2, m, k, j, i) = 0.25 * (e2_l3 + e2_r3 + e2_l1 + e2_r1 + S4(e2x3, m, k, j, i - 1) + S4(e2x3, m, k, j, i) + S4(e2x1, m, k - 1, j, i) + S4(e2x1, m, k, j, i));

          double e3_l2, e3_r2, e3_l1, e3_r1;
          if (F1(flx1, m, IDN, k, j - 1, i) >= 0.0) e3_l2 = S4(e3x2, m, k, j, i - 1) - S4(e3cc, m, k, j - 1, i - 1);
          else e3_l2 = S4(e3x2, m, k, j, i) - S4(e3cc, m, k, j - 1, i);
          if (F1(flx1, m, IDN, k, j, i) >= 0.0) e3_r2 = S4(e3x2, m, k, j, i - 1) - S4(e3cc, m, k, j, i - 1);
          else e3_r2 = S4(e3x2, m, k, j, i) - S4(e3cc, m, k, j, i);
          if (F2(flx2, m, IDN, k, j, i - 1) >= 0.0) e3_l1 = S4(e3x1, m, k, j - 1, i) - S4(e3cc, m, k, j - 1, i - 1);
          else e3_l1 = S4(e3x1, m, k, j, i) - S4(e3cc, m, k, j, i - 1);
          if (F2(flx2, m, IDN, k, j, i) >= 0.0) e3_r1 = S4(e3x1, m, k, j - 1, i) - S4(e3cc, m, k, j - 1, i);
          else e3_r1 = S4(e3x1, m, k, j, i) - S4(e3cc, m, k, j, i);
          X3E(e3, m, k, j, i) = 0.25 * (e3_l1 + e3_r1 + e3_l2 + e3_r2 + S4(e3x2, m, k, j, i - 1) + S4(e3x2, m, k, j, i) + S4(e3x1, m, k, j - 1, i) + S4(e3x1, m, k, j, i));
        }
  free(e1cc); free(e2cc); free(e3cc);
}

/* MHD::CT, mhd_ct.cpp:23-80 (3-D) */
void oracle_ct(const ocfg *c, double *b1f, double *b2f, double *b3f, const double *o1f,
               const double *o2f, const double *o3f, const double *e1, const double *e2,
               const double *e3, const double *mb_size, double gam0, double gam1, double beta_dt) {
  odims D = dims(c);
  for (int m = 0; m < c->nmb; ++m) {
    const double dx1 = mb_size[m * 9 + 6], dx2 = mb_size[m * 9 + 7], dx3 = mb_size[m * 9 + 8];
    const int multi_d = c->nx2 > 1, three_d = c->nx3 > 1;
    if (multi_d)   /* B1 is not touched at all in 1-D (mhd_ct.cpp:41-53) */
    for (int k = D.ks; k <= D.ke; ++k)
      for (int j = D.js; j <= D.je; ++j)
        for (int i = D.is; i <= D.ie + 1; ++i) {
          double b = gam0 * X1F(b1f, m, k, j, i) + gam1 * X1F(o1f, m, k, j, i);
          b -= beta_dt * (X3E(e3, m, k, j + 1, i) - X3E(e3, m, k, j, i)) / dx2;
          if (three_d) b += beta_dt * (X2E(e2, m, k + 1, j, i) - X2E(e2, m, k, j, i)) / dx3;
          X1F(b1f, m, k, j, i) = b;
        }
    for (int k = D.ks; k <= D.ke; ++k)
      for (int j = D.js; j <= D.je + 1; ++j)
        for (int i = D.is; i <= D.ie; ++i) {
          double b = gam0 * X2F(b2f, m, k, j, i) + gam1 * X2F(o2f, m, k, j, i);
          b += beta_dt * (X3E(e3, m, k, j, i + 1) - X3E(e3, m, k, j, i)) / dx1;
          if (three_d) b -= beta_dt * (X1E(e1, m, k + 1, j, i) - X1E(e1, m, k, j, i)) / dx3;
          X2F(b2f, m, k, j, i) = b;
        }
    for (int k = D.ks; k <= D.ke + 1; ++k)
      for (int j = D.js; j <= D.je; ++j)
        for (int i = D.is; i <= D.ie; ++i) {
          double b = gam0 * X3F(b3f, m, k, j, i) + gam1 * X3F(o3f, m, k, j, i);
          b -= beta_dt * (X2E(e2, m, k, j, i + 1) - X2E(e2, m, k, j, i)) / dx1;
          if (multi_d) b += beta_dt * (X1E(e1, m, k, j + 1, i) - X1E(e1, m, k, j, i)) / dx2;
          X3F(b3f, m, k, j, i) = b;
        }
  }
}

/* IdealMHD::ConsToPrim + SingleC2P_IdealMHD, ideal_mhd.cpp:30-134, ideal_c2p_mhd.hpp:20-67 */
void oracle_cons2prim(const ocfg *c, double *u0, const double *b1f, const double *b2f,
                      const double *b3f, double *w0, double *bcc, int il, int iu, int jl, int ju,
                      int kl, int ku, int *nfloor) {
  odims D = dims(c);
  int nd = 0, ne = 0, nt = 0;
  const double gm1 = c->gamma - 1.0;
  for (int m = 0; m < c->nmb; ++m)
    for (int k = kl; k <= ku; ++k)
      for (int j = jl; j <= ju; ++j)
        for (int i = il; i <= iu; ++i) {
          double ud = CC(u0, D.nv, m, IDN, k, j, i), umx = CC(u0, D.nv, m, IM1, k, j, i);
          double umy = CC(u0, D.nv, m, IM2, k, j, i), umz = CC(u0, D.nv, m, IM3, k, j, i);
          double ubx = 0.5 * (X1F(b1f, m, k, j, i) + X1F(b1f, m, k, j, i + 1));
          double uby = 0.5 * (X2F(b2f, m, k, j, i) + X2F(b2f, m, k, j + 1, i));
          double ubz = 0.5 * (X3F(b3f, m, k, j, i) + X3F(b3f, m, k + 1, j, i));
          if (c->is_iso) {   /* isothermal_mhd.cpp:28-46,95-135 */
            const double b2i = SQR(ubx) + SQR(uby) + SQR(ubz);
            const double dfl_ = fmax(c->dfloor, b2i / c->sigma_max);
            if (ud < dfl_) { ud = dfl_; CC(u0, D.nv, m, IDN, k, j, i) = ud; nd++; }
            double dii = 1.0 / ud;
            CC(w0, D.nv, m, IDN, k, j, i) = ud; CC(w0, D.nv, m, IVX, k, j, i) = dii * umx;
            CC(w0, D.nv, m, IVY, k, j, i) = dii * umy; CC(w0, D.nv, m, IVZ, k, j, i) = dii * umz;
            CC(bcc, 3, m, 0, k, j, i) = ubx; CC(bcc, 3, m, 1, k, j, i) = uby; CC(bcc, 3, m, 2, k, j, i) = ubz;
            for (int n = D.nfl; n < D.nv; ++n)   /* no scalar floor in the isothermal c2p */
              CC(w0, D.nv, m, n, k, j, i) = CC(u0, D.nv, m, n, k, j, i) / ud;
            continue;
          }
          double ue = CC(u0, D.nv, m, IEN, k, j, i);
          int dfl = 0, efl = 0, tfl = 0;
          const double b2 = SQR(ubx) + SQR(uby) + SQR(ubz);
          const double dfloor_ = fmax(c->dfloor, b2 / c->sigma_max);
          double efloor = c->pfloor / (c->gamma - 1.0);
          double tfloor = c->tfloor, sfloor = c->sfloor;
          if (ud < dfloor_) { ud = dfloor_; dfl = 1; }
          double wd = ud;
          double di = 1.0 / ud;
          double wvx = di * umx, wvy = di * umy, wvz = di * umz;
          double e_k = 0.5 * di * (SQR(umx) + SQR(umy) + SQR(umz));
          double e_m = 0.5 * (SQR(ubx) + SQR(uby) + SQR(ubz));
          double we = (ue - e_k - e_m);
          if (we < efloor) { we = efloor; ue = efloor + e_k + e_m; efl = 1; }
          if (gm1 * we * di < tfloor) { we = wd * tfloor / gm1; ue = we + e_k + e_m; tfl = 1; }
          double spe_over_eps = gm1 / pow(wd, gm1);
          double spe = spe_over_eps * we * di;
          if (spe <= sfloor) { we = wd * sfloor / spe_over_eps; efl = 1; }
          if (dfl) { CC(u0, D.nv, m, IDN, k, j, i) = ud; nd++; }
          if (efl) { CC(u0, D.nv, m, IEN, k, j, i) = ue; ne++; }
          if (tfl) { CC(u0, D.nv, m, IEN, k, j, i) = ue; nt++; }
          CC(w0, D.nv, m, IDN, k, j, i) = wd; CC(w0, D.nv, m, IVX, k, j, i) = wvx;
          CC(w0, D.nv, m, IVY, k, j, i) = wvy; CC(w0, D.nv, m, IVZ, k, j, i) = wvz;
          CC(w0, D.nv, m, IEN, k, j, i) = we;
          CC(bcc, 3, m, 0, k, j, i) = ubx; CC(bcc, 3, m, 1, k, j, i) = uby; CC(bcc, 3, m, 2, k, j, i) = ubz;
          for (int n = D.nfl; n < D.nv; ++n) {   /* ideal_mhd.cpp:113-120: scalar floor, then s/d */
            if (CC(u0, D.nv, m, n, k, j, i) < 0.0) CC(u0, D.nv, m, n, k, j, i) = 0.0;
            CC(w0, D.nv, m, n, k, j, i) = CC(u0, D.nv, m, n, k, j, i) / ud;
          }
        }
  if (nfloor) { nfloor[0] = nd; nfloor[1] = ne; nfloor[2] = nt; }
}

/* MHD::NewTimeStep Newtonian ideal branch, mhd_newdt.cpp:31-174; returns min(dt1,dt2,dt3)
 * and the three directional minima in dt3[]. */
double oracle_new_dt(const ocfg *c, const double *w0, const double *bcc, const double *mb_size,
                     double *dt3) {
  odims D = dims(c);
  double dt1 = (double)FLT_MAX, dt2 = (double)FLT_MAX, dtz = (double)FLT_MAX;
  const double gm1 = c->gamma - 1.0;
  for (int m = 0; m < c->nmb; ++m) {
    const double dx1 = mb_size[m * 9 + 6], dx2 = mb_size[m * 9 + 7], dx3 = mb_size[m * 9 + 8];
    for (int k = D.ks; k <= D.ke; ++k)
      for (int j = D.js; j <= D.je; ++j)
        for (int i = D.is; i <= D.ie; ++i) {
          double w_d = CC(w0, D.nv, m, IDN, k, j, i);
          double w_bx = CC(bcc, 3, m, 0, k, j, i), w_by = CC(bcc, 3, m, 1, k, j, i), w_bz = CC(bcc, 3, m, 2, k, j, i);
          double p = c->is_iso ? 0.0 : gm1 * CC(w0, D.nv, m, IEN, k, j, i);
          double cf = c->is_iso ? fast_speed_iso(c->iso_cs, w_d, w_bx, w_by, w_bz) : fast_speed(c->gamma, w_d, p, w_bx, w_by, w_bz);
          double max_dv1 = fabs(CC(w0, D.nv, m, IVX, k, j, i)) + cf;
          cf = c->is_iso ? fast_speed_iso(c->iso_cs, w_d, w_by, w_bz, w_bx) : fast_speed(c->gamma, w_d, p, w_by, w_bz, w_bx);
          double max_dv2 = fabs(CC(w0, D.nv, m, IVY, k, j, i)) + cf;
          cf = c->is_iso ? fast_speed_iso(c->iso_cs, w_d, w_bz, w_bx, w_by) : fast_speed(c->gamma, w_d, p, w_bz, w_bx, w_by);
          double max_dv3 = fabs(CC(w0, D.nv, m, IVZ, k, j, i)) + cf;
          dt1 = fmin((dx1 / max_dv1), dt1);
          dt2 = fmin((dx2 / max_dv2), dt2);
          dtz = fmin((dx3 / max_dv3), dtz);
        }
  }
  if (dt3) { dt3[0] = dt1; dt3[1] = dt2; dt3[2] = dtz; }
  double dtnew = dt1;                                  /* mhd_newdt.cpp:152-155 */
  if (c->nx2 > 1 && dt2 < dtnew) dtnew = dt2;
  if (c->nx3 > 1 && dtz < dtnew) dtnew = dtz;
  return dtnew;
}

/* ------------------------------------------------------------------------------------ */
/* Ghost-zone exchange, same level, all blocks on this "rank".
 * NeighborIndex: nghbr_index.hpp:28-54.  nghbr_gid: [nmb][56] neighbour gid or -1
 * (meshblock.cpp:142-430); gids are 0..nmb-1 here. */
int oracle_neighbor_index(int ix, int iy, int iz, int n1, int n2) {
  if ((abs(ix) + abs(iy) + abs(iz)) == 0) return -1;
  if (abs(ix * iy * iz) > 1) return -1;
  if (iz == 0) {
    if (ix * iy == 0) {
      int subface = n1 + 2 * n2;
      return abs(ix) * 2 * (ix + 1) + abs(iy) * 2 * (iy + 5) + subface;
    } else {
      return 16 + (ix + 1) + 2 * (iy + 1) + n1;
    }
  } else {
    if (ix * iy == 0) {
      int subface = n1 + 2 * n2;
      return 24 + abs(ix) * (ix + 9) + abs(iy) * (iy + 17) + 2 * (iz + 1) + subface;
    } else {
      return 48 + (ix + 1) / 2 + (iy + 1) + 2 * (iz + 1);
    }
  }
}

/* CC: send box buffs_cc.cpp:36-46, recv box buffs_cc.cpp:170-196, pack/unpack order
 * bvals_cc.cpp:104-135,318-396.  The data path (pack into buf, unpack from buf) is kept. */
void oracle_halo_cc(const ocfg *c, double *a, int nv, const int *nghbr_gid) {
  odims D = dims(c);
  const int ng = c->ng, ng1 = ng - 1;
  size_t maxdat = (size_t)D.n1 * D.n2 * ng * nv + 16;
  if ((size_t)D.n1 * D.n3 * ng * nv > maxdat) maxdat = (size_t)D.n1 * D.n3 * ng * nv;
  if ((size_t)D.n2 * D.n3 * ng * nv > maxdat) maxdat = (size_t)D.n2 * D.n3 * ng * nv;
  /* all sends complete before any unpack: recv buffers [m][slot] */
  double **rbuf = calloc((size_t)c->nmb * 56, sizeof(double *));
  for (int m = 0; m < c->nmb; ++m)
    for (int oz = -1; oz <= 1; ++oz) for (int oy = -1; oy <= 1; ++oy) for (int ox = -1; ox <= 1; ++ox) {
      if (!ox && !oy && !oz) continue;
      int n = oracle_neighbor_index(ox, oy, oz, 0, 0);
      int gid = nghbr_gid[m * 56 + n];
      if (gid < 0) continue;
      int dn = oracle_neighbor_index(-ox, -oy, -oz, 0, 0);
      int il = (ox > 0) ? (D.ie - ng1) : D.is, iu = (ox < 0) ? (D.is + ng1) : D.ie;
      int jl = (oy > 0) ? (D.je - ng1) : D.js, ju = (oy < 0) ? (D.js + ng1) : D.je;
      int kl = (oz > 0) ? (D.ke - ng1) : D.ks, ku = (oz < 0) ? (D.ks + ng1) : D.ke;
      int ni = iu - il + 1, nj = ju - jl + 1, nk = ku - kl + 1;
      double *b = malloc(sizeof(double) * (size_t)ni * nj * nk * nv);
      for (int v = 0; v < nv; ++v) for (int k = kl; k <= ku; ++k) for (int j = jl; j <= ju; ++j) for (int i = il; i <= iu; ++i)
        b[(i - il) + ni * ((j - jl) + nj * ((k - kl) + nk * v))] = CC(a, nv, m, v, k, j, i);
      rbuf[(size_t)gid * 56 + dn] = b;
    }
  for (int m = 0; m < c->nmb; ++m)
    for (int oz = -1; oz <= 1; ++oz) for (int oy = -1; oy <= 1; ++oy) for (int ox = -1; ox <= 1; ++ox) {
      if (!ox && !oy && !oz) continue;
      int n = oracle_neighbor_index(ox, oy, oz, 0, 0);
      double *b = rbuf[(size_t)m * 56 + n];
      if (!b) continue;
      int il = (ox == 0) ? D.is : (ox > 0 ? D.ie + 1 : D.is - ng), iu = (ox == 0) ? D.ie : (ox > 0 ? D.ie + ng : D.is - 1);
      int jl = (oy == 0) ? D.js : (oy > 0 ? D.je + 1 : D.js - ng), ju = (oy == 0) ? D.je : (oy > 0 ? D.je + ng : D.js - 1);
      int kl = (oz == 0) ? D.ks : (oz > 0 ? D.ke + 1 : D.ks - ng), ku = (oz == 0) ? D.ke : (oz > 0 ? D.ke + ng : D.ks - 1);
      int ni = iu - il + 1, nj = ju - jl + 1, nk = ku - kl + 1;
      for (int v = 0; v < nv; ++v) for (int k = kl; k <= ku; ++k) for (int j = jl; j <= ju; ++j) for (int i = il; i <= iu; ++i)
        CC(a, nv, m, v, k, j, i) = b[(i - il) + ni * ((j - jl) + nj * ((k - kl) + nk * v))];
      free(b);
    }
  free(rbuf);
}

/* FC: send boxes buffs_fc.cpp:38-83, recv boxes buffs_fc.cpp:396-441 (uniform mesh: the
 * multilevel extensions at :84-98/:442-456 do not apply). */
static void fc_box(const odims *Dp, int ng, int comp, int ox, int oy, int oz, int send, int *b6) {
  const odims D = *Dp;
  const int ng1 = ng - 1;
  int s[3] = {D.is, D.js, D.ks}, e[3] = {D.ie, D.je, D.ke}, o[3] = {ox, oy, oz};
  for (int d = 0; d < 3; ++d) {
    int lo, hi;
    if (send) {
      if (o[d] == 0) { lo = s[d]; hi = e[d] + (comp == d ? 1 : 0); }
      else if (o[d] > 0) { lo = e[d] - ng1; hi = e[d]; }
      else { lo = s[d] + (comp == d ? 1 : 0); hi = s[d] + (comp == d ? ng : ng1); }
    } else {
      if (o[d] == 0) { lo = s[d]; hi = e[d] + (comp == d ? 1 : 0); }
      else if (o[d] > 0) { lo = e[d] + (comp == d ? 2 : 1); hi = e[d] + (comp == d ? ng + 1 : ng); }
      else { lo = s[d] - ng; hi = s[d] - 1; }
    }
    b6[2 * d] = lo; b6[2 * d + 1] = hi;
  }
}

void oracle_halo_fc(const ocfg *c, double *b1f, double *b2f, double *b3f, const int *nghbr_gid) {
  odims D = dims(c);
  double *comp[3] = {b1f, b2f, b3f};
  double **rbuf = calloc((size_t)c->nmb * 56 * 3, sizeof(double *));
  for (int m = 0; m < c->nmb; ++m)
    for (int oz = -1; oz <= 1; ++oz) for (int oy = -1; oy <= 1; ++oy) for (int ox = -1; ox <= 1; ++ox) {
      if (!ox && !oy && !oz) continue;
      int n = oracle_neighbor_index(ox, oy, oz, 0, 0);
      int gid = nghbr_gid[m * 56 + n];
      if (gid < 0) continue;
      int dn = oracle_neighbor_index(-ox, -oy, -oz, 0, 0);
      for (int v = 0; v < 3; ++v) {
        int b6[6]; fc_box(&D, c->ng, v, ox, oy, oz, 1, b6);
        int ni = b6[1] - b6[0] + 1, nj = b6[3] - b6[2] + 1, nk = b6[5] - b6[4] + 1;
        double *b = malloc(sizeof(double) * (size_t)ni * nj * nk);
        for (int k = b6[4]; k <= b6[5]; ++k) for (int j = b6[2]; j <= b6[3]; ++j) for (int i = b6[0]; i <= b6[1]; ++i) {
          double val = (v == 0) ? X1F(b1f, m, k, j, i) : (v == 1) ? X2F(b2f, m, k, j, i) : X3F(b3f, m, k, j, i);
          b[(i - b6[0]) + ni * ((j - b6[2]) + nj * (k - b6[4]))] = val;
        }
        rbuf[((size_t)gid * 56 + dn) * 3 + v] = b;
      }
    }
  (void)comp;
  for (int m = 0; m < c->nmb; ++m)
    for (int oz = -1; oz <= 1; ++oz) for (int oy = -1; oy <= 1; ++oy) for (int ox = -1; ox <= 1; ++ox) {
      if (!ox && !oy && !oz) continue;
      int n = oracle_neighbor_index(ox, oy, oz, 0, 0);
      for (int v = 0; v < 3; ++v) {
        double *b = rbuf[((size_t)m * 56 + n) * 3 + v];
        if (!b) continue;
        int b6[6]; fc_box(&D, c->ng, v, ox, oy, oz, 0, b6);
        int ni = b6[1] - b6[0] + 1, nj = b6[3] - b6[2] + 1;
        for (int k = b6[4]; k <= b6[5]; ++k) for (int j = b6[2]; j <= b6[3]; ++j) for (int i = b6[0]; i <= b6[1]; ++i) {
          double val = b[(i - b6[0]) + ni * ((j - b6[2]) + nj * (k - b6[4]))];
          if (v == 0) X1F(b1f, m, k, j, i) = val; else if (v == 1) X2F(b2f, m, k, j, i) = val; else X3F(b3f, m, k, j, i) = val;
        }
        free(b);
      }
    }
  free(rbuf);
}

/* ------------------------------------------------------------------------------------ */
/* SendE/RecvE: same-level EMF flux correction, flux_correct_fc.cpp:29-366 (pack),
 * :374-437 (RecvAndUnpackFluxFC), :445-644 (SumBoundaryFluxes), :801-1034
 * (AverageBoundaryFluxes); index boxes buffs_fc.cpp:327-368,892-935.  Every block sends the
 * edge EMFs lying on its faces/edges to the face and edge neighbours (slots n<48); the receiver
 * ADDS the buffers in increasing slot order and then scales: x0.5 in face interiors, / nflx on
 * block edges (nflx = 1 + number of contributing neighbours). */
static void emf_box(const odims *Dp, int v, int ox, int oy, int oz, int *b6) {
  const odims D = *Dp;
  int s[3] = {D.is, D.js, D.ks}, e[3] = {D.ie, D.je, D.ke}, o[3] = {ox, oy, oz};
  for (int d = 0; d < 3; ++d) {
    if (o[d] == 0) { b6[2 * d] = s[d]; b6[2 * d + 1] = e[d] + (v == d ? 0 : 1); }  /* edge field v is cell-like along d==v */
    else if (o[d] > 0) { b6[2 * d] = e[d] + 1; b6[2 * d + 1] = e[d] + 1; }
    else { b6[2 * d] = s[d]; b6[2 * d + 1] = s[d]; }
  }
}

static double *emf_at(const odims *Dp, double *e1, double *e2, double *e3, int v, int m, int k, int j, int i) {
  const odims D = *Dp;
  if (v == 0) return &X1E(e1, m, k, j, i);
  if (v == 1) return &X2E(e2, m, k, j, i);
  return &X3E(e3, m, k, j, i);
}

/* which component(s) a slot carries: faces carry the two tangential ones, edges the one along
 * the edge (flux_correct_fc.cpp:517-635) */
static int emf_slot_has(int ox, int oy, int oz, int v) {
  int nz = (ox != 0) + (oy != 0) + (oz != 0);
  int o[3] = {ox, oy, oz};
  if (nz == 1) return o[v] == 0;           /* face: components tangential to the face */
  if (nz == 2) return o[v] == 0;           /* edge: the component along the edge */
  return 0;
}

void oracle_emf_exchange(const ocfg *c, double *e1, double *e2, double *e3, const int *nghbr_gid) {
  odims D = dims(c);
  const int nmb = c->nmb;
  /* 1-D / 2-D: only slots in the active dimensions exist (flux_correct_fc.cpp:898,953,963,1013,1023:
   * multi_d / three_d gates), and a face interior keeps its full extent along an inactive one (:846,873) */
  const int act[3] = {1, c->nx2 > 1, c->nx3 > 1};
  double **rbuf = calloc((size_t)nmb * 48 * 3, sizeof(double *));
  /* pack + send */
  for (int m = 0; m < nmb; ++m)
    for (int oz = -1; oz <= 1; ++oz) for (int oy = -1; oy <= 1; ++oy) for (int ox = -1; ox <= 1; ++ox) {
      int nz = (ox != 0) + (oy != 0) + (oz != 0);
      if (nz == 0 || nz == 3) continue;
      if ((oy && !act[1]) || (oz && !act[2])) continue;
      int n = oracle_neighbor_index(ox, oy, oz, 0, 0);
      int gid = nghbr_gid[m * 56 + n];
      if (gid < 0) continue;
      int dn = oracle_neighbor_index(-ox, -oy, -oz, 0, 0);
      for (int v = 0; v < 3; ++v) {
        if (!emf_slot_has(ox, oy, oz, v)) continue;
        int b6[6]; emf_box(&D, v, ox, oy, oz, b6);
        int ni = b6[1] - b6[0] + 1, nj = b6[3] - b6[2] + 1, nk = b6[5] - b6[4] + 1;
        double *b = malloc(sizeof(double) * (size_t)ni * nj * nk);
        for (int k = b6[4]; k <= b6[5]; ++k) for (int j = b6[2]; j <= b6[3]; ++j) for (int i = b6[0]; i <= b6[1]; ++i)
          b[(i - b6[0]) + ni * ((j - b6[2]) + nj * (k - b6[4]))] = *emf_at(&D, e1, e2, e3, v, m, k, j, i);
        rbuf[((size_t)gid * 48 + dn) * 3 + v] = b;
      }
    }
  /* sum in increasing slot order (SumBoundaryFluxes), counting contributions per block edge */
  int *nflx = malloc(sizeof(int) * (size_t)nmb * 48);
  for (int q = 0; q < nmb * 48; ++q) nflx[q] = 1;
  for (int m = 0; m < nmb; ++m)
    for (int n = 0; n < 48; ++n) {
      int ox = 0, oy = 0, oz = 0, found = 0;
      for (int z = -1; z <= 1 && !found; ++z) for (int y = -1; y <= 1 && !found; ++y) for (int x = -1; x <= 1 && !found; ++x)
        if ((x || y || z) && oracle_neighbor_index(x, y, z, 0, 0) == n) { ox = x; oy = y; oz = z; found = 1; }
      if (!found) continue;
      if ((oy && !act[1]) || (oz && !act[2])) continue;
      if (nghbr_gid[m * 56 + n] < 0) continue;
      int nz = (ox != 0) + (oy != 0) + (oz != 0);
      if (nz == 1) {  /* a face neighbour also contributes to the four edges of that face */
        for (int t = -1; t <= 1; t += 2) {
          if (ox) { if (act[1]) nflx[m * 48 + oracle_neighbor_index(ox, t, 0, 0, 0)]++; if (act[2]) nflx[m * 48 + oracle_neighbor_index(ox, 0, t, 0, 0)]++; }
          if (oy) { nflx[m * 48 + oracle_neighbor_index(t, oy, 0, 0, 0)]++; if (act[2]) nflx[m * 48 + oracle_neighbor_index(0, oy, t, 0, 0)]++; }
          if (oz) { nflx[m * 48 + oracle_neighbor_index(t, 0, oz, 0, 0)]++; nflx[m * 48 + oracle_neighbor_index(0, t, oz, 0, 0)]++; }
        }
      } else {
        nflx[m * 48 + n]++;
      }
      for (int v = 0; v < 3; ++v) {
        double *b = rbuf[((size_t)m * 48 + n) * 3 + v];
        if (!b) continue;
        int b6[6]; emf_box(&D, v, ox, oy, oz, b6);
        int ni = b6[1] - b6[0] + 1, nj = b6[3] - b6[2] + 1;
        for (int k = b6[4]; k <= b6[5]; ++k) for (int j = b6[2]; j <= b6[3]; ++j) for (int i = b6[0]; i <= b6[1]; ++i)
          *emf_at(&D, e1, e2, e3, v, m, k, j, i) += b[(i - b6[0]) + ni * ((j - b6[2]) + nj * (k - b6[4]))];
        free(b);
      }
    }
  /* AverageBoundaryFluxes */
  for (int m = 0; m < nmb; ++m)
    for (int oz = -1; oz <= 1; ++oz) for (int oy = -1; oy <= 1; ++oy) for (int ox = -1; ox <= 1; ++ox) {
      int nz = (ox != 0) + (oy != 0) + (oz != 0);
      if (nz == 0 || nz == 3) continue;
      if ((oy && !act[1]) || (oz && !act[2])) continue;
      int n = oracle_neighbor_index(ox, oy, oz, 0, 0);
      for (int v = 0; v < 3; ++v) {
        if (!emf_slot_has(ox, oy, oz, v)) continue;
        int b6[6]; emf_box(&D, v, ox, oy, oz, b6);
        if (nz == 1) {
          if (nghbr_gid[m * 56 + n] < 0) continue;      /* only where a same-level neighbour exists */
          /* interior of the face: drop the two end lines in the OTHER tangential direction */
          int o[3] = {ox, oy, oz};
          for (int d = 0; d < 3; ++d) if (o[d] == 0 && d != v && act[d]) { b6[2 * d] += 1; b6[2 * d + 1] -= 1; }
          for (int k = b6[4]; k <= b6[5]; ++k) for (int j = b6[2]; j <= b6[3]; ++j) for (int i = b6[0]; i <= b6[1]; ++i)
            *emf_at(&D, e1, e2, e3, v, m, k, j, i) *= 0.5;
        } else {
          for (int k = b6[4]; k <= b6[5]; ++k) for (int j = b6[2]; j <= b6[3]; ++j) for (int i = b6[0]; i <= b6[1]; ++i)
            *emf_at(&D, e1, e2, e3, v, m, k, j, i) /= (double)nflx[m * 48 + n];
        }
      }
    }
  free(nflx);
  free(rbuf);
}

/* ------------------------------------------------------------------------------------ */
/* Task 26 ApplyPhysicalBCs (mhd_tasks.cpp:504-520): HydroBCs (bvals/physics/hydro_bcs.cpp:68-335)
 * then BFieldBCs (bvals/physics/bfield_bcs.cpp:67-358), skipped on strictly periodic meshes.
 * Flags follow BoundaryFlag (bvals.hpp:19): block 0, reflect 1, inflow 2, outflow 3, diode 4,
 * user 5, periodic 6, shear_periodic 7, vacuum 8.  The passes run x1, x2, x3 in that order, each
 * over the FULL transverse extent (ghosts included), so edge/corner ghosts are filled from the
 * ghosts the earlier passes wrote.  A pass is gated on the MESH flag of the inner face and then
 * switches on the per-block flag of each face (block faces: untouched).  inflow/user need
 * problem-generator data and are not restated (oracle_apply_bcs returns 1 for them). */
enum { BC_BLOCK = 0, BC_REFLECT = 1, BC_INFLOW = 2, BC_OUTFLOW = 3, BC_DIODE = 4, BC_USER = 5,
       BC_PERIODIC = 6, BC_SHEAR = 7, BC_VACUUM = 8 };

/* one ghost line: a[base + x*st] for x along the BC direction; cell-like (face_like=0) or the
 * face component normal to the boundary (face_like=1); s,e = first/last ACTIVE cell index */
/* clamp: 1 = normal momentum (diode clamps it), 0 = other conserved variable, -1 = magnetic
 * field (diode and vacuum both act as outflow, bfield_bcs.cpp:90-92) */
static void bc_line(double *a, size_t st, int ng, int s, int e, int side, int flag, int face_like,
                    int flip, int clamp) {
  for (int g = 0; g < ng; ++g) {
    int dst, mir, edge;
    if (side == 0) { dst = s - g - 1; mir = s + g + face_like; edge = s; }
    else { dst = e + g + 1 + face_like; mir = e - g; edge = e + face_like; }
    switch (flag) {
      case BC_REFLECT: a[dst * st] = flip ? -a[mir * st] : a[mir * st]; break;
      case BC_OUTFLOW: a[dst * st] = a[edge * st]; break;
      case BC_DIODE:
        if (clamp > 0) a[dst * st] = (side == 0) ? fmin(0.0, a[edge * st]) : fmax(0.0, a[edge * st]);
        else a[dst * st] = a[edge * st];
        break;
      case BC_VACUUM: a[dst * st] = (clamp >= 0) ? 0.0 : a[edge * st]; break;  /* B: vacuum == outflow */
      default: break;
    }
  }
}

/* all ghost lines of one array [nmb][nvar][x3][x2][x1] for BC direction d.  kind: -1 cell-centred
 * conserved variables, 0/1/2 face component x1f/x2f/x3f */
static void bc_array(const ocfg *c, double *a, int nvar, int kind, int d, const int *mb_bcs) {
  odims D = dims(c);
  int ext[3] = {D.n1 + (kind == 0), D.n2 + (kind == 1), D.n3 + (kind == 2)};
  int s3[3] = {D.is, D.js, D.ks}, e3[3] = {D.ie, D.je, D.ke};
  size_t st[3] = {1, (size_t)ext[0], (size_t)ext[0] * ext[1]};
  size_t blk = st[2] * ext[2];
  int ta = (d + 1) % 3, tb = (d + 2) % 3;
  for (int m = 0; m < c->nmb; ++m)
    for (int n = 0; n < nvar; ++n)
      for (int qb = 0; qb < ext[tb]; ++qb) for (int qa = 0; qa < ext[ta]; ++qa) {
        double *line = a + ((size_t)m * nvar + n) * blk + qa * st[ta] + qb * st[tb];
        for (int side = 0; side < 2; ++side) {
          int flag = mb_bcs[m * 6 + 2 * d + side];
          if (kind < 0) bc_line(line, st[d], c->ng, s3[d], e3[d], side, flag, 0, n == IVX + d, n == IVX + d);
          else bc_line(line, st[d], c->ng, s3[d], e3[d], side, flag, kind == d, kind == d, -1);
        }
      }
}

int oracle_apply_bcs(const ocfg *c, const int *mesh_bcs, const int *mb_bcs, double *u0, double *b1f,
                     double *b2f, double *b3f) {
  for (int q = 0; q < 6 * c->nmb; ++q)
    if (mb_bcs[q] == BC_INFLOW || mb_bcs[q] == BC_USER || mb_bcs[q] == BC_SHEAR) return 1;
  for (int d = 0; d < 3; ++d) {        /* hydro_bcs.cpp:80-81,167,252: gate on the mesh flag */
    if ((d == 1 && c->nx2 <= 1) || (d == 2 && c->nx3 <= 1)) continue;   /* :164 one_d, :249 two_d */
    if (mesh_bcs[2 * d] == BC_PERIODIC || (d == 0 && mesh_bcs[0] == BC_SHEAR)) continue;
    bc_array(c, u0, (c->is_iso ? 4 : 5) + c->nscalars, -1, d, mb_bcs);
  }
  for (int d = 0; d < 3; ++d) {        /* bfield_bcs.cpp:78,169,262 */
    if ((d == 1 && c->nx2 <= 1) || (d == 2 && c->nx3 <= 1)) continue;   /* :165 one_d, :258 two_d */
    if (mesh_bcs[2 * d] == BC_PERIODIC) continue;
    bc_array(c, b1f, 1, 0, d, mb_bcs);
    bc_array(c, b2f, 1, 1, d, mb_bcs);
    bc_array(c, b3f, 1, 2, d, mb_bcs);
  }
  return 0;
}

/* physical BCs used by oracle_stage / oracle_run; NULL = strictly periodic mesh */
static const int *g_mesh_bcs = NULL, *g_mb_bcs = NULL;
void oracle_set_bcs(const int *mesh_bcs, const int *mb_bcs) { g_mesh_bcs = mesh_bcs; g_mb_bcs = mb_bcs; }

/* MHD::FOFC, mhd_fofc.cpp:31-489 (Newtonian, no excision).  Estimates the updated state over
 * [is-1,ie+1] (x [js-1,je+1] x [ks-1,ke+1] in active dimensions), flags the cells whose c2p would
 * need a floor (ideal_mhd.cpp:57-88 / isothermal_mhd.cpp:105-118 with only_testfloors), then replaces
 * the fluxes and face EMFs on all faces of flagged cells with first-order LLF values.  Returns
 * the number of flagged cells (ecounter.nfofc). */
static int g_nfofc = 0;
int oracle_get_nfofc(int reset) { int n = g_nfofc; if (reset) g_nfofc = 0; return n; }

static void fofc_llf(const ocfg *c, int dir, const double *w0, const double *bcc0, int nv, int m, int kl_, int jl_, int il_,
                     int kr_, int jr_, int ir_, double bxi, double *f5, double *ey, double *ez) {
  odims D = dims(c);
  const int ivx = IVX + dir, ivy = IVX + (dir + 1) % 3, ivz = IVX + (dir + 2) % 3;
  const int iby = (dir + 1) % 3, ibz = (dir + 2) % 3;
  prim1d l, r;
  l.d = CC(w0, nv, m, IDN, kl_, jl_, il_); l.vx = CC(w0, nv, m, ivx, kl_, jl_, il_); l.vy = CC(w0, nv, m, ivy, kl_, jl_, il_);
  l.vz = CC(w0, nv, m, ivz, kl_, jl_, il_); l.e = c->is_iso ? 0.0 : CC(w0, nv, m, IEN, kl_, jl_, il_);
  l.by = CC(bcc0, 3, m, iby, kl_, jl_, il_); l.bz = CC(bcc0, 3, m, ibz, kl_, jl_, il_);
  r.d = CC(w0, nv, m, IDN, kr_, jr_, ir_); r.vx = CC(w0, nv, m, ivx, kr_, jr_, ir_); r.vy = CC(w0, nv, m, ivy, kr_, jr_, ir_);
  r.vz = CC(w0, nv, m, ivz, kr_, jr_, ir_); r.e = c->is_iso ? 0.0 : CC(w0, nv, m, IEN, kr_, jr_, ir_);
  r.by = CC(bcc0, 3, m, iby, kr_, jr_, ir_); r.bz = CC(bcc0, 3, m, ibz, kr_, jr_, ir_);
  cons1d fx;
  if (c->is_iso) llf_face_iso(c, &l, &r, bxi, &fx); else llf_face(c, &l, &r, bxi, &fx);
  f5[IDN] = fx.d; f5[ivx] = fx.mx; f5[ivy] = fx.my; f5[ivz] = fx.mz; f5[IEN] = fx.e;
  *ey = fx.by; *ez = fx.bz;     /* mhd_fofc.cpp:196-197: the single-state LLF already carries the EMF signs */
}

int oracle_fofc(const ocfg *c, const double *u0, const double *u1, const double *w0, const double *bcc0,
                const double *b0x1f, const double *b0x2f, const double *b0x3f,
                const double *b1x1f, const double *b1x2f, const double *b1x3f,
                double *flx1, double *flx2, double *flx3, double *e3x1, double *e2x1, double *e1x2,
                double *e3x2, double *e2x3, double *e1x3, const double *mb_size, double gam0, double gam1,
                double beta_dt) {
  odims D = dims(c);
  const int multi_d = c->nx2 > 1, three_d = c->nx3 > 1;
  const int il = D.is - 1, iu = D.ie + 1, jl = D.js - multi_d, ju = D.je + multi_d, kl = D.ks - three_d, ku = D.ke + three_d;
  unsigned char *flag = calloc((size_t)c->nmb * D.cc, 1);
  int count = 0;
  const double gm1 = c->gamma - 1.0;
  for (int m = 0; m < c->nmb; ++m) {
    const double dtodx1 = beta_dt / mb_size[m * 9 + 6], dtodx2 = beta_dt / mb_size[m * 9 + 7], dtodx3 = beta_dt / mb_size[m * 9 + 8];
    for (int k = kl; k <= ku; ++k)
      for (int j = jl; j <= ju; ++j)
        for (int i = il; i <= iu; ++i) {
          double ut[5] = {0, 0, 0, 0, 0};
          for (int n = 0; n < D.nfl; ++n) {   /* mhd_fofc.cpp:78-88 */
            double divf = dtodx1 * (F1(flx1, m, n, k, j, i + 1) - F1(flx1, m, n, k, j, i));
            if (multi_d) divf += dtodx2 * (F2(flx2, m, n, k, j + 1, i) - F2(flx2, m, n, k, j, i));
            if (three_d) divf += dtodx3 * (F3(flx3, m, n, k + 1, j, i) - F3(flx3, m, n, k, j, i));
            ut[n] = gam0 * CC(u0, D.nv, m, n, k, j, i) + gam1 * CC(u1, D.nv, m, n, k, j, i) - divf;
          }
          double b1old = 0.5 * (X1F(b1x1f, m, k, j, i) + X1F(b1x1f, m, k, j, i + 1));   /* :91-107 */
          double b2old = 0.5 * (X2F(b1x2f, m, k, j, i) + X2F(b1x2f, m, k, j + 1, i));
          double b3old = 0.5 * (X3F(b1x3f, m, k, j, i) + X3F(b1x3f, m, k + 1, j, i));
          double bx = gam0 * CC(bcc0, 3, m, 0, k, j, i) + gam1 * b1old;
          double by = gam0 * CC(bcc0, 3, m, 1, k, j, i) + gam1 * b2old;
          double bz = gam0 * CC(bcc0, 3, m, 2, k, j, i) + gam1 * b3old;
          by += dtodx1 * (S4(e3x1, m, k, j, i + 1) - S4(e3x1, m, k, j, i));
          bz -= dtodx1 * (S4(e2x1, m, k, j, i + 1) - S4(e2x1, m, k, j, i));
          if (multi_d) {
            bx -= dtodx2 * (S4(e3x2, m, k, j + 1, i) - S4(e3x2, m, k, j, i));
            bz += dtodx2 * (S4(e1x2, m, k, j + 1, i) - S4(e1x2, m, k, j, i));
          }
          if (three_d) {
            bx += dtodx3 * (S4(e2x3, m, k + 1, j, i) - S4(e2x3, m, k, j, i));
            by -= dtodx3 * (S4(e1x3, m, k + 1, j, i) - S4(e1x3, m, k, j, i));
          }
          int used = 0;
          if (c->is_iso) {   /* isothermal_mhd.cpp:95-118: the floor test averages the b0 FACES it is handed */
            double fbx = 0.5 * (X1F(b0x1f, m, k, j, i) + X1F(b0x1f, m, k, j, i + 1));
            double fby = 0.5 * (X2F(b0x2f, m, k, j, i) + X2F(b0x2f, m, k, j + 1, i));
            double fbz = 0.5 * (X3F(b0x3f, m, k, j, i) + X3F(b0x3f, m, k + 1, j, i));
            const double b2 = SQR(fbx) + SQR(fby) + SQR(fbz);
            if (ut[IDN] < fmax(c->dfloor, b2 / c->sigma_max)) used = 1;
          } else {           /* ideal_c2p_mhd.hpp:20-67 on (utest, bcctest) */
            const double b2 = SQR(bx) + SQR(by) + SQR(bz);
            const double dfloor_ = fmax(c->dfloor, b2 / c->sigma_max);
            double efloor = c->pfloor / (c->gamma - 1.0);
            double ud = ut[IDN];
            if (ud < dfloor_) { ud = dfloor_; used = 1; }
            double wd = ud, di = 1.0 / ud;
            double e_k = 0.5 * di * (SQR(ut[IM1]) + SQR(ut[IM2]) + SQR(ut[IM3]));
            double e_m = 0.5 * (SQR(bx) + SQR(by) + SQR(bz));
            double we = (ut[IEN] - e_k - e_m);
            if (we < efloor) { we = efloor; used = 1; }
            if (gm1 * we * di < c->tfloor) { we = wd * c->tfloor / gm1; used = 1; }
            double spe_over_eps = gm1 / pow(wd, gm1);
            double spe = spe_over_eps * we * di;
            if (spe <= c->sfloor) used = 1;
          }
          if (used) { flag[(size_t)m * D.cc + ((size_t)k * D.n2 + j) * D.n1 + i] = 1; count++; }
        }
  }
  /* mhd_fofc.cpp:131-474: lower faces of flagged cells, then upper faces (a face shared by two flagged
   * cells gets the same value twice) */
  for (int pass = 0; pass < 2; ++pass)
    for (int m = 0; m < c->nmb; ++m)
      for (int k = kl; k <= ku; ++k)
        for (int j = jl; j <= ju; ++j)
          for (int i = il; i <= iu; ++i) {
            if (!flag[(size_t)m * D.cc + ((size_t)k * D.n2 + j) * D.n1 + i]) continue;
            double f5[5], ey, ez;
            const int s = pass;   /* 0: faces (i,j,k); 1: faces (i+1), (j+1), (k+1) */
            fofc_llf(c, 0, w0, bcc0, D.nv, m, k, j, i - 1 + s, k, j, i + s, X1F(b0x1f, m, k, j, i + s), f5, &ey, &ez);
            for (int n = 0; n < D.nfl; ++n) F1(flx1, m, n, k, j, i + s) = f5[n];
            S4(e3x1, m, k, j, i + s) = ey; S4(e2x1, m, k, j, i + s) = ez;
            if (multi_d) {
              fofc_llf(c, 1, w0, bcc0, D.nv, m, k, j - 1 + s, i, k, j + s, i, X2F(b0x2f, m, k, j + s, i), f5, &ey, &ez);
              for (int n = 0; n < D.nfl; ++n) F2(flx2, m, n, k, j + s, i) = f5[n];
              S4(e1x2, m, k, j + s, i) = ey; S4(e3x2, m, k, j + s, i) = ez;
            }
            if (three_d) {
              fofc_llf(c, 2, w0, bcc0, D.nv, m, k - 1 + s, j, i, k + s, j, i, X3F(b0x3f, m, k + s, j, i), f5, &ey, &ez);
              for (int n = 0; n < D.nfl; ++n) F3(flx3, m, n, k + s, j, i) = f5[n];
              S4(e2x3, m, k + s, j, i) = ey; S4(e1x3, m, k + s, j, i) = ez;
            }
          }
  free(flag);
  g_nfofc += count;
  return count;
}

/* ------------------------------------------------------------------------------------ */
/* One RK stage = the "stagen" task chain of mhd_tasks.cpp:48-75 on a uniform mesh
 * (tasks 1,2,5,10/11,14,15/16,17,21/22,27).  SendE/RecvE is a bitwise identity only when the
 * faces shared by neighbouring blocks START bit-identical (true for the blast ICs, SURVEY.md
 * section 0 fact 6; false for vector-potential ICs evaluated on both sides of a periodic
 * boundary, e.g. linear waves), so it is always executed, as in the reference. */
typedef struct {
  double *u0, *u1, *w0, *bcc0;
  double *b0x1f, *b0x2f, *b0x3f, *b1x1f, *b1x2f, *b1x3f;
} ostate;

void oracle_stage(const ocfg *c, ostate *s, const double *mb_size, const int *nghbr_gid, int stage,
                  double gam0, double gam1, double beta_dt) {
  odims D = dims(c);
  size_t nm = (size_t)c->nmb;
  if (stage == 1) { /* CopyCons mhd_tasks.cpp:162-170 */
    memcpy(s->u1, s->u0, 8 * D.nv * nm * D.cc);
    memcpy(s->b1x1f, s->b0x1f, 8 * nm * D.f1);
    memcpy(s->b1x2f, s->b0x2f, 8 * nm * D.f2);
    memcpy(s->b1x3f, s->b0x3f, 8 * nm * D.f3);
  }
  double *flx1 = calloc(D.nv * nm * D.f1, 8), *flx2 = calloc(D.nv * nm * D.f2, 8), *flx3 = calloc(D.nv * nm * D.f3, 8);
  double *fe[6];
  for (int q = 0; q < 6; ++q) fe[q] = calloc(nm * D.cc, 8);
  double *e1 = calloc(nm * D.e1, 8), *e2 = calloc(nm * D.e2, 8), *e3 = calloc(nm * D.e3, 8);
  oracle_fluxes(c, s->w0, s->bcc0, s->b0x1f, s->b0x2f, s->b0x3f, flx1, flx2, flx3, fe[0], fe[1], fe[2], fe[3], fe[4], fe[5]);
  if (c->fofc)   /* mhd_tasks.cpp:208-211 */
    oracle_fofc(c, s->u0, s->u1, s->w0, s->bcc0, s->b0x1f, s->b0x2f, s->b0x3f, s->b1x1f, s->b1x2f, s->b1x3f, flx1, flx2, flx3,
                fe[0], fe[1], fe[2], fe[3], fe[4], fe[5], mb_size, gam0, gam1, beta_dt);
  oracle_rk_update(c, s->u0, s->u1, flx1, flx2, flx3, mb_size, gam0, gam1, beta_dt);
  oracle_halo_cc(c, s->u0, D.nv, nghbr_gid);
  oracle_corner_e(c, s->w0, s->bcc0, flx1, flx2, flx3, fe[0], fe[1], fe[2], fe[3], fe[4], fe[5], e1, e2, e3);
  oracle_emf_exchange(c, e1, e2, e3, nghbr_gid);
  oracle_ct(c, s->b0x1f, s->b0x2f, s->b0x3f, s->b1x1f, s->b1x2f, s->b1x3f, e1, e2, e3, mb_size, gam0, gam1, beta_dt);
  oracle_halo_fc(c, s->b0x1f, s->b0x2f, s->b0x3f, nghbr_gid);
  if (g_mesh_bcs) oracle_apply_bcs(c, g_mesh_bcs, g_mb_bcs, s->u0, s->b0x1f, s->b0x2f, s->b0x3f);
  int nf[3];
  oracle_cons2prim(c, s->u0, s->b0x1f, s->b0x2f, s->b0x3f, s->w0, s->bcc0, 0, D.n1 - 1, 0, D.n2 - 1, 0, D.n3 - 1, nf);
  free(flx1); free(flx2); free(flx3);
  for (int q = 0; q < 6; ++q) free(fe[q]);
  free(e1); free(e2); free(e3);
}

/* Driver::Execute + Mesh::NewTimeStep, driver.cpp:93-130,380-459, mesh.cpp:573-643.
 * integrator: 1 rk1, 2 rk2, 3 rk3.  *dt is the step to take (already cfl-limited);
 * on return *dt is the next step and *time the new time.  Returns cycles done. */
int oracle_run(const ocfg *c, ostate *s, const double *mb_size, const int *nghbr_gid,
               int integrator, double cfl, double tlim, int ncycles, double *time, double *dt,
               double *dt_hist) {
  double g0[3] = {0.0, 0.0, 0.0}, g1[3] = {1.0, 0.0, 0.0}, be[3] = {1.0, 0.0, 0.0};
  int nst = integrator;
  if (integrator == 2) { g0[1] = 0.5; g1[1] = 0.5; be[1] = 0.5; }
  if (integrator == 3) { g0[1] = 0.25; g1[1] = 0.75; be[1] = 0.25; g0[2] = 2.0 / 3.0; g1[2] = 1.0 / 3.0; be[2] = 2.0 / 3.0; }
  int n = 0;
  while (*time < tlim && n < ncycles) {
    if (dt_hist) dt_hist[n] = *dt;
    for (int st = 1; st <= nst; ++st) oracle_stage(c, s, mb_size, nghbr_gid, st, g0[st - 1], g1[st - 1], be[st - 1] * (*dt));
    double dtnew = oracle_new_dt(c, s->w0, s->bcc0, mb_size, NULL);
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

/* first-cycle dt: Driver::Initialize -> NewTimeStep with dt = FLT_MAX sentinel
 * (build_tree.cpp:301, mesh.cpp:573-640) */
double oracle_initial_dt(const ocfg *c, const ostate *s, const double *mb_size, double cfl,
                         double tlim, double time) {
  double dtnew = oracle_new_dt(c, s->w0, s->bcc0, mb_size, NULL);
  double d = 2.0 * (double)FLT_MAX;
  double lim = cfl * dtnew;
  if (lim < d) d = lim;
  if ((time < tlim) && ((time + d) > tlim)) d = tlim - time;
  return d;
}

#include "hydro_oracle.c"
#include "smr_oracle.c"
#include "smr_fc_oracle.c"
#include "amr_oracle.c"
