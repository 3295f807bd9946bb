// mhd_kernels.h -- host-side declarations of the kernel launchers (internal to the library).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace akb {

struct EosParams;

// number of partial-result slots of the new-dt min-reduction (3 doubles each)
#define DT_SLOTS 64

struct Geom {  // RegionIndcs (src/mesh/mesh.hpp:35-41) + derived array extents
  int nmb, ng, nx1, nx2, nx3;
  int nvar;         // fluid variables + passive scalars per cell-centred / flux array
  int nfl;          // fluid variables: 5 (ideal gas) or 4 (isothermal EOS: no IEN)
  int n1, n2, n3;
  int is, ie, js, je, ks, ke;
  long cc;          // cells per variable per block
  long f1, f2, f3;  // faces per block
  long e1, e2, e3;  // edges per block
  // dimensionality (mesh.cpp:70-79 multi_d/three_d) and the transverse extension of the flux
  // sweeps: MHD computes fluxes one cell beyond the active zone in every ACTIVE transverse
  // direction (the corner EMFs need them, mhd_fluxes.cpp:100-248); hydro does not
  // (hydro_fluxes.cpp:91-101).  Inactive dimensions have n=1, s=e=0.
  int multi_d, three_d;
  int ext1, ext2, ext3;
  int next;         // 1 with FOFC: the sweeps cover one more face on each side along the face normal
};

struct Arrays {  // mirror of ak_mhd_arrays (include/athenak_b200.h)
  double *u0, *u1, *w0, *bcc0;
  double *b0x1f, *b0x2f, *b0x3f, *b1x1f, *b1x2f, *b1x3f;
  double *flx1, *flx2, *flx3;
  double *e3x1, *e2x1, *e1x2, *e3x2, *e2x3, *e1x3;
  double *e1, *e2, *e3;
  const double *mb_size;
};

cudaError_t launch_fluxes(int recon, int rs, const Geom &g, const EosParams &eos, const Arrays &a, cudaStream_t st);
// flux + update fusion: sweeps that turn the fluxes into the divergence terms on chip and finish the conserved update in the
// x3 sweep; du = scratch of 2 * nmb * nfl * cc doubles.  3-D MHD without passive scalars / FOFC (fluxes_update_supported).
bool fluxes_update_supported(const Geom &g);
cudaError_t launch_fluxes_update(int recon, int rs, const Geom &g, const EosParams &eos, const Arrays &a, double *du, bool save,
                                 double gam0, double gam1, double beta_dt, cudaStream_t st);
// hydro twin (hydro_kernels.cu): rs = RS_HYD_*; a.bcc0, face fields and EMF arrays are unused
cudaError_t launch_hydro_fluxes(int recon, int rs, const Geom &g, const EosParams &eos, const Arrays &a, cudaStream_t st);
cudaError_t launch_hydro_cons2prim(const Geom &g, const EosParams &eos, const Arrays &a, int il, int iu, int jl, int ju,
                                   int kl, int ku, int *counters, bool newdt, unsigned long long *dtmin, cudaStream_t st);
cudaError_t launch_hydro_newdt(const Geom &g, const EosParams &eos, const Arrays &a, unsigned long long *dtmin, cudaStream_t st);
cudaError_t launch_corner_e(const Geom &g, const Arrays &a, cudaStream_t st);
cudaError_t launch_update_ct(const Geom &g, const Arrays &a, bool do_u, bool do_b, bool save, double gam0, double gam1,
                             double beta_dt, cudaStream_t st);
cudaError_t launch_cons2prim(const Geom &g, const EosParams &eos, const Arrays &a, int il, int iu, int jl, int ju,
                             int kl, int ku, int *counters, bool newdt, unsigned long long *dtmin, cudaStream_t st);
cudaError_t launch_newdt(const Geom &g, const EosParams &eos, const Arrays &a, unsigned long long *dtmin, cudaStream_t st);
// history sums (hst.cu): scratch holds (nblocks+1)*11 doubles, the 11 results are its last row
cudaError_t launch_hst_sums(const Geom &g, const Arrays &a, bool mhd, double *scratch, int nblocks, cudaStream_t st);
// internal solver code: public AK_RSOLVER_* value, + 20 for the isothermal variant (iso_device.cuh)
// FOFC (fofc.cu): flag = nmb*cc bytes, counter = device int accumulating ecounter.nfofc
cudaError_t launch_fofc(const Geom &g, const EosParams &eos, const Arrays &a, double gam0, double gam1, double beta_dt,
                        unsigned char *flag, int *counter, cudaStream_t st);
inline int internal_rsolver(int rsolver, int eos) { return rsolver + (eos == 1 ? 20 : 0); }
cudaError_t launch_copy(const double *src, double *dst, long n, cudaStream_t st);
int set_march_tma(int on);     // x2/x3 sweeps through the TMA bulk-copy form (flux_march_tma_kernel); returns the previous setting

// fused stage tail (stage_tail.cu): corner EMFs + conserved update + CT + ConsToPrim (+ new dt) in one kernel, out of place
// for the face field and the primitives; shell kernels for the outermost layer
struct TailArrays {
  const double *u_in, *u1r;          // u1r: register copy read at stages > 1 (nullptr at stage 1)
  double *u_out, *u1w;               // u1w: stage-1 save of the old state (nullptr otherwise, or when the host rotates u too)
  const double *b_in[3], *b1[3];     // b1 == nullptr at stage 1 (the register IS the old field)
  double *b_out[3];
  const double *w_in, *bcc_in;
  double *w_out, *bcc_out;
  const double *flx[3];
  const double *e3x1, *e2x1, *e1x2, *e3x2, *e2x3, *e1x3;
  double *e1, *e2, *e3;
  const double *mb_size;
};
bool stage_tail_supported(const Geom &g);
cudaError_t launch_stage_tail(const Geom &g, const EosParams &eos, const TailArrays &A, bool stage1, double gam0, double gam1,
                              double beta_dt, int *counters, bool newdt, unsigned long long *dtmin, bool split, cudaStream_t st);
cudaError_t launch_shell_ct(const Geom &g, const TailArrays &A, bool stage1, double gam0, double gam1, double beta_dt, cudaStream_t st);
cudaError_t launch_shell_c2p(const Geom &g, const EosParams &eos, const Arrays &a, int *counters, bool newdt, unsigned long long *dtmin,
                             cudaStream_t st);

}  // namespace akb
