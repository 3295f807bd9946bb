// mhd_flux_update_a.cu -- instantiations of the flux + update fusion sweeps (AccArgs in flux_kernels.cuh) for RECON_DC, RECON_PLM.
// Separate translation units only to keep the build parallel; see mhd_kernels.cu for the kernel inventory.
#include "mhd_kernels.h"
#include "flux_kernels.cuh"

namespace akb {

template <int RECON>
static cudaError_t launch_fluxes_update_r(int rs, const Geom &g, const EosParams &eos, const Arrays &a, const AccArgs &acc, cudaStream_t st) {
  switch (rs) {
    case RS_LLF: return launch_fluxes_t<RECON, RS_LLF, true>(g, eos, a, st, &acc);
    case RS_HLLE: return launch_fluxes_t<RECON, RS_HLLE, true>(g, eos, a, st, &acc);
    case RS_HLLD: return launch_fluxes_t<RECON, RS_HLLD, true>(g, eos, a, st, &acc);
    case RS_LLF_ISO: return launch_fluxes_t<RECON, RS_LLF_ISO, true>(g, eos, a, st, &acc);
    case RS_HLLE_ISO: return launch_fluxes_t<RECON, RS_HLLE_ISO, true>(g, eos, a, st, &acc);
    case RS_HLLD_ISO: return launch_fluxes_t<RECON, RS_HLLD_ISO, true>(g, eos, a, st, &acc);
  }
  return cudaErrorInvalidValue;
}
cudaError_t launch_fluxes_update_dc(int rs, const Geom &g, const EosParams &eos, const Arrays &a, const AccArgs &acc, cudaStream_t st) { return launch_fluxes_update_r<RECON_DC>(rs, g, eos, a, acc, st); }
cudaError_t launch_fluxes_update_plm(int rs, const Geom &g, const EosParams &eos, const Arrays &a, const AccArgs &acc, cudaStream_t st) { return launch_fluxes_update_r<RECON_PLM>(rs, g, eos, a, acc, st); }

cudaError_t launch_fluxes_update_ppm4(int, const Geom &, const EosParams &, const Arrays &, const AccArgs &, cudaStream_t);
cudaError_t launch_fluxes_update_ppmx(int, const Geom &, const EosParams &, const Arrays &, const AccArgs &, cudaStream_t);
cudaError_t launch_fluxes_update_wenoz(int, const Geom &, const EosParams &, const Arrays &, const AccArgs &, cudaStream_t);
cudaError_t launch_fluxes_update_teno(int, const Geom &, const EosParams &, const Arrays &, const AccArgs &, cudaStream_t);

bool fluxes_update_supported(const Geom &g) { return g.three_d && g.multi_d && g.nvar == g.nfl && g.next == 0 && g.ext1 == 1; }

cudaError_t launch_fluxes_update(int recon, int rs, const Geom &g, const EosParams &eos, const Arrays &a, double *du, bool save,
                                 double gam0, double gam1, double beta_dt, cudaStream_t st) {
  if (!fluxes_update_supported(g)) return cudaErrorInvalidConfiguration;
  AccArgs acc;
  acc.du1 = du; acc.du2 = du + (long)g.nmb * g.nfl * g.cc; acc.mb_size = a.mb_size; acc.u0 = a.u0; acc.u1 = a.u1;
  acc.gam0 = gam0; acc.gam1 = gam1; acc.beta_dt = beta_dt; acc.save = save ? 1 : 0;
  switch (recon) {
    case RECON_DC: return launch_fluxes_update_dc(rs, g, eos, a, acc, st);
    case RECON_PLM: return launch_fluxes_update_plm(rs, g, eos, a, acc, st);
    case RECON_PPM4: return launch_fluxes_update_ppm4(rs, g, eos, a, acc, st);
    case RECON_PPMX: return launch_fluxes_update_ppmx(rs, g, eos, a, acc, st);
    case RECON_WENOZ: return launch_fluxes_update_wenoz(rs, g, eos, a, acc, st);
    case RECON_TENO: return launch_fluxes_update_teno(rs, g, eos, a, acc, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace akb
