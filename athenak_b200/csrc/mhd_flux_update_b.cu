// mhd_flux_update_b.cu -- instantiations of the flux + update fusion sweeps (AccArgs in flux_kernels.cuh) for RECON_PPM4.
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
cudaError_t launch_fluxes_update_ppm4(int rs, const Geom &g, const EosParams &eos, const Arrays &a, const AccArgs &acc, cudaStream_t st) { return launch_fluxes_update_r<RECON_PPM4>(rs, g, eos, a, acc, st); }

}  // namespace akb
