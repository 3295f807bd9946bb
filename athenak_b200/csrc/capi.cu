// capi.cu -- extern "C" boundary declared in include/athenak_b200.h.
#include <cuda_runtime.h>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/athenak_b200.h"
#include "halo.h"
#include "mhd_device.cuh"
#include "mhd_kernels.h"

namespace {

struct Handle {
  ak_mhd_cfg cfg;
  akb::Geom g;
  akb::EosParams eos;
  bool hydro = false;                       // hydro twin (ak_hydro_*): no magnetic arrays, 1-D/2-D/3-D
  int rs = 0;                               // internal solver code (cfg.rsolver, + 20 when isothermal)
  int *d_counters = nullptr;                // 4 ints: dfloor, efloor, tfloor, fofc
  unsigned char *d_fofc = nullptr;          // FOFC cell flags, nmb*cc bytes (cfg.fofc)
  double stage_rk[3] = {0.0, 1.0, 0.0};     // gam0, gam1, beta*dt for the fused FOFC call (ak_mhd_set_stage)
  // ak_mhd_stage_begin: the U half of the stage update runs on an auxiliary stream next to the corner-EMF kernel
  int armed_stage = 0;                      // > 0: ak_mhd_stage_flux_emf launches the U update of this stage early
  bool u_pending = false;                   // the early U update is in flight: ak_mhd_stage_update does CT only
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_u = nullptr;
  unsigned long long *d_dtmin = nullptr;    // 3 ordered-bit doubles
  double *d_hst = nullptr;                  // history-sum scratch, allocated on first use
  std::string err;
};

thread_local std::string g_err;  // errors raised before a handle exists

int fail(Handle *h, const std::string &msg) {
  if (h) h->err = msg; else g_err = msg;
  return 1;
}
int cuda_fail(Handle *h, const char *where, cudaError_t e) {
  return fail(h, std::string(where) + ": " + cudaGetErrorString(e));
}

// Small device -> host reads (new-dt slots, floor / FOFC counters, history sums, the all-reduced dt) do NOT use cudaMemcpyAsync:
// a D2H copy of any size queues on the device's download copy engine behind whatever bulk download another stream has in
// flight.  Under ak_pipe_* (host_driver.cpp) that is the previous job's 1.2 GB state, 21 ms: every step's first dt read waited
// for it and steps and downloads ran strictly one after the other.  Instead a one-block kernel stores the words into mapped
// pinned host memory (posted PCIe writes from the SM, no copy engine); the values are handed over after the stream sync that
// every such read is followed by anyway.  One mirror per host thread (a handle is driven by one thread at a time).
__global__ void mirror_words_kernel(const unsigned *__restrict__ src, unsigned *__restrict__ dst, int n) {
  for (int q = threadIdx.x; q < n; q += blockDim.x) dst[q] = src[q];
  __threadfence_system();
}
struct Mirror {
  static constexpr size_t CAP = 8192;
  unsigned char *host = nullptr;
  size_t used = 0;
  struct Pend { void *dst; size_t off, bytes; };
  std::vector<Pend> pend;
  bool failed = false;
  ~Mirror() { if (host) cudaFreeHost(host); }
};
thread_local Mirror t_mirror;
cudaError_t small_d2h(void *dst, const void *src, size_t bytes, cudaStream_t st) {
  Mirror &m = t_mirror;
  if (!m.host && !m.failed) {
    if (cudaHostAlloc((void **)&m.host, Mirror::CAP, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { m.host = nullptr; m.failed = true; cudaGetLastError(); }
  }
  const size_t need = (bytes + 15) & ~(size_t)15;
  if (!m.host || (bytes & 3) || m.used + need > Mirror::CAP) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
  mirror_words_kernel<<<1, 128, 0, st>>>((const unsigned *)src, (unsigned *)(m.host + m.used), (int)(bytes / 4));
  m.pend.push_back({dst, m.used, bytes});
  m.used += need;
  return cudaGetLastError();
}
cudaError_t small_sync(cudaStream_t st) {
  const cudaError_t e = cudaStreamSynchronize(st);
  Mirror &m = t_mirror;
  for (const auto &p : m.pend) std::memcpy(p.dst, m.host + p.off, p.bytes);
  m.pend.clear();
  m.used = 0;
  return e;
}

// new-dt partial slots: identity = (double)FLT_MAX, the reference's initial value (mhd_newdt.cpp:40-42)
cudaError_t dt_slots_init(Handle *h, cudaStream_t st) {
  static unsigned long long init[3 * DT_SLOTS];
  static bool ready = false;
  if (!ready) {
    const double ident = (double)FLT_MAX;
    for (int q = 0; q < 3 * DT_SLOTS; ++q) std::memcpy(&init[q], &ident, 8);
    ready = true;
  }
  return cudaMemcpyAsync(h->d_dtmin, init, sizeof(init), cudaMemcpyHostToDevice, st);
}
cudaError_t dt_slots_fetch(Handle *h, double *dt3, cudaStream_t st) {
  double part[3 * DT_SLOTS];
  cudaError_t e = small_d2h(part, h->d_dtmin, sizeof(part), st);
  if (e != cudaSuccess) return e;
  if ((e = small_sync(st)) != cudaSuccess) return e;
  for (int c = 0; c < 3; ++c) {
    double v = part[c];
    for (int q = 1; q < DT_SLOTS; ++q) v = (part[3 * q + c] < v) ? part[3 * q + c] : v;
    dt3[c] = v;
  }
  return cudaSuccess;
}

akb::Arrays to_arrays(const ak_mhd_arrays *a) {
  akb::Arrays r;
  static_assert(sizeof(akb::Arrays) == sizeof(ak_mhd_arrays), "ak_mhd_arrays / akb::Arrays mismatch");
  std::memcpy(&r, a, sizeof(r));
  return r;
}

int check_cfg(const ak_mhd_cfg *c, bool hydro, std::string &why) {
  if (!c) { why = "null cfg"; return 1; }
  if (c->nmb < 1) { why = "nmb < 1"; return 1; }
  // mesh.cpp:70-79: 1-D (nx2 = nx3 = 1), 2-D (nx3 = 1) or 3-D
  if (c->nx1 < 2 || c->nx2 < 1 || c->nx3 < 1 || (c->nx2 == 1 && c->nx3 > 1)) { why = "MeshBlock must be 1-D (nx2=nx3=1), 2-D (nx3=1) or 3-D"; return 1; }
  if (hydro) {
    if (c->rsolver == AK_RSOLVER_HYD_ADVECT) { why = "rsolver advect needs time/evolution=kinematic (hydro.cpp:255-265), which is not on this path"; return 1; }
    if (c->rsolver < AK_RSOLVER_HYD_LLF || c->rsolver > AK_RSOLVER_HYD_ROE) { why = "unknown hydro Riemann solver"; return 1; }
  } else {
    if (c->rsolver < AK_RSOLVER_LLF || c->rsolver > AK_RSOLVER_HLLD) { why = "unknown Riemann solver"; return 1; }
  }
  if (c->recon < AK_RECON_DC || c->recon > AK_RECON_TENO) { why = "unknown reconstruction method"; return 1; }
  // reference src/mhd/mhd.cpp:217-228: 5-point stencils need nghost >= 3, PLM/DC need 2
  const int need = (c->recon >= AK_RECON_PPM4) ? 3 : 2;
  if (c->ng < need) { why = "nghost too small for the reconstruction stencil"; return 1; }
  if (c->eos != AK_EOS_IDEAL && c->eos != AK_EOS_ISOTHERMAL) { why = "unknown eos"; return 1; }
  if (c->eos == AK_EOS_ISOTHERMAL) {
    // hydro.cpp:252-262: hllc cannot be used with the isothermal EOS
    if (hydro && c->rsolver == AK_RSOLVER_HYD_HLLC) { why = "rsolver hllc cannot be used with isothermal EOS"; return 1; }
    if (!(c->iso_sound_speed > 0.0)) { why = "iso_sound_speed must be positive"; return 1; }
  } else if (!(c->gamma > 1.0)) { why = "gamma must exceed 1"; return 1; }
  if (c->fofc != 0 && c->fofc != 1) { why = "fofc must be 0 or 1"; return 1; }
  if (c->fofc) {
    if (hydro) { why = "hydro FOFC is not built: the reference's Hydro never allocates utest/fofc (hydro.cpp:38-39) and crashes"; return 1; }
    // mhd.cpp:209-235: FOFC needs one more ghost cell than the stencil
    if (c->ng < need + 1 && c->recon != AK_RECON_DC) { why = "FOFC requires one more ghost zone than the reconstruction stencil"; return 1; }
  }
  if (c->nscalars < 0 || c->nscalars > 64) { why = "nscalars must be in [0, 64]"; return 1; }
  return 0;
}

}  // namespace

extern "C" {

const char *ak_last_error(void *handle) {
  Handle *h = static_cast<Handle *>(handle);
  return h ? h->err.c_str() : g_err.c_str();
}

static int create_impl(const ak_mhd_cfg *cfg, bool hydro, void **handle) {
  std::string why;
  if (!handle) return fail(nullptr, "ak_mhd_create: null handle pointer");
  *handle = nullptr;
  if (check_cfg(cfg, hydro, why)) return fail(nullptr, "ak_mhd_create: " + why);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, std::string("ak_mhd_create: no CUDA device (this engine has no CPU path): ") + cudaGetErrorString(e));
  Handle *h = new Handle();
  h->cfg = *cfg;
  h->hydro = hydro;
  akb::Geom &g = h->g;
  g.nmb = cfg->nmb; g.ng = cfg->ng; g.nx1 = cfg->nx1; g.nx2 = cfg->nx2; g.nx3 = cfg->nx3;
  g.nfl = (cfg->eos == AK_EOS_ISOTHERMAL) ? 4 : 5;
  g.nvar = g.nfl + cfg->nscalars;
  h->rs = akb::internal_rsolver(cfg->rsolver, cfg->eos);
  g.multi_d = cfg->nx2 > 1; g.three_d = cfg->nx3 > 1;          // mesh.cpp:70-79
  g.n1 = cfg->nx1 + 2 * cfg->ng;
  g.n2 = g.multi_d ? cfg->nx2 + 2 * cfg->ng : 1;               // mhd.cpp:149-152 / hydro.cpp
  g.n3 = g.three_d ? cfg->nx3 + 2 * cfg->ng : 1;
  g.is = cfg->ng; g.ie = g.is + cfg->nx1 - 1;   // reference src/mesh/mesh.cpp:285-333
  g.js = g.multi_d ? cfg->ng : 0; g.je = g.js + cfg->nx2 - 1;
  g.ks = g.three_d ? cfg->ng : 0; g.ke = g.ks + cfg->nx3 - 1;
  g.next = cfg->fofc ? 1 : 0;
  g.ext1 = hydro ? 0 : 1; g.ext2 = (!hydro && g.multi_d) ? 1 : 0; g.ext3 = (!hydro && g.three_d) ? 1 : 0;
  g.cc = (long)g.n3 * g.n2 * g.n1;
  g.f1 = (long)g.n3 * g.n2 * (g.n1 + 1); g.f2 = (long)g.n3 * (g.n2 + 1) * g.n1; g.f3 = (long)(g.n3 + 1) * g.n2 * g.n1;
  g.e1 = (long)(g.n3 + 1) * (g.n2 + 1) * g.n1; g.e2 = (long)(g.n3 + 1) * g.n2 * (g.n1 + 1); g.e3 = (long)g.n3 * (g.n2 + 1) * (g.n1 + 1);
  h->eos = akb::EosParams{cfg->gamma, cfg->dfloor, cfg->pfloor, cfg->tfloor, cfg->sfloor, cfg->sigma_max,
                          1.0 / (cfg->gamma - 1.0), cfg->iso_sound_speed};
  if ((e = cudaMalloc(&h->d_counters, 4 * sizeof(int))) != cudaSuccess ||
      (cfg->fofc && (e = cudaMalloc(&h->d_fofc, (size_t)g.nmb * g.cc)) != cudaSuccess) ||
      (e = cudaMalloc(&h->d_dtmin, 3 * DT_SLOTS * sizeof(unsigned long long))) != cudaSuccess) {
    int rc = cuda_fail(nullptr, "ak_mhd_create", e);
    delete h;
    return rc;
  }
  if ((e = cudaMemset(h->d_counters, 0, 4 * sizeof(int))) != cudaSuccess) {
    int rc = cuda_fail(nullptr, "ak_mhd_create", e);
    delete h;
    return rc;
  }
  *handle = h;
  return 0;
}

int ak_mhd_create(const ak_mhd_cfg *cfg, void **handle) { return create_impl(cfg, false, handle); }
int ak_hydro_create(const ak_mhd_cfg *cfg, void **handle) { return create_impl(cfg, true, handle); }
int ak_hydro_destroy(void *handle) { return ak_mhd_destroy(handle); }

int ak_mhd_destroy(void *handle) {
  Handle *h = static_cast<Handle *>(handle);
  if (!h) return 0;
  cudaFree(h->d_counters);
  cudaFree(h->d_fofc);
  if (h->aux) cudaStreamDestroy(h->aux);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_u) cudaEventDestroy(h->ev_u);
  cudaFree(h->d_dtmin);
  cudaFree(h->d_hst);
  delete h;
  return 0;
}

#define AK_PROLOGUE(name)                                          \
  Handle *h = static_cast<Handle *>(handle);                       \
  if (!h) return fail(nullptr, name ": null handle");              \
  if (!a) return fail(h, name ": null arrays");                    \
  cudaStream_t st = static_cast<cudaStream_t>(s);                  \
  akb::Arrays A = to_arrays(a);                                    \
  cudaError_t e = cudaSuccess; (void)e; (void)st; (void)A;

int ak_mhd_copy_cons(void *handle, const ak_mhd_arrays *a, ak_stream_t s) {
  AK_PROLOGUE("ak_mhd_copy_cons")
  const akb::Geom &g = h->g;
  if ((e = akb::launch_copy(A.u0, A.u1, (long)g.nvar * g.nmb * g.cc, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_copy_cons", e);
  if ((e = akb::launch_copy(A.b0x1f, A.b1x1f, g.nmb * g.f1, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_copy_cons", e);
  if ((e = akb::launch_copy(A.b0x2f, A.b1x2f, g.nmb * g.f2, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_copy_cons", e);
  if ((e = akb::launch_copy(A.b0x3f, A.b1x3f, g.nmb * g.f3, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_copy_cons", e);
  return 0;
}

int ak_mhd_fluxes(void *handle, const ak_mhd_arrays *a, ak_stream_t s) {
  AK_PROLOGUE("ak_mhd_fluxes")
  if ((e = akb::launch_fluxes(h->cfg.recon, h->rs, h->g, h->eos, A, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_fluxes", e);
  return 0;
}

int ak_mhd_rk_update(void *handle, const ak_mhd_arrays *a, double gam0, double gam1, double beta_dt, ak_stream_t s) {
  AK_PROLOGUE("ak_mhd_rk_update")
  if ((e = akb::launch_update_ct(h->g, A, true, false, false, gam0, gam1, beta_dt, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_rk_update", e);
  return 0;
}

int ak_mhd_corner_e(void *handle, const ak_mhd_arrays *a, ak_stream_t s) {
  AK_PROLOGUE("ak_mhd_corner_e")
  if ((e = akb::launch_corner_e(h->g, A, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_corner_e", e);
  return 0;
}

int ak_mhd_ct(void *handle, const ak_mhd_arrays *a, double gam0, double gam1, double beta_dt, ak_stream_t s) {
  AK_PROLOGUE("ak_mhd_ct")
  if ((e = akb::launch_update_ct(h->g, A, false, true, false, gam0, gam1, beta_dt, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_ct", e);
  return 0;
}

static int c2p_impl(Handle *h, const akb::Arrays &A, int il, int iu, int jl, int ju, int kl, int ku, int *nfloor,
                    double *dt3, cudaStream_t st, const char *name) {
  const akb::Geom &g = h->g;
  if (il < 0 || jl < 0 || kl < 0 || iu >= g.n1 || ju >= g.n2 || ku >= g.n3 || il > iu || jl > ju || kl > ku)
    return fail(h, std::string(name) + ": cell box outside the array");
  cudaError_t e;
  // floor counters accumulate on the device across calls and are zeroed when they are read
  if (dt3) {
    if ((e = dt_slots_init(h, st)) != cudaSuccess) return cuda_fail(h, name, e);
  }
  e = h->hydro ? akb::launch_hydro_cons2prim(g, h->eos, A, il, iu, jl, ju, kl, ku, h->d_counters, dt3 != nullptr, h->d_dtmin, st)
               : akb::launch_cons2prim(g, h->eos, A, il, iu, jl, ju, kl, ku, h->d_counters, dt3 != nullptr, h->d_dtmin, st);
  if (e != cudaSuccess) return cuda_fail(h, name, e);
  if (nfloor) {
    if ((e = small_d2h(nfloor, h->d_counters, 3 * sizeof(int), st)) != cudaSuccess) return cuda_fail(h, name, e);
    if ((e = cudaMemsetAsync(h->d_counters, 0, 3 * sizeof(int), st)) != cudaSuccess) return cuda_fail(h, name, e);
  }
  if (dt3) {
    if ((e = dt_slots_fetch(h, dt3, st)) != cudaSuccess) return cuda_fail(h, name, e);
  } else if (nfloor) {
    if ((e = small_sync(st)) != cudaSuccess) return cuda_fail(h, name, e);
  }
  return 0;
}

int ak_mhd_cons2prim(void *handle, const ak_mhd_arrays *a, int il, int iu, int jl, int ju, int kl, int ku,
                     int *nfloor, ak_stream_t s) {
  AK_PROLOGUE("ak_mhd_cons2prim")
  return c2p_impl(h, A, il, iu, jl, ju, kl, ku, nfloor, nullptr, st, "ak_mhd_cons2prim");
}

int ak_mhd_cons2prim_newdt(void *handle, const ak_mhd_arrays *a, int *nfloor, double *dt3, ak_stream_t s) {
  AK_PROLOGUE("ak_mhd_cons2prim_newdt")
  if (!dt3) return fail(h, "ak_mhd_cons2prim_newdt: null dt3");
  return c2p_impl(h, A, 0, h->g.n1 - 1, 0, h->g.n2 - 1, 0, h->g.n3 - 1, nfloor, dt3, st, "ak_mhd_cons2prim_newdt");
}

int ak_mhd_new_dt(void *handle, const ak_mhd_arrays *a, double *dt3, ak_stream_t s) {
  AK_PROLOGUE("ak_mhd_new_dt")
  if (!dt3) return fail(h, "ak_mhd_new_dt: null dt3");
  if ((e = dt_slots_init(h, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_new_dt", e);
  if ((e = akb::launch_newdt(h->g, h->eos, A, h->d_dtmin, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_new_dt", e);
  if ((e = dt_slots_fetch(h, dt3, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_new_dt", e);
  return 0;
}

int ak_mhd_fofc(void *handle, const ak_mhd_arrays *a, double gam0, double gam1, double beta_dt, int *nfofc, ak_stream_t s) {
  AK_PROLOGUE("ak_mhd_fofc")
  if (!h->cfg.fofc || h->hydro) return fail(h, "ak_mhd_fofc: handle was not created with cfg.fofc = 1");
  if ((e = akb::launch_fofc(h->g, h->eos, A, gam0, gam1, beta_dt, h->d_fofc, h->d_counters + 3, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_fofc", e);
  if (nfofc) {
    if ((e = small_d2h(nfofc, h->d_counters + 3, sizeof(int), st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_fofc", e);
    if ((e = cudaMemsetAsync(h->d_counters + 3, 0, sizeof(int), st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_fofc", e);
    if ((e = small_sync(st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_fofc", e);
  }
  return 0;
}

int ak_mhd_set_stage(void *handle, double gam0, double gam1, double beta_dt) {
  Handle *h = static_cast<Handle *>(handle);
  if (!h) return fail(nullptr, "ak_mhd_set_stage: null handle");
  h->stage_rk[0] = gam0; h->stage_rk[1] = gam1; h->stage_rk[2] = beta_dt;
  return 0;
}

int ak_mhd_stage_begin(void *handle, int stage, double gam0, double gam1, double beta_dt) {
  Handle *h = static_cast<Handle *>(handle);
  if (!h || h->hydro) return fail(h, "ak_mhd_stage_begin: bad handle");
  if (stage < 1) return fail(h, "ak_mhd_stage_begin: stage is 1-based");
  h->stage_rk[0] = gam0; h->stage_rk[1] = gam1; h->stage_rk[2] = beta_dt;
  const char *env = getenv("AKB_OVERLAP_U");
  if (env && env[0] == '0') { h->armed_stage = 0; return 0; }
  if (!h->aux) {
    if (cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_u, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      h->armed_stage = 0;      // no auxiliary stream: the plain order still works
      return 0;
    }
  }
  h->armed_stage = stage;
  return 0;
}

/* the auxiliary stream ak_mhd_stage_begin / ak_mhd_stage_flux_emf run the early U update on (NULL when that overlap is off):
 * the host layer queues the ghost-zone push of u0 behind it, so the push overlaps the corner EMFs, SendE/RecvE and CT */
int ak_mhd_aux_stream(void *handle, ak_stream_t *aux) {
  Handle *h = static_cast<Handle *>(handle);
  if (!h || !aux) return fail(h, "ak_mhd_aux_stream: null argument");
  *aux = h->u_pending ? static_cast<ak_stream_t>(h->aux) : nullptr;
  return 0;
}

int ak_mhd_stage_flux_emf(void *handle, const ak_mhd_arrays *a, ak_stream_t s) {
  AK_PROLOGUE("ak_mhd_stage_flux_emf")
  if ((e = akb::launch_fluxes(h->cfg.recon, h->rs, h->g, h->eos, A, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_stage_flux_emf", e);
  if (h->cfg.fofc && (e = akb::launch_fofc(h->g, h->eos, A, h->stage_rk[0], h->stage_rk[1], h->stage_rk[2], h->d_fofc, h->d_counters + 3, st)) != cudaSuccess)
    return cuda_fail(h, "ak_mhd_stage_flux_emf", e);
  if (h->armed_stage > 0) {
    // The conserved-variable update needs only the fluxes, the corner EMFs only w0/bcc0 and the face EMFs: run
    // both at once (each alone leaves ~40% of the HBM bandwidth idle); ak_mhd_stage_update joins and does CT.
    const bool save = (h->armed_stage == 1) && !h->cfg.fofc;
    cudaEventRecord(h->ev_fork, st);
    cudaStreamWaitEvent(h->aux, h->ev_fork, 0);
    if ((e = akb::launch_update_ct(h->g, A, true, false, save, h->stage_rk[0], h->stage_rk[1], h->stage_rk[2], h->aux)) != cudaSuccess)
      return cuda_fail(h, "ak_mhd_stage_flux_emf", e);
    cudaEventRecord(h->ev_u, h->aux);
    h->u_pending = true;
    h->armed_stage = 0;
  }
  if ((e = akb::launch_corner_e(h->g, A, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_stage_flux_emf", e);
  return 0;
}

int ak_mhd_stage_update(void *handle, const ak_mhd_arrays *a, int stage, double gam0, double gam1, double beta_dt,
                        ak_stream_t s) {
  AK_PROLOGUE("ak_mhd_stage_update")
  // stage 1: CopyCons is folded into the update kernel (it saves u1/b1 for the ACTIVE zones it
  // updates -- the only entries later stages read -- instead of copying whole arrays)
  // With FOFC the estimate reads u1/b1 in the ghost zones too, so the caller ran the full CopyCons.
  const bool save = stage == 1 && !h->cfg.fofc;
  if (h->u_pending) {      // U was updated next to the corner-EMF kernel (ak_mhd_stage_begin): CT only, after the join
    h->u_pending = false;
    cudaStreamWaitEvent(st, h->ev_u, 0);
    if ((e = akb::launch_update_ct(h->g, A, false, true, save, gam0, gam1, beta_dt, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_stage_update", e);
    return 0;
  }
  if ((e = akb::launch_update_ct(h->g, A, true, true, save, gam0, gam1, beta_dt, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_stage_update", e);
  return 0;
}

int ak_mhd_stage_fused(void *handle, const ak_mhd_arrays *a, int stage, double gam0, double gam1, double beta_dt,
                       void *halo_plan, ak_stream_t s) {
  int rc = 0;
  Handle *hh = static_cast<Handle *>(handle);
  if (hh && hh->cfg.fofc) {       // CopyCons over the whole arrays at stage 1, RK coefficients for the FOFC estimate
    if (stage == 1 && (rc = ak_mhd_copy_cons(handle, a, s)) != 0) return rc;
    ak_mhd_set_stage(handle, gam0, gam1, beta_dt);
  }
  if ((rc = ak_mhd_stage_flux_emf(handle, a, s)) != 0) return rc;
  if (halo_plan) {
    if (ak_halo_npeers(halo_plan) != 0)
      return fail(static_cast<Handle *>(handle), "ak_mhd_stage_fused: MeshBlocks on other ranks; use ak_mhd_stage_flux_emf + "
                                                 "ak_halo_emf_* + ak_mhd_stage_update");
    if ((rc = ak_halo_emf_pack_local(halo_plan, a->e1, a->e2, a->e3, s)) != 0) return rc;
    if ((rc = ak_halo_emf_sumavg(halo_plan, a->e1, a->e2, a->e3, s)) != 0) return rc;
  }
  return ak_mhd_stage_update(handle, a, stage, gam0, gam1, beta_dt, s);
}

// ---------------------------------------------------------------------------------------
// Two-kernel stage: flux sweeps (ak_mhd_fluxes) + ONE tail kernel for tasks 14, 1, 5, 17, 27 (and 28 on
// the last stage), plus the two shell kernels for the outermost layer around SendE/RecvE and the ghost exchange.
// Flux + update fusion: tasks 2 + 1 + 5 (Fluxes, CopyCons of u, RKUpdate) as the three sweeps alone -- the flux arrays are never
// written (only their IDN component, for the corner EMFs) and no update kernel runs; du = scratch, 2 * nmb * nfl * cells doubles.
int ak_mhd_fluxes_update_supported(void *handle) {
  Handle *h = static_cast<Handle *>(handle);
  return (h && !h->hydro && !h->cfg.fofc && akb::fluxes_update_supported(h->g)) ? 1 : 0;
}
int ak_mhd_fluxes_update(void *handle, const ak_mhd_arrays *a, double *du, int stage, double gam0, double gam1, double beta_dt,
                         ak_stream_t s) {
  AK_PROLOGUE("ak_mhd_fluxes_update")
  if (!ak_mhd_fluxes_update_supported(handle)) return fail(h, "ak_mhd_fluxes_update: needs 3-D MHD without passive scalars or FOFC");
  if (!du) return fail(h, "ak_mhd_fluxes_update: null scratch array");
  if (stage < 1) return fail(h, "ak_mhd_fluxes_update: stage is 1-based");
  if ((e = akb::launch_fluxes_update(h->cfg.recon, h->rs, h->g, h->eos, A, du, stage == 1, gam0, gam1, beta_dt, st)) != cudaSuccess)
    return cuda_fail(h, "ak_mhd_fluxes_update", e);
  return 0;
}

// task 17 with the stage-1 CopyCons of the face field folded in (b1 <- old b0 on the faces it updates): the CT half of
// ak_mhd_stage_update for callers whose conserved update already happened (ak_mhd_fluxes_update)
int ak_mhd_ct_save(void *handle, const ak_mhd_arrays *a, int stage, double gam0, double gam1, double beta_dt, ak_stream_t s) {
  AK_PROLOGUE("ak_mhd_ct_save")
  if (h->hydro) return fail(h, "ak_mhd_ct_save: MHD handles only");
  if ((e = akb::launch_update_ct(h->g, A, false, true, stage == 1 && !h->cfg.fofc, gam0, gam1, beta_dt, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_ct_save", e);
  return 0;
}

static akb::TailArrays tail_arrays(const akb::Arrays &in, const akb::Arrays &out, bool stage1) {
  akb::TailArrays T;
  T.u_in = in.u0; T.u1r = stage1 ? nullptr : in.u1; T.u_out = out.u0; T.u1w = stage1 ? in.u1 : nullptr;
  T.b_in[0] = in.b0x1f; T.b_in[1] = in.b0x2f; T.b_in[2] = in.b0x3f;
  T.b1[0] = stage1 ? nullptr : in.b1x1f; T.b1[1] = stage1 ? nullptr : in.b1x2f; T.b1[2] = stage1 ? nullptr : in.b1x3f;
  T.b_out[0] = out.b0x1f; T.b_out[1] = out.b0x2f; T.b_out[2] = out.b0x3f;
  T.w_in = in.w0; T.bcc_in = in.bcc0; T.w_out = out.w0; T.bcc_out = out.bcc0;
  T.flx[0] = in.flx1; T.flx[1] = in.flx2; T.flx[2] = in.flx3;
  T.e3x1 = in.e3x1; T.e2x1 = in.e2x1; T.e1x2 = in.e1x2; T.e3x2 = in.e3x2; T.e2x3 = in.e2x3; T.e1x3 = in.e1x3;
  T.e1 = in.e1; T.e2 = in.e2; T.e3 = in.e3;
  T.mb_size = in.mb_size;
  return T;
}
static int tail_check(Handle *h, const ak_mhd_arrays *a, const ak_mhd_arrays *out, const char *name) {
  if (h->hydro) return fail(h, std::string(name) + ": MHD handles only");
  if (!akb::stage_tail_supported(h->g)) return fail(h, std::string(name) + ": needs 3-D blocks of at least 3 cells per direction, nghost >= 2, no FOFC");
  if (!out) return fail(h, std::string(name) + ": null output arrays");
  if (out->b0x1f == a->b0x1f || out->b0x2f == a->b0x2f || out->b0x3f == a->b0x3f || out->w0 == a->w0 || out->bcc0 == a->bcc0)
    return fail(h, std::string(name) + ": the face field and the primitives are written out of place (out->b0x?f, out->w0, out->bcc0 must be other buffers)");
  return 0;
}

int ak_mhd_tail_supported(void *handle) {
  Handle *h = static_cast<Handle *>(handle);
  return (h && !h->hydro && akb::stage_tail_supported(h->g)) ? 1 : 0;
}

int ak_mhd_stage_tail(void *handle, const ak_mhd_arrays *a, const ak_mhd_arrays *out, int stage, double gam0, double gam1,
                      double beta_dt, int want_dt, ak_stream_t s) {
  AK_PROLOGUE("ak_mhd_stage_tail")
  if (tail_check(h, a, out, "ak_mhd_stage_tail")) return 1;
  if (stage < 1) return fail(h, "ak_mhd_stage_tail: stage is 1-based");
  if (want_dt && (e = dt_slots_init(h, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_stage_tail", e);
  const akb::TailArrays T = tail_arrays(A, to_arrays(out), stage == 1);
  // AKB_TAIL_SPLIT=1: field kernel (corner EMFs + CT) + cell kernel (update + c2p) instead of the single tail kernel
  static const bool split = [] { const char *v = getenv("AKB_TAIL_SPLIT"); return v && v[0] == '1'; }();
  if ((e = akb::launch_stage_tail(h->g, h->eos, T, stage == 1, gam0, gam1, beta_dt, h->d_counters, want_dt != 0, h->d_dtmin, split, st)) != cudaSuccess)
    return cuda_fail(h, "ak_mhd_stage_tail", e);
  return 0;
}

int ak_mhd_stage_shell_ct(void *handle, const ak_mhd_arrays *a, const ak_mhd_arrays *out, int stage, double gam0, double gam1,
                          double beta_dt, ak_stream_t s) {
  AK_PROLOGUE("ak_mhd_stage_shell_ct")
  if (tail_check(h, a, out, "ak_mhd_stage_shell_ct")) return 1;
  const akb::TailArrays T = tail_arrays(A, to_arrays(out), stage == 1);
  if ((e = akb::launch_shell_ct(h->g, T, stage == 1, gam0, gam1, beta_dt, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_stage_shell_ct", e);
  return 0;
}

int ak_mhd_stage_shell_c2p(void *handle, const ak_mhd_arrays *a, int *nfloor, double *dt3, ak_stream_t s) {
  AK_PROLOGUE("ak_mhd_stage_shell_c2p")
  if (h->hydro || !akb::stage_tail_supported(h->g)) return fail(h, "ak_mhd_stage_shell_c2p: not available for this handle (see ak_mhd_tail_supported)");
  if ((e = akb::launch_shell_c2p(h->g, h->eos, A, h->d_counters, dt3 != nullptr, h->d_dtmin, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_stage_shell_c2p", e);
  if (nfloor) {
    if ((e = small_d2h(nfloor, h->d_counters, 3 * sizeof(int), st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_stage_shell_c2p", e);
    if ((e = cudaMemsetAsync(h->d_counters, 0, 3 * sizeof(int), st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_stage_shell_c2p", e);
  }
  if (dt3) {
    if ((e = dt_slots_fetch(h, dt3, st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_stage_shell_c2p", e);
  } else if (nfloor) {
    if ((e = small_sync(st)) != cudaSuccess) return cuda_fail(h, "ak_mhd_stage_shell_c2p", e);
  }
  return 0;
}

// history sums (src/outputs/history.cpp:78-160,272-372); synchronises the stream
int ak_hst_sums(void *handle, const ak_mhd_arrays *a, double *sums, int *nhist, ak_stream_t s) {
  AK_PROLOGUE("ak_hst_sums")
  if (!sums) return fail(h, "ak_hst_sums: null output");
  const int nb = 148 * 4;
  if (!h->d_hst && (e = cudaMalloc(&h->d_hst, sizeof(double) * 11 * (nb + 1))) != cudaSuccess) return cuda_fail(h, "ak_hst_sums", e);
  if ((e = akb::launch_hst_sums(h->g, A, !h->hydro, h->d_hst, nb, st)) != cudaSuccess) return cuda_fail(h, "ak_hst_sums", e);
  const int n = (h->hydro ? 8 : 11) - (5 - h->g.nfl);   // isothermal: no tot-E column (history.cpp:84-88,278-282)
  if ((e = small_d2h(sums, h->d_hst + 11L * nb, sizeof(double) * n, st)) != cudaSuccess) return cuda_fail(h, "ak_hst_sums", e);
  if ((e = small_sync(st)) != cudaSuccess) return cuda_fail(h, "ak_hst_sums", e);
  if (nhist) *nhist = n;
  return 0;
}

// ---------------------------------------------------------------------------------------
// hydro twin (reference src/hydro/hydro_tasks.cpp:47-79)
#define AK_HYD_PROLOGUE(name)                                                   \
  AK_PROLOGUE(name)                                                             \
  if (!h->hydro) return fail(h, name ": handle was not made by ak_hydro_create");

int ak_hydro_copy_cons(void *handle, const ak_mhd_arrays *a, ak_stream_t s) {
  AK_HYD_PROLOGUE("ak_hydro_copy_cons")
  if ((e = akb::launch_copy(A.u0, A.u1, (long)h->g.nvar * h->g.nmb * h->g.cc, st)) != cudaSuccess) return cuda_fail(h, "ak_hydro_copy_cons", e);
  return 0;
}
int ak_hydro_fluxes(void *handle, const ak_mhd_arrays *a, ak_stream_t s) {
  AK_HYD_PROLOGUE("ak_hydro_fluxes")
  if ((e = akb::launch_hydro_fluxes(h->cfg.recon, h->rs, h->g, h->eos, A, st)) != cudaSuccess) return cuda_fail(h, "ak_hydro_fluxes", e);
  return 0;
}
int ak_hydro_rk_update(void *handle, const ak_mhd_arrays *a, double gam0, double gam1, double beta_dt, ak_stream_t s) {
  AK_HYD_PROLOGUE("ak_hydro_rk_update")
  if ((e = akb::launch_update_ct(h->g, A, true, false, false, gam0, gam1, beta_dt, st)) != cudaSuccess) return cuda_fail(h, "ak_hydro_rk_update", e);
  return 0;
}
int ak_hydro_stage_fused(void *handle, const ak_mhd_arrays *a, int stage, double gam0, double gam1, double beta_dt, ak_stream_t s) {
  AK_HYD_PROLOGUE("ak_hydro_stage_fused")
  if ((e = akb::launch_hydro_fluxes(h->cfg.recon, h->rs, h->g, h->eos, A, st)) != cudaSuccess) return cuda_fail(h, "ak_hydro_stage_fused", e);
  if ((e = akb::launch_update_ct(h->g, A, true, false, stage == 1, gam0, gam1, beta_dt, st)) != cudaSuccess) return cuda_fail(h, "ak_hydro_stage_fused", e);
  return 0;
}
int ak_hydro_cons2prim(void *handle, const ak_mhd_arrays *a, int il, int iu, int jl, int ju, int kl, int ku,
                       int *nfloor, ak_stream_t s) {
  AK_HYD_PROLOGUE("ak_hydro_cons2prim")
  return c2p_impl(h, A, il, iu, jl, ju, kl, ku, nfloor, nullptr, st, "ak_hydro_cons2prim");
}
int ak_hydro_cons2prim_newdt(void *handle, const ak_mhd_arrays *a, int *nfloor, double *dt3, ak_stream_t s) {
  AK_HYD_PROLOGUE("ak_hydro_cons2prim_newdt")
  if (!dt3) return fail(h, "ak_hydro_cons2prim_newdt: null dt3");
  return c2p_impl(h, A, 0, h->g.n1 - 1, 0, h->g.n2 - 1, 0, h->g.n3 - 1, nfloor, dt3, st, "ak_hydro_cons2prim_newdt");
}
int ak_hydro_new_dt(void *handle, const ak_mhd_arrays *a, double *dt3, ak_stream_t s) {
  AK_HYD_PROLOGUE("ak_hydro_new_dt")
  if (!dt3) return fail(h, "ak_hydro_new_dt: null dt3");
  if ((e = dt_slots_init(h, st)) != cudaSuccess) return cuda_fail(h, "ak_hydro_new_dt", e);
  if ((e = akb::launch_hydro_newdt(h->g, h->eos, A, h->d_dtmin, st)) != cudaSuccess) return cuda_fail(h, "ak_hydro_new_dt", e);
  if ((e = dt_slots_fetch(h, dt3, st)) != cudaSuccess) return cuda_fail(h, "ak_hydro_new_dt", e);
  return 0;
}

// ---------------------------------------------------------------------------------------
// halo

int ak_halo_create(const ak_mhd_cfg *cfg, const int *nghbr, int my_rank, int gids_start, void **plan) {
  std::string why;
  if (!plan) return fail(nullptr, "ak_halo_create: null plan pointer");
  *plan = nullptr;
  if (check_cfg(cfg, cfg && cfg->rsolver >= AK_RSOLVER_HYD_ADVECT, why)) return fail(nullptr, "ak_halo_create: " + why);
  if (!nghbr) return fail(nullptr, "ak_halo_create: null neighbour table");
  akb::HaloPlan *P = akb::halo_build(*cfg, nghbr, my_rank, gids_start, why);
  if (!P) return fail(nullptr, "ak_halo_create: " + why);
  *plan = P;
  return 0;
}

// Host-side introspection of the resolved copy plan (no device needed): used by the CPU tests
// of the partition/exchange logic and by hosts that want to drive their own transport.
// kind: 0 local_cc, 1 local_fc, 2 pack_cc, 3 unpack_cc, 4 pack_fc, 5 unpack_fc (ipeer ignored
// for 0,1).  out: [max][13] longs = m, src_m, comp, lo[3], hi[3], shift[3], boff.
int ak_halo_descs(void *plan, int kind, int ipeer, long *out, int max) {
  akb::HaloPlan *P = static_cast<akb::HaloPlan *>(plan);
  if (!P) return -1;
  const std::vector<akb::HaloDesc> *v = nullptr;
  if (kind == 0) v = &P->local_cc;
  else if (kind == 1) v = &P->local_fc;
  else {
    if (ipeer < 0 || ipeer >= (int)P->peers.size()) return -1;
    akb::HaloPeer &p = P->peers[ipeer];
    v = kind == 2 ? &p.pack_cc : kind == 3 ? &p.unpack_cc : kind == 4 ? &p.pack_fc : kind == 5 ? &p.unpack_fc : nullptr;
  }
  if (!v) return -1;
  const int n = (int)v->size();
  for (int q = 0; q < n && q < max && out; ++q) {
    const akb::HaloDesc &d = (*v)[q];
    long *o = out + (long)q * 13;
    o[0] = d.m; o[1] = d.src_m; o[2] = d.comp;
    for (int a = 0; a < 3; ++a) { o[3 + a] = d.lo[a]; o[6 + a] = d.hi[a]; o[9 + a] = d.shift[a]; }
    o[12] = d.boff;
  }
  return n;
}

int ak_halo_destroy(void *plan) {
  akb::halo_free(static_cast<akb::HaloPlan *>(plan));
  return 0;
}

int ak_halo_npeers(void *plan) { return plan ? (int)static_cast<akb::HaloPlan *>(plan)->peers.size() : 0; }

int ak_halo_peer(void *plan, int ipeer, int *rank, long *cc_count_per_var, long *fc_count) {
  akb::HaloPlan *P = static_cast<akb::HaloPlan *>(plan);
  if (!P || ipeer < 0 || ipeer >= (int)P->peers.size()) return fail(nullptr, "ak_halo_peer: bad peer index");
  if (rank) *rank = P->peers[ipeer].rank;
  if (cc_count_per_var) *cc_count_per_var = P->peers[ipeer].cc_count;
  if (fc_count) *fc_count = P->peers[ipeer].fc_count;
  return 0;
}

static int halo_ready(akb::HaloPlan *P, const char *where) {
  cudaError_t e = akb::halo_ensure_uploaded(P);
  if (e == cudaSuccess) return 0;
  P->err = std::string(where) + ": " + cudaGetErrorString(e);
  g_err = P->err;
  return 1;
}

static int halo_rc(akb::HaloPlan *P, const char *where, cudaError_t e) {
  if (e == cudaSuccess) return 0;
  P->err = std::string(where) + ": " + cudaGetErrorString(e);
  g_err = P->err;
  return 1;
}

int ak_halo_local_cc(void *plan, double *a, int nvar, ak_stream_t s) {
  akb::HaloPlan *P = static_cast<akb::HaloPlan *>(plan);
  if (!P || !a) return fail(nullptr, "ak_halo_local_cc: null argument");
  if (halo_ready(P, "ak_halo_local_cc")) return 1;
  akb::HaloArrays A = akb::halo_arrays_cc(P->cfg, a, nvar);
  const long cc = A.bstride[0] / nvar;
  return halo_rc(P, "ak_halo_local_cc", akb::halo_launch(P, P->local_cc, P->d_local_cc, A, nvar, cc, nullptr, akb::HALO_LOCAL, (cudaStream_t)s));
}

int ak_halo_local_fc(void *plan, double *x1f, double *x2f, double *x3f, ak_stream_t s) {
  akb::HaloPlan *P = static_cast<akb::HaloPlan *>(plan);
  if (!P || !x1f || !x2f || !x3f) return fail(nullptr, "ak_halo_local_fc: null argument");
  if (halo_ready(P, "ak_halo_local_fc")) return 1;
  akb::HaloArrays A = akb::halo_arrays_fc(P->cfg, x1f, x2f, x3f);
  return halo_rc(P, "ak_halo_local_fc", akb::halo_launch(P, P->local_fc, P->d_local_fc, A, 1, 0, nullptr, akb::HALO_LOCAL, (cudaStream_t)s));
}

#define AK_PEER(name)                                                                         \
  akb::HaloPlan *P = static_cast<akb::HaloPlan *>(plan);                                      \
  if (!P || ipeer < 0 || ipeer >= (int)P->peers.size()) return fail(nullptr, name ": bad plan or peer index"); \
  if (halo_ready(P, name)) return 1;                                                          \
  akb::HaloPeer &peer = P->peers[ipeer];

int ak_halo_pack_cc(void *plan, const double *a, int nvar, int ipeer, double *sendbuf, ak_stream_t s) {
  AK_PEER("ak_halo_pack_cc")
  akb::HaloArrays A = akb::halo_arrays_cc(P->cfg, const_cast<double *>(a), nvar);
  return halo_rc(P, "ak_halo_pack_cc", akb::halo_launch(P, peer.pack_cc, peer.d_pack_cc, A, nvar, A.bstride[0] / nvar, sendbuf, akb::HALO_PACK, (cudaStream_t)s));
}
int ak_halo_unpack_cc(void *plan, double *a, int nvar, int ipeer, const double *recvbuf, ak_stream_t s) {
  AK_PEER("ak_halo_unpack_cc")
  akb::HaloArrays A = akb::halo_arrays_cc(P->cfg, a, nvar);
  return halo_rc(P, "ak_halo_unpack_cc", akb::halo_launch(P, peer.unpack_cc, peer.d_unpack_cc, A, nvar, A.bstride[0] / nvar, const_cast<double *>(recvbuf), akb::HALO_UNPACK, (cudaStream_t)s));
}
int ak_halo_pack_fc(void *plan, const double *x1f, const double *x2f, const double *x3f, int ipeer, double *sendbuf, ak_stream_t s) {
  AK_PEER("ak_halo_pack_fc")
  akb::HaloArrays A = akb::halo_arrays_fc(P->cfg, const_cast<double *>(x1f), const_cast<double *>(x2f), const_cast<double *>(x3f));
  return halo_rc(P, "ak_halo_pack_fc", akb::halo_launch(P, peer.pack_fc, peer.d_pack_fc, A, 1, 0, sendbuf, akb::HALO_PACK, (cudaStream_t)s));
}
int ak_halo_unpack_fc(void *plan, double *x1f, double *x2f, double *x3f, int ipeer, const double *recvbuf, ak_stream_t s) {
  AK_PEER("ak_halo_unpack_fc")
  akb::HaloArrays A = akb::halo_arrays_fc(P->cfg, x1f, x2f, x3f);
  return halo_rc(P, "ak_halo_unpack_fc", akb::halo_launch(P, peer.unpack_fc, peer.d_unpack_fc, A, 1, 0, const_cast<double *>(recvbuf), akb::HALO_UNPACK, (cudaStream_t)s));
}

// ---- SendE/RecvE: EMF flux correction on a same-level mesh (flux_correct_fc.cpp) ------------
int ak_halo_emf_peer(void *plan, int ipeer, long *count, double **recv_ptr) {
  AK_PEER("ak_halo_emf_peer")
  if (count) *count = peer.emf_count;
  if (recv_ptr) *recv_ptr = P->d_emf_rbuf + peer.emf_recv_base;
  return 0;
}
int ak_halo_emf_rbuf(void *plan, double **base, long *count) {
  akb::HaloPlan *P = static_cast<akb::HaloPlan *>(plan);
  if (!P) return fail(nullptr, "ak_halo_emf_rbuf: null plan");
  if (halo_ready(P, "ak_halo_emf_rbuf")) return 1;
  if (base) *base = P->d_emf_rbuf;
  if (count) *count = P->emf_rbuf_count;
  return 0;
}
int ak_halo_emf_pack_local(void *plan, const double *e1, const double *e2, const double *e3, ak_stream_t s) {
  akb::HaloPlan *P = static_cast<akb::HaloPlan *>(plan);
  if (!P || !e1 || !e2 || !e3) return fail(nullptr, "ak_halo_emf_pack_local: null argument");
  if (halo_ready(P, "ak_halo_emf_pack_local")) return 1;
  akb::HaloArrays A = akb::halo_arrays_ec(P->cfg, const_cast<double *>(e1), const_cast<double *>(e2), const_cast<double *>(e3));
  return halo_rc(P, "ak_halo_emf_pack_local", akb::halo_launch(P, P->emf_pack_local, P->d_emf_pack_local, A, 1, 0, P->d_emf_rbuf, akb::HALO_PACK, (cudaStream_t)s));
}
int ak_halo_emf_pack(void *plan, const double *e1, const double *e2, const double *e3, int ipeer, double *sendbuf, ak_stream_t s) {
  AK_PEER("ak_halo_emf_pack")
  akb::HaloArrays A = akb::halo_arrays_ec(P->cfg, const_cast<double *>(e1), const_cast<double *>(e2), const_cast<double *>(e3));
  return halo_rc(P, "ak_halo_emf_pack", akb::halo_launch(P, peer.pack_emf, peer.d_pack_emf, A, 1, 0, sendbuf, akb::HALO_PACK, (cudaStream_t)s));
}
int ak_halo_emf_sumavg(void *plan, double *e1, double *e2, double *e3, ak_stream_t s) {
  akb::HaloPlan *P = static_cast<akb::HaloPlan *>(plan);
  if (!P || !e1 || !e2 || !e3) return fail(nullptr, "ak_halo_emf_sumavg: null argument");
  if (halo_ready(P, "ak_halo_emf_sumavg")) return 1;
  akb::HaloArrays A = akb::halo_arrays_ec(P->cfg, e1, e2, e3);
  return halo_rc(P, "ak_halo_emf_sumavg", akb::emf_sumavg_launch(P, A, (cudaStream_t)s));
}

// ---- peer-memory ghost exchange ---------------------------------------------------------------
int ak_halo_push_setup(void *plan, int ipeer, int peer_gids_start, double *pu0, double *px1f, double *px2f, double *px3f) {
  akb::HaloPlan *P = static_cast<akb::HaloPlan *>(plan);
  if (!P || ipeer < 0 || ipeer >= (int)P->peers.size() || ipeer >= AKB_MAX_PUSH_PEERS) return fail(nullptr, "ak_halo_push_setup: bad peer");
  if (!pu0) return fail(nullptr, "ak_halo_push_setup: null peer array");
  P->push_ptrs.p[ipeer][0] = pu0; P->push_ptrs.p[ipeer][1] = px1f; P->push_ptrs.p[ipeer][2] = px2f; P->push_ptrs.p[ipeer][3] = px3f;
  for (auto &d : P->push)
    if (d.peer == ipeer) {
      d.dst_m = d.dst_gid - peer_gids_start;
      if (d.dst_m < 0) return fail(nullptr, "ak_halo_push_setup: neighbour gid below the peer's first gid");
      if (d.comp >= 0 && !P->push_ptrs.p[ipeer][1 + d.comp]) return fail(nullptr, "ak_halo_push_setup: null peer face array");
    }
  return 0;
}

// the peer's face-field arrays of the next push (the two-kernel stage rotates its face-field buffers; every rank does so at
// the same point of the task list, so the sender follows the receiver's rotation with its own copy of the mapped pointers)
int ak_halo_push_set_faces(void *plan, int ipeer, double *px1f, double *px2f, double *px3f) {
  akb::HaloPlan *P = static_cast<akb::HaloPlan *>(plan);
  if (!P || ipeer < 0 || ipeer >= (int)P->peers.size() || ipeer >= AKB_MAX_PUSH_PEERS) return fail(nullptr, "ak_halo_push_set_faces: bad peer");
  if (!px1f || !px2f || !px3f) return fail(nullptr, "ak_halo_push_set_faces: null peer face array");
  P->push_ptrs.p[ipeer][1] = px1f; P->push_ptrs.p[ipeer][2] = px2f; P->push_ptrs.p[ipeer][3] = px3f;
  return 0;
}

int ak_halo_push_flags(void *plan, int my_rank, int nranks, unsigned long long *local_flags,
                       unsigned long long *const *peer_flags) {
  akb::HaloPlan *P = static_cast<akb::HaloPlan *>(plan);
  if (!P || !local_flags || !peer_flags) return fail(nullptr, "ak_halo_push_flags: null argument");
  const int np = (int)P->peers.size();
  // kinds: 0/1 = ghost zones free / filled, 2/3 = EMF receive segment free / filled
  std::vector<unsigned long long *> sg(np);
  std::vector<int> ws(4 * np);
  cudaError_t e;
  if ((e = cudaMalloc(&P->d_wait_slots, sizeof(int) * std::max(4 * np, 1))) != cudaSuccess) return cuda_fail(nullptr, "ak_halo_push_flags", e);
  for (int kind = 0; kind < 4; ++kind) {
    for (int p = 0; p < np; ++p) {
      sg[p] = peer_flags[p] + kind * nranks + my_rank;       // my slot of this kind in peer p's memory
      ws[kind * np + p] = kind * nranks + P->peers[p].rank;  // where peer p writes its counter in MY memory
    }
    if ((e = cudaMalloc(&P->d_sig[kind], sizeof(void *) * std::max(np, 1))) != cudaSuccess) return cuda_fail(nullptr, "ak_halo_push_flags", e);
    if (np) cudaMemcpy(P->d_sig[kind], sg.data(), sizeof(void *) * np, cudaMemcpyHostToDevice);
  }
  if (np) cudaMemcpy(P->d_wait_slots, ws.data(), sizeof(int) * 4 * np, cudaMemcpyHostToDevice);
  P->flags_local = local_flags; P->my_rank = my_rank; P->nranks = nranks;
  return 0;
}

int ak_halo_push_commit(void *plan) {
  akb::HaloPlan *P = static_cast<akb::HaloPlan *>(plan);
  if (!P) return fail(nullptr, "ak_halo_push_commit: null plan");
  for (const auto &d : P->push) if (d.dst_m < 0) return fail(nullptr, "ak_halo_push_commit: a peer was not set up");
  if (!P->flags_local) return fail(nullptr, "ak_halo_push_commit: flags were not set up");
  cudaFree(P->d_push); P->d_push = nullptr;
  for (auto &kv : P->tiles) cudaFree(kv.second.d);     // launch tables are keyed by the descriptor arrays: rebuild
  P->tiles.clear();
  if (!P->push.empty()) {
    cudaError_t e = cudaMalloc(&P->d_push, sizeof(akb::PushDesc) * P->push.size());
    if (e == cudaSuccess) e = cudaMemcpy(P->d_push, P->push.data(), sizeof(akb::PushDesc) * P->push.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(nullptr, "ak_halo_push_commit", e);
  }
  P->push_ready = true;
  return 0;
}

int ak_halo_push(void *plan, double *u0, int nvar, double *x1f, double *x2f, double *x3f, ak_stream_t s) {
  akb::HaloPlan *P = static_cast<akb::HaloPlan *>(plan);
  if (!P || !P->push_ready) return fail(nullptr, "ak_halo_push: plan has no committed peer-memory set-up");
  const bool do_cc = u0 != nullptr, do_fc = x1f != nullptr;
  if (do_fc && (!x2f || !x3f)) return fail(nullptr, "ak_halo_push: null face array");
  akb::HaloArrays CCa = akb::halo_arrays_cc(P->cfg, u0, nvar > 0 ? nvar : 1);
  akb::HaloArrays FCa = akb::halo_arrays_fc(P->cfg, x1f, x2f, x3f);
  const long cc = CCa.bstride[0] / (nvar > 0 ? nvar : 1);
  return halo_rc(P, "ak_halo_push", akb::halo_push_launch(P, CCa, nvar, cc, FCa, do_cc, do_fc, (cudaStream_t)s));
}

int ak_halo_signal(void *plan, int kind, unsigned long long value, ak_stream_t s) {
  akb::HaloPlan *P = static_cast<akb::HaloPlan *>(plan);
  if (!P || !P->push_ready || kind < 0 || kind > 3) return fail(nullptr, "ak_halo_signal: bad argument");
  return halo_rc(P, "ak_halo_signal", akb::peer_signal_launch(P->d_sig[kind], (int)P->peers.size(), value, (cudaStream_t)s));
}

int ak_halo_wait(void *plan, int kind, unsigned long long value, ak_stream_t s) {
  akb::HaloPlan *P = static_cast<akb::HaloPlan *>(plan);
  if (!P || !P->push_ready || kind < 0 || kind > 3) return fail(nullptr, "ak_halo_wait: bad argument");
  const int np = (int)P->peers.size();
  return halo_rc(P, "ak_halo_wait", akb::peer_wait_launch(P->flags_local, P->d_wait_slots + kind * np, np, value, (cudaStream_t)s));
}

/* Synchronous read of a few words of device memory into host memory WITHOUT the download copy engine (see small_d2h above):
 * the host driver's all-reduced dt.  bytes must be a multiple of 4. */
int ak_small_d2h_sync(void *dst_host, const void *src_dev, long bytes, ak_stream_t s) {
  cudaError_t e = small_d2h(dst_host, src_dev, (size_t)bytes, (cudaStream_t)s);
  if (e == cudaSuccess) e = small_sync((cudaStream_t)s);
  return e == cudaSuccess ? 0 : cuda_fail(nullptr, "ak_small_d2h_sync", e);
}

/* x2 / x3 flux sweeps of the PPM4 path through the TMA form (flux_march_tma_kernel: ring rows fetched by cp.async.bulk, one issuing
 * thread per CTA) instead of per-thread cp.async; process-wide, returns the previous setting.  Same bits either way. */
int ak_mhd_set_march_tma(int on) { return akb::set_march_tma(on); }

}  // extern "C"
