/* athenak_b200.h -- C ABI of the B200-native MHD stage-update engine.
 *
 * Drop-in boundary: every entry point replaces the body of one task
 * that AthenaK's MHD module registers with TaskList::AddTask (reference
 * src/mhd/mhd_tasks.cpp:38-84; task signature src/tasklist/task_list.hpp:85-110).  The host
 * keeps owning every array (reference src/mhd/mhd.hpp:99-137); this library only borrows raw
 * device pointers (Kokkos View::data(), LayoutRight, no padding) for the duration of a call.
 *
 * Conventions
 *  - all functions return 0 on success, non-zero on error; ak_last_error() gives the text.
 *    A task wrapper turns non-zero into the reference's "### FATAL ERROR" + exit
 *    (e.g. src/mhd/mhd.cpp:85-87) and otherwise returns TaskStatus::complete.
 *  - single host thread per handle; all work is enqueued on the given stream, in order;
 *    only calls that return host scalars (ak_mhd_cons2prim counters, ak_mhd_new_dt)
 *    synchronise that stream.
 *  - there is no CPU fallback: without a CUDA device every compute call fails.
 *
 * Array layout (reference src/athena.hpp:111-130,178-231), FP64, i fastest:
 *   cell  (m,n,k,j,i)            n1=nx1+2ng, n2=nx2+2ng, n3=nx3+2ng
 *   faces x1f (m,[n,]k,j,i)  extents (n3,n2,n1+1); x2f (n3,n2+1,n1); x3f (n3+1,n2,n1)
 *   edges x1e (m,k,j,i) extents (n3+1,n2+1,n1); x2e (n3+1,n2,n1+1); x3e (n3,n2+1,n1+1)
 *   face EMFs e3x1,e2x1,e1x2,e3x2,e2x3,e1x3: cell-shaped (m,k,j,i) (src/mhd/mhd.cpp:349-354)
 */
#ifndef ATHENAK_B200_H_
#define ATHENAK_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef void *ak_stream_t; /* a cudaStream_t (0 = default stream) */

/* reconstruction / Riemann-solver selectors: reference src/athena.hpp:75, src/mhd/mhd.hpp:39 */
enum { AK_RECON_DC = 0, AK_RECON_PLM = 1, AK_RECON_PPM4 = 2, AK_RECON_PPMX = 3, AK_RECON_WENOZ = 4, AK_RECON_TENO = 5 };
enum { AK_RSOLVER_LLF = 0, AK_RSOLVER_HLLE = 1, AK_RSOLVER_HLLD = 2 };
/* equation of state: reference src/eos/eos.hpp:27-33 (EOS_Data::is_ideal).  The isothermal MHD solvers are the
 * `else` branches of llf/hlle (src/mhd/rsolvers) and HLLD of hlld_mhd.hpp:350-546; isothermal hydro has
 * llf, hlle and roe only (hydro.cpp:252-262 rejects hllc). */
enum { AK_EOS_IDEAL = 0, AK_EOS_ISOTHERMAL = 1 };
/* hydro solvers: reference src/hydro/hydro.hpp:32 (Hydro_RSolver), Newtonian ones only */
enum { AK_RSOLVER_HYD_ADVECT = 10, AK_RSOLVER_HYD_LLF = 11, AK_RSOLVER_HYD_HLLE = 12, AK_RSOLVER_HYD_HLLC = 13,
       AK_RSOLVER_HYD_ROE = 14 };

typedef struct {
  int nmb;           /* MeshBlocks in this pack (reference MeshBlockPack::nmb_thispack) */
  int ng;            /* ghost cells per side (mesh/nghost)                               */
  int nx1, nx2, nx3; /* active cells per block (RegionIndcs, src/mesh/mesh.hpp:35-41)    */
  int recon, rsolver;
  double gamma, dfloor, pfloor, tfloor, sfloor, sigma_max; /* EOS_Data, src/eos/eos.hpp:27-33 */
  int nscalars;      /* passive scalars (<mhd>/<hydro> nscalars, src/mhd/mhd.cpp, src/hydro/hydro.cpp:75):
                        stored after the fluid variables in u0, u1, w0 and the flux arrays */
  int eos;           /* AK_EOS_IDEAL (5 fluid variables) or AK_EOS_ISOTHERMAL (4: no IEN; <mhd>/<hydro> eos,
                        src/eos/isothermal_mhd.cpp, isothermal_hyd.cpp; gamma and the p/T/s floors are ignored) */
  double iso_sound_speed; /* EOS_Data::iso_cs (src/eos/eos.hpp:28), isothermal only */
  int fofc;          /* 1: first-order flux correction (<mhd>/fofc, src/mhd/mhd.cpp:197-201): ak_mhd_fluxes covers one
                        more face on each side along the face normal (src/mhd/mhd_fluxes.cpp:96-106) and the Fluxes task
                        must call ak_mhd_fofc after it.  Needs nghost >= 3 (dc/plm) or 4 (5-point stencils),
                        src/mhd/mhd.cpp:209-235.  MHD only: the reference's hydro FOFC never allocates its
                        utest / fofc arrays (src/hydro/hydro.cpp:38-39 vs src/mhd/mhd.cpp:369-374) and crashes. */
} ak_mhd_cfg;

typedef struct {
  double *u0, *u1, *w0, *bcc0;                       /* src/mhd/mhd.hpp:99-111 */
  double *b0x1f, *b0x2f, *b0x3f, *b1x1f, *b1x2f, *b1x3f;
  double *flx1, *flx2, *flx3;                        /* uflx.x1f/x2f/x3f      */
  double *e3x1, *e2x1, *e1x2, *e3x2, *e2x3, *e1x3;   /* face-centred EMFs     */
  double *e1, *e2, *e3;                              /* efld.x1e/x2e/x3e      */
  const double *mb_size; /* DEVICE [nmb][9]: x1min,x2min,x3min,x1max,x2max,x3max,dx1,dx2,dx3
                            (RegionSize, src/mesh/mesh.hpp:25-29)                        */
} ak_mhd_arrays;

int ak_mhd_create(const ak_mhd_cfg *cfg, void **handle);
int ak_mhd_destroy(void *handle);
const char *ak_last_error(void *handle);

/* task 1  MHD::CopyCons      src/mhd/mhd_tasks.cpp:162-170  (caller decides stage==1) */
int ak_mhd_copy_cons(void *h, const ak_mhd_arrays *a, ak_stream_t s);
/* task 2  MHD::Fluxes -> CalculateFluxes<rsolver>  src/mhd/mhd_fluxes.cpp:85-266 */
int ak_mhd_fluxes(void *h, const ak_mhd_arrays *a, ak_stream_t s);
/* task 5  MHD::RKUpdate      src/mhd/mhd_update.cpp:24-84 */
int ak_mhd_rk_update(void *h, const ak_mhd_arrays *a, double gam0, double gam1, double beta_dt, ak_stream_t s);
/* task 14 MHD::CornerE       src/mhd/mhd_corner_e.cpp:26-415 (3-D Newtonian branch) */
int ak_mhd_corner_e(void *h, const ak_mhd_arrays *a, ak_stream_t s);
/* task 17 MHD::CT            src/mhd/mhd_ct.cpp:23-80 */
int ak_mhd_ct(void *h, const ak_mhd_arrays *a, double gam0, double gam1, double beta_dt, ak_stream_t s);
/* task 27 MHD::ConToPrim -> IdealMHD::ConsToPrim  src/eos/ideal_mhd.cpp:30-134.
 * nfloor (host int[3], may be NULL) receives the dfloor/efloor/tfloor counts accumulated on the
 * device since they were last read (and zeroes them); when it is non-NULL the call synchronises
 * the stream, with NULL it is fully asynchronous. */
int ak_mhd_cons2prim(void *h, const ak_mhd_arrays *a, int il, int iu, int jl, int ju, int kl, int ku,
                     int *nfloor, ak_stream_t s);
/* task 28 MHD::NewTimeStep   src/mhd/mhd_newdt.cpp:31-174: dt3 (host double[3]) = min over
 * active cells of dx_d/(|v_d|+c_f,d); synchronises the stream. */
int ak_mhd_new_dt(void *h, const ak_mhd_arrays *a, double *dt3, ak_stream_t s);

/* Fused stage pieces (fewer kernels, no CopyCons pass, no e_cc / L-R state arrays):
 *   ak_mhd_stage_flux_emf : tasks 2 + 14 (Fluxes, CornerE)
 *   [tasks 15/16 SendE/RecvE: ak_halo_emf_* below]
 *   ak_mhd_stage_update   : tasks 1 + 5 + 17 (CopyCons folded into the stage-1 update, RKUpdate, CT);
 *                           u1/b1 (the RK registers) are saved for the active zones only.
 * Legal whenever nothing else consumes uflx/efld between the tasks (no diffusion/FOFC/source
 * terms).  ak_mhd_stage_fused = flux_emf + same-rank EMF correction through halo_plan (may be
 * NULL to skip it; it is a bitwise identity only if neighbouring blocks' shared faces hold
 * identical bits) + update; it refuses plans that have off-rank neighbours. */
int ak_mhd_stage_flux_emf(void *h, const ak_mhd_arrays *a, ak_stream_t s);
/* second half of task 2 when cfg.fofc = 1: MHD::FOFC, src/mhd/mhd_fofc.cpp:31-489 (Newtonian, no excision).
 * Estimates the updated state over [is-1,ie+1]^d from uflx / the face EMFs, u0, u1, bcc0 and b1 (so CopyCons must have
 * filled u1/b1 INCLUDING ghost zones: call ak_mhd_copy_cons at stage 1; ak_mhd_stage_update then skips its folded
 * copy), flags the cells whose ConsToPrim would need a floor and replaces the fluxes and face EMFs on every face of
 * a flagged cell with single-state LLF values.  nfofc (host int, may be NULL) receives the cells flagged since the
 * count was last read (EventCounters::nfofc, src/mesh/mesh.hpp:70-74) and makes the call synchronise.
 * With cfg.fofc = 1 ak_mhd_stage_flux_emf / ak_mhd_stage_fused take gam0 = gam1 = beta_dt from
 * ak_mhd_set_stage, which must precede them. */
int ak_mhd_fofc(void *h, const ak_mhd_arrays *a, double gam0, double gam1, double beta_dt, int *nfofc, ak_stream_t s);
int ak_mhd_set_stage(void *h, double gam0, double gam1, double beta_dt);
/* Optional, before ak_mhd_stage_flux_emf: announces the stage and its RK coefficients (as ak_mhd_set_stage) and lets the
 * library start the conserved-variable half of the stage update on an auxiliary stream as soon as the fluxes (and FOFC)
 * are done, concurrently with the corner-EMF kernel; the following ak_mhd_stage_update(h, a, stage, same coefficients)
 * joins it and performs the CT half only.  Same kernels, same results; AKB_OVERLAP_U=0 disables the overlap. */
int ak_mhd_stage_begin(void *h, int stage, double gam0, double gam1, double beta_dt);
/* after ak_mhd_stage_flux_emf: the auxiliary stream its early U update was queued on (NULL when that overlap is off); work queued
 * there runs after the U update and beside the corner-EMF kernel -- the host driver uses it for the ghost-zone push of u0 */
int ak_mhd_aux_stream(void *h, ak_stream_t *aux);
int ak_mhd_stage_update(void *h, const ak_mhd_arrays *a, int stage, double gam0, double gam1,
                        double beta_dt, ak_stream_t s);
int ak_mhd_stage_fused(void *h, const ak_mhd_arrays *a, int stage, double gam0, double gam1,
                       double beta_dt, void *halo_plan, ak_stream_t s);

/* ---- the two-kernel stage (north star: "one or two hand-written kernels" per MeshBlock update) --------------------------
 * Kernel 1 = ak_mhd_fluxes (task 2: reconstruction + Riemann solve, three sweeps in one forked launch group).
 * Kernel 2 = ak_mhd_stage_tail: tasks 14 (MHD::CornerE, mhd_corner_e.cpp:206-414), 1 (CopyCons, mhd_tasks.cpp:162-170),
 * 5 (RKUpdate, mhd_update.cpp:24-84), 17 (CT, mhd_ct.cpp:23-80), 27 (ConToPrim, eos/ideal_mhd.cpp:30-134,
 * isothermal_mhd.cpp:20-135) and, with want_dt = 1, the reduction of 28 (NewTimeStep, mhd_newdt.cpp:31-174) in ONE launch:
 * corner EMFs never leave the chip except within two edges of the block surface.  3-D blocks, same-level mesh, no FOFC
 * (ak_mhd_tail_supported says whether the handle qualifies; otherwise use the per-task or ak_mhd_stage_* calls above).
 *
 * in  a:   u0, u1 (read at stage > 1; at stage 1 its active cells receive the old u0, as CopyCons does), b0x?f, b1x?f
 *          (stage > 1), w0, bcc0, flx?, the six face EMFs; e1,e2,e3 are WRITTEN within two edges of the block surface.
 * out out: u0 (may be a->u0: updated in place), b0x?f, w0, bcc0 -- these three MUST be other buffers than a's: neighbouring
 *          tiles still read the old values.  The caller rotates its pointers afterwards (std::swap on the Views); at
 *          stage 1 the old b0 buffers ARE b1 from then on, so no face-field copy is made.
 * The tail kernel finishes every cell and face that does not touch the block surface.  The rest needs other blocks:
 *   SendE/RecvE on a->e1,e2,e3 (ak_halo_emf_*), then
 *   ak_mhd_stage_shell_ct  : task 17 for the faces with an edge on the block surface (same in / out convention), then
 *   the ghost exchange of out->u0, out->b0x?f and the physical boundaries, then
 *   ak_mhd_stage_shell_c2p : task 27 for the ghost zones and the outermost active layer of the NEW state a (= the rotated
 *          pointers); nfloor (3 ints) and dt3 (3 doubles), when not NULL, return the floor counters and the new-dt minima
 *          accumulated by both kernels (dt3 requires want_dt = 1 in the tail call) and synchronise the stream. */
int ak_mhd_tail_supported(void *h);

/* Flux + update fusion: tasks 2 (Fluxes, mhd_fluxes.cpp:85-266), 1 (CopyCons of u) and 5 (RKUpdate, mhd_update.cpp:24-84) as the
 * three sweep kernels alone.  Each sweep turns the fluxes at the two faces of a cell into its term of the flux divergence while
 * they are in registers; x1 and x2 store theirs in du, x3 adds its own in the reference's order and writes
 * u0 = gam0*u0 + gam1*u1 - beta_dt*divf (stage 1: u1's active cells receive the old u0).  a->flx1/2/3 receive ONLY their IDN
 * component (the corner EMFs upwind with it), the face EMFs are written as by ak_mhd_fluxes.  du = scratch of
 * 2 * nmb * nfl * n3*n2*n1 doubles.  3-D MHD without passive scalars / FOFC (ak_mhd_fluxes_update_supported).
 * Follow with ak_mhd_corner_e, SendE/RecvE, ak_mhd_ct (ak_mhd_stage_update with the U half already done), the ghost exchange
 * and ak_mhd_cons2prim. */
int ak_mhd_fluxes_update_supported(void *h);
/* task 17 (CT) with the stage-1 CopyCons of the face field folded in, for callers of ak_mhd_fluxes_update */
int ak_mhd_ct_save(void *h, const ak_mhd_arrays *a, int stage, double gam0, double gam1, double beta_dt, ak_stream_t s);
int ak_mhd_fluxes_update(void *h, const ak_mhd_arrays *a, double *du, int stage, double gam0, double gam1, double beta_dt,
                         ak_stream_t s);
int ak_mhd_stage_tail(void *h, const ak_mhd_arrays *a, const ak_mhd_arrays *out, int stage, double gam0, double gam1,
                      double beta_dt, int want_dt, ak_stream_t s);
int ak_mhd_stage_shell_ct(void *h, const ak_mhd_arrays *a, const ak_mhd_arrays *out, int stage, double gam0, double gam1,
                          double beta_dt, ak_stream_t s);
int ak_mhd_stage_shell_c2p(void *h, const ak_mhd_arrays *a, int *nfloor, double *dt3, ak_stream_t s);
/* c2p over all cells fused with the task-28 reduction over the active cells (last stage). */
int ak_mhd_cons2prim_newdt(void *h, const ak_mhd_arrays *a, int *nfloor, double *dt3, ak_stream_t s);

/* ---- hydro twin: the same stage update without magnetic fields --------
 * Replaces the task bodies Hydro::AssembleHydroTasks registers (reference src/hydro/
 * hydro_tasks.cpp:47-79): CopyCons :137-160, Fluxes -> CalculateFluxes<rsolver> (src/hydro/
 * hydro_fluxes.cpp:78-229; solvers src/hydro/rsolvers/{advect,llf,hlle,hllc,roe}_hyd.hpp),
 * RKUpdate (hydro_update.cpp:23-82), ConToPrim (hydro_tasks.cpp:404-412 -> src/eos/ideal_hyd.cpp:
 * 28-113), NewTimeStep (hydro_newdt.cpp:31-138).  cfg->rsolver is an AK_RSOLVER_HYD_* value,
 * sigma_max is ignored; MeshBlocks may be 1-D (nx2 = nx3 = 1), 2-D (nx3 = 1) or 3-D, inactive
 * dimensions having extent 1 and no ghost cells (reference src/mesh/mesh.cpp:285-333).  Only
 * u0, u1, w0, flx1..3 and mb_size of ak_mhd_arrays are read.  Ghost exchange: ak_halo_*_cc with
 * nvar = 5; physical boundaries: ak_bcs_hydro. */
int ak_hydro_create(const ak_mhd_cfg *cfg, void **handle);
int ak_hydro_destroy(void *handle);
int ak_hydro_copy_cons(void *h, const ak_mhd_arrays *a, ak_stream_t s);
int ak_hydro_fluxes(void *h, const ak_mhd_arrays *a, ak_stream_t s);
int ak_hydro_rk_update(void *h, const ak_mhd_arrays *a, double gam0, double gam1, double beta_dt, ak_stream_t s);
int ak_hydro_cons2prim(void *h, const ak_mhd_arrays *a, int il, int iu, int jl, int ju, int kl, int ku,
                       int *nfloor, ak_stream_t s);
int ak_hydro_new_dt(void *h, const ak_mhd_arrays *a, double *dt3, ak_stream_t s);
/* CopyCons (folded into the stage-1 update) + Fluxes + RKUpdate; c2p over all cells + new dt */
int ak_hydro_stage_fused(void *h, const ak_mhd_arrays *a, int stage, double gam0, double gam1,
                         double beta_dt, ak_stream_t s);
int ak_hydro_cons2prim_newdt(void *h, const ak_mhd_arrays *a, int *nfloor, double *dt3, ak_stream_t s);

/* History-output sums over the active cells of the pack (reference src/outputs/history.cpp:78-160
 * LoadHydroHistoryData, :272-372 LoadMHDHistoryData): sums[0..4] = volume-weighted mass, 1/2/3-mom,
 * tot-E; [5..7] = 1/2/3-KE; MHD handles add [8..10] = 1/2/3-ME.  *nhist = 8 (hydro) or 11 (MHD).
 * h from ak_mhd_create or ak_hydro_create.  Deterministic two-pass device reduction; the caller
 * adds the per-rank results (the reference's MPI_Reduce, history.cpp:385-392).  Synchronises. */
int ak_hst_sums(void *h, const ak_mhd_arrays *a, double *sums, int *nhist, ak_stream_t s);

/* ---- ghost-zone exchange (src/bvals): tasks 10/11 SendU/RecvU, 21/22 SendB/RecvB ---------
 * nghbr: HOST [nmb][56][4] = {gid, lev, rank, dest} per neighbour slot (NeighborBlock,
 * src/mesh/mesh.hpp:47-52; slot numbering src/mesh/nghbr_index.hpp:28-54).  gids_start is the
 * first gid owned by my_rank (contiguous ranges, src/mesh/load_balance.cpp:38-88).
 * Same-rank neighbours are filled by one gather kernel (the reference's direct-write path,
 * src/bvals/bvals_cc.cpp:122-135, without the intermediate buffer).  Off-rank neighbours go
 * through per-peer send/recv buffers laid out as the reference's rank-packed messages:
 * entries ordered by (m, slot), each nvar*ndat values linearised as
 * (i-il) + ni*((j-jl) + nj*((k-kl) + nk*v))  (src/bvals/bvals_cc.cpp:127). */
int ak_halo_create(const ak_mhd_cfg *cfg, const int *nghbr, int my_rank, int gids_start, void **plan);
int ak_halo_destroy(void *plan);
int ak_halo_npeers(void *plan);
/* peer rank and message length (doubles) for CC data with nvar variables, and for FC data */
int ak_halo_peer(void *plan, int ipeer, int *rank, long *cc_count_per_var, long *fc_count);
/* host-side view of the resolved copy plan (needs no device): kind 0 local_cc, 1 local_fc,
 * 2 pack_cc, 3 unpack_cc, 4 pack_fc, 5 unpack_fc; out [max][13] longs = m, src_m, comp,
 * lo[3], hi[3], shift[3], message offset.  Returns the number of descriptors, <0 on error. */
int ak_halo_descs(void *plan, int kind, int ipeer, long *out, int max);
int ak_halo_local_cc(void *plan, double *a, int nvar, ak_stream_t s);
int ak_halo_local_fc(void *plan, double *x1f, double *x2f, double *x3f, ak_stream_t s);
int ak_halo_pack_cc(void *plan, const double *a, int nvar, int ipeer, double *sendbuf, ak_stream_t s);
int ak_halo_unpack_cc(void *plan, double *a, int nvar, int ipeer, const double *recvbuf, ak_stream_t s);
int ak_halo_pack_fc(void *plan, const double *x1f, const double *x2f, const double *x3f, int ipeer,
                    double *sendbuf, ak_stream_t s);
int ak_halo_unpack_fc(void *plan, double *x1f, double *x2f, double *x3f, int ipeer,
                      const double *recvbuf, ak_stream_t s);

/* Peer-memory ghost exchange (one process per GPU on one NVLink/NVSwitch node).  Instead of
 * pack -> message -> unpack, ONE kernel writes every off-rank boundary slab of this rank straight
 * into the ghost zones of the neighbouring blocks inside the peer GPUs' arrays.  The host layer maps
 * each peer's u0 / x1f / x2f / x3f and flag array into this process (cudaIpcOpenMemHandle) and
 * registers them per peer (ipeer as in ak_halo_peer) together with the peer's first gid; flags are
 * 4*nranks monotone counters per rank ([kind][writer rank], kind 0 = "my ghost zones may be
 * overwritten", 1 = "your ghost zones are filled"; kinds 2 / 3 are the same pair for the EMF
 * receive segments: ak_halo_emf_pack may be handed a peer-mapped pointer into the neighbour's
 * segment (ak_halo_emf_rbuf base + that rank's ak_halo_emf_peer offset) instead of a send buffer).  Protocol per exchange with sequence number n:
 *   receiver: ak_halo_signal(plan, 0, n) once nothing reads its ghost zones any more
 *   sender:   ak_halo_wait(plan, 0, n); ak_halo_push(...); ak_halo_signal(plan, 1, n)
 *   receiver: ak_halo_wait(plan, 1, n) before the ghost zones are read
 * all stream-ordered (the waits are spinning single-block kernels).  Same bytes land in the same
 * places as with ak_halo_pack_* + message + ak_halo_unpack_*. */
int ak_halo_push_setup(void *plan, int ipeer, int peer_gids_start, double *peer_u0, double *peer_x1f,
                       double *peer_x2f, double *peer_x3f);
/* the peer's face-field arrays for the pushes that follow (the two-kernel stage below rotates its face-field buffers) */
int ak_halo_push_set_faces(void *plan, int ipeer, double *peer_x1f, double *peer_x2f, double *peer_x3f);
int ak_halo_push_flags(void *plan, int my_rank, int nranks, unsigned long long *local_flags,
                       unsigned long long *const *peer_flags /* host [npeers] */);
int ak_halo_push_commit(void *plan);
/* u0 == NULL skips the cell-centred data, x1f == NULL the face fields */
int ak_halo_push(void *plan, double *u0, int nvar, double *x1f, double *x2f, double *x3f, ak_stream_t s);
int ak_halo_signal(void *plan, int kind, unsigned long long value, ak_stream_t s);
/* base and length (doubles) of this rank's EMF receive buffer (all local + per-peer segments) */
int ak_halo_emf_rbuf(void *plan, double **base, long *count);
int ak_halo_wait(void *plan, int kind, unsigned long long value, ak_stream_t s);

/* Orszag-Tang vortex ICs (src/pgen/tests/orszag_tang.cpp:32-123), BASELINE config 3. */
int ak_pgen_orszag_tang(void *mesh, int rank, int ng, double gamma, double *u0, double *b1f,
                        double *b2f, double *b3f);

/* tasks 15/16 SendE/RecvE: same-level EMF flux correction (src/bvals/flux_correct_fc.cpp:29-437,
 * 445-644, 801-1034).  Every block's boundary edge EMFs are copied into the receive segments of
 * its face/edge neighbours (same rank: ak_halo_emf_pack_local; other ranks: ak_halo_emf_pack into
 * one message per peer, to be delivered at the address ak_halo_emf_peer returns), then
 * ak_halo_emf_sumavg adds the segments in increasing slot order and scales (x0.5 in face
 * interiors, / nflx on block edges) exactly as SumBoundaryFluxes + AverageBoundaryFluxes. */
int ak_halo_emf_peer(void *plan, int ipeer, long *count, double **recv_ptr);
int ak_halo_emf_pack_local(void *plan, const double *e1, const double *e2, const double *e3, ak_stream_t s);
int ak_halo_emf_pack(void *plan, const double *e1, const double *e2, const double *e3, int ipeer,
                     double *sendbuf, ak_stream_t s);
int ak_halo_emf_sumavg(void *plan, double *e1, double *e2, double *e3, ak_stream_t s);

/* task 26 MHD::ApplyPhysicalBCs (src/mhd/mhd_tasks.cpp:504-520): physical boundary conditions
 * on the blocks touching the mesh edge, MeshBoundaryValues::HydroBCs (src/bvals/physics/
 * hydro_bcs.cpp:27-335) and ::BFieldBCs (src/bvals/physics/bfield_bcs.cpp:25-358).  Flags are
 * the reference's BoundaryFlag values (src/bvals/bvals.hpp:19).  mesh_bcs[6] = ix1,ox1,ix2,ox2,
 * ix3,ox3 of the mesh; mb_bcs = host [nmb][6] per-block flags (MeshBlock::mb_bcs, block faces =
 * AK_BC_BLOCK).  The x1, x2, x3 passes run in that order over the full transverse extent.  A
 * strictly periodic mesh makes every call a no-op.  inflow / user / shear_periodic need
 * problem-generator data and are rejected by ak_bcs_create. */
enum ak_bc_flag { AK_BC_UNDEF = -1, AK_BC_BLOCK = 0, AK_BC_REFLECT = 1, AK_BC_INFLOW = 2, AK_BC_OUTFLOW = 3,
                  AK_BC_DIODE = 4, AK_BC_USER = 5, AK_BC_PERIODIC = 6, AK_BC_SHEAR_PERIODIC = 7, AK_BC_VACUUM = 8 };
int ak_bcs_create(const ak_mhd_cfg *cfg, const int *mesh_bcs, const int *mb_bcs, void **bcs);
int ak_bcs_destroy(void *bcs);
const char *ak_bcs_last_error(void *bcs);
int ak_bcs_strictly_periodic(void *bcs);
int ak_bcs_hydro(void *bcs, double *u0, int nvar, ak_stream_t s);
int ak_bcs_bfield(void *bcs, double *x1f, double *x2f, double *x3f, ak_stream_t s);
/* both of the above with one launch per direction */
int ak_bcs_apply(void *bcs, double *u0, int nvar, double *x1f, double *x2f, double *x3f, ak_stream_t s);

/* Device self-test: compares the library's FP64 reciprocal / division / sqrt (nvcc's fast paths + a zero-aware
 * range guard, csrc/mhd_device.cuh) with the built-in IEEE operators on n random operands.  exponent_span <= 1022:
 * |exponent| <= exponent_span, every 64th numerator exactly 0.  exponent_span > 1022: the whole double range --
 * subnormals, the denormal floor, +-0, infinities, NaNs, negative sqrt arguments -- in every operand position.
 * mismatches[5] = differing results for 1/x, a/b, sqrt(x), a/y by the reciprocal+FMA quotient used for the
 * launch-uniform divisors (dx, 12, 6), and -- [4] -- results of the DEFERRED form of the guard (flag accumulated, caller
 * re-evaluates) that differ from IEEE although the flag stayed clear. */
int ak_selftest_divsqrt(long n, unsigned long long seed, int exponent_span, long *mismatches);

/* Measured FP64 issue peak of the current device (csrc/peaks.cu): a register-only kernel of dependent chains of
 * one FP64 instruction.  op: 0 DFMA, 1 DADD, 2 DMUL, 3 alternating DADD/DMUL (what an --fmad=false build issues);
 * ilp = independent chains per thread (1, 2, 4, 8); warps_per_smsp = resident warps per SM sub-partition (1..16);
 * iters x 16 instructions per chain.  Returns thread-level FP64 instructions per second (the denominator of
 * bench.py's roofline.fp64) and the best-of-3 kernel time. */
int ak_peak_fp64(int op, int ilp, int warps_per_smsp, int iters, double *lane_instr_per_s, double *ms);

/* ---- host-side mesh metadata (CPU only; reference src/mesh, kept on the host) ---------------
 * Uniform (single-level) meshes.  Bit-exact mirrors of: Z-ordered gids (meshblock_tree.cpp:
 * 60-90,336-350), LoadBalance (load_balance.cpp:38-88), MeshBlock sizes and BC flags
 * (meshblock.cpp:25-131; enum ak_bc_flag), neighbour table {gid,lev,rank,dest}
 * (meshblock.cpp:142-430).  bcs[6] = ix1,ox1,ix2,ox2,ix3,ox3 as enum ak_bc_flag. */
int ak_mesh_create(const int *nmesh, const int *nblock, const double *xmin, const double *xmax,
                   const int *bcs, int nranks, void **mesh);
/* Static mesh refinement, host maps: the same tables for a mesh with <refined_region>
 * blocks (src/mesh/build_tree.cpp:79-230 -> MeshBlockTree::AddNode/Refine, src/mesh/meshblock_tree.cpp:97-249, which
 * keeps adjacent leaves within one level of each other).  regions = [nregions][7] doubles: physical level
 * (>= 1), x1min, x1max, x2min, x2max, x3min, x3max.  Z-ordered gids across levels, lloc[.][3] = logical level,
 * NeighborBlock.lev and the sub-block slots / destinations of coarser and finer neighbours as in
 * src/mesh/meshblock.cpp:142-430.  The device path for fine/coarse boundaries is ak_smr_* below. */
int ak_mesh_create_smr(const int *nmesh, const int *nblock, const double *xmin, const double *xmax, const int *bcs,
                       int nranks, int nregions, const double *regions, void **mesh);
int ak_mesh_destroy(void *mesh);
/* Adaptive mesh refinement, host maps: the reference's MeshRefinement bookkeeping for one cycle --
 * src/mesh/mesh_refinement.cpp:196-266 (cycle counters, level limits, <mesh_refinement>/refinement_interval and
 * ncycle_check), :274-419 (UpdateMeshBlockTree: refine flagged leaves keeping adjacent leaves within one level,
 * derefine complete sibling sets the tree allows), :427-510 (new Z-ordered block list and the new->old / old->new gid
 * maps) and :611-656.  ak_mesh_set_amr switches a mesh created by ak_mesh_create[_smr] to adaptive; ak_mesh_adapt takes
 * the refinement criteria's verdict per block (+1 / -1 / 0; ignored on cycles that are not check cycles), returns in the
 * same array what happened to each OLD block (+1 refined, -nleaf derefined, 0) and the number of blocks created /
 * deleted; when either is non-zero every ak_mesh_* table describes the NEW mesh afterwards.  Data movement: ak_amr_* below. */
int ak_mesh_set_amr(void *mesh, int num_levels, int refinement_interval, int ncycle_check);
int ak_mesh_adapt(void *mesh, int ncycle, int *refine_flag /*[nmb_total] in/out*/, int *nnew, int *ndel);
int ak_mesh_amr_maps(void *mesh, int *newtoold, int *oldtonew, int *old_flag, int *old_nmb);
/* ---- adaptive mesh refinement on the device (dynamic refinement) ----
 * ak_amr_criterion: one <amr_criterion> block (src/mesh/refinement_criteria.cpp:177-386; method 0 min_max, 1 slope,
 * 2 second_deriv, 3 location) as per-block extrema on the device + the reference's flag rules on the host.
 * ak_amr_plan / ak_amr_move: the data movement of MeshRefinement::RedistAndRefineMeshBlocks (src/mesh/mesh_refinement.cpp:
 * 523-603: DerefineCC/FC, CopyCC/FC, CopyForRefinementCC/FC, RefineCC/FC) as gathers from the OLD arrays into the NEW
 * arrays: the reference moves blocks in place inside over-allocated Views and stages through the coarse arrays; here
 * the host hands over both sets of arrays and every element of the new ones is computed in one launch per field with
 * the reference's expressions (RestrictCC/FC for derefined blocks, ProlongCC and the divergence-preserving face
 * operators for blocks created by refinement).  ak_amr_repair_fc: RepairAMRFC (:1189-1221).  Host sequence
 * (AdaptiveMeshRefinement, :145-190): criteria -> ak_mesh_adapt -> ak_amr_plan + ak_amr_move -> ghost exchange +
 * ConToPrim -> ak_amr_repair_fc -> ghost exchange + ConToPrim -> NewTimeStep.  plan rows = {kind, old block, offset[3]}. */
const char *ak_amr_last_error(void);
int ak_amr_criterion(const ak_mhd_cfg *cfg, int method, const double *q, int nvar, int var, double value_min, double value_max,
                     const double *loc4, const double *mb_size, int *flag /* host [nmb], in/out */, ak_stream_t s);
int ak_amr_plan(const ak_mhd_cfg *cfg_new, int old_nmb, const int *newtoold, const int *old_flag, const int *lloc_new, int *plan /*[new_nmb][5]*/);
int ak_amr_move(const ak_mhd_cfg *cfg_new, const int *plan, const double *u_old, double *u_new, int nvar,
                const double *b1_old, const double *b2_old, const double *b3_old, double *b1_new, double *b2_new, double *b3_new, ak_stream_t s);
int ak_amr_repair_fc(const ak_mhd_cfg *cfg, const int *blockflag /* host [nmb] */, double *b1, double *b2, double *b3, ak_stream_t s);

/* ---- fine/coarse boundaries of cell-centred data on a statically refined mesh, hydro ----
 * Replaces, for a pack whose neighbours are all on this rank, the multilevel parts of the tasks RestrictU
 * (src/mesh/mesh_refinement.cpp:1226-1284), SendU/RecvU (src/bvals/bvals_cc.cpp:42-167,273-402 with the
 * same/coarser/finer index tables of src/bvals/buffs_cc.cpp:28-484), Prolongate (src/bvals/prolongation.cpp:
 * 366-547: FillCoarseInBndryCC + ProlongateCC, operator src/mesh/prolongation.hpp:21-60) and SendFlux/RecvFlux
 * (src/bvals/flux_correct_cc.cpp:28-290), in the order of src/hydro/hydro_tasks.cpp:55-71:
 *   Fluxes -> ak_smr_flux_correct_cc -> RKUpdate -> ak_smr_restrict_cc -> ak_smr_exchange_cc ->
 *   ak_smr_prolongate_cc -> [physical BCs] -> ConToPrim.
 * nghbr = [nmb][56][4] NeighborBlock entries (gid, lev, rank, dest), mb_lev = [nmb] logical levels
 * (ak_mesh_create_smr); u / cu are the fine array and the host's coarse array coarse_u0
 * (nmb, nvar, cnx3+2ng, cnx2+2ng, cnx1+2ng in active dimensions; ak_smr_coarse_cells per variable and block). */
int ak_smr_create(const ak_mhd_cfg *cfg, const int *nghbr, const int *mb_lev, int my_rank, int gids_start, void **plan);
int ak_smr_destroy(void *plan);
const char *ak_smr_last_error(void *plan);
long ak_smr_coarse_cells(void *plan);
int ak_smr_counts(void *plan, long *counts /* [4]: exchange, fill, prolongation, flux descriptors */);
int ak_smr_restrict_cc(void *plan, const double *u, double *cu, int nvar, ak_stream_t s);
int ak_smr_exchange_cc(void *plan, double *u, double *cu, int nvar, ak_stream_t s);
int ak_smr_prolongate_cc(void *plan, double *u, double *cu, int nvar, ak_stream_t s);
/* steps: 1 = FillCoarseInBndryCC only, 2 = ProlongateCC only, 3 = both.  With physical boundaries the host applies
 * them to the COARSE array between the two steps (HydroBCsCoarse, src/bvals/physics/hydro_bcs.cpp:51-67: ak_bcs_hydro
 * on a plan created with nx -> nx/2) and to the fine array after prolongation (src/hydro/hydro_tasks.cpp:357-398). */
int ak_smr_prolongate_cc_steps(void *plan, double *u, double *cu, int nvar, int steps, ak_stream_t s);
int ak_smr_flux_correct_cc(void *plan, double *flx1, double *flx2, double *flx3, int nvar, ak_stream_t s);
/* ---- the same boundaries for FACE-CENTRED fields and edge EMFs: MHD on a statically refined mesh ----
 * Replaces the multilevel parts of RestrictB (src/mesh/mesh_refinement.cpp:1290-1383), SendB/RecvB
 * (src/bvals/bvals_fc.cpp:63-436 with the tables of src/bvals/buffs_fc.cpp:29-790), Prolongate for B
 * (src/bvals/prolongation.cpp:556-785: FillCoarseInBndryFC, shared faces with the ownership rule of :92-133, then the
 * divergence-preserving interior faces of :253-357) and SendE/RecvE with coarser and finer neighbours
 * (src/bvals/flux_correct_fc.cpp:29-1034), in the order of src/mhd/mhd_tasks.cpp:48-75:
 *   Fluxes -> ak_smr_flux_correct_cc -> RKUpdate -> ak_smr_restrict_cc -> ak_smr_exchange_cc -> CornerE ->
 *   ak_smr_emf_correct -> CT -> ak_smr_restrict_fc -> ak_smr_exchange_fc -> prolongate_cc(1) -> prolongate_fc(1)
 *   -> prolongate_cc(2) -> prolongate_fc(2) -> ConToPrim.
 * b1..b3 = b0.x1f/x2f/x3f, cb1..cb3 = the host's coarse_b0 (ak_smr_coarse_faces(plan, v) doubles per block),
 * e1..e3 = efld.x1e/x2e/x3e.  Where several neighbours deliver the same ghost face the reference's serial unpack
 * order decides (the highest slot wins); the plan resolves that once on the host.  With physical boundaries the host
 * applies them to the coarse arrays between steps 1 and 2 (BFieldBCsCoarse, src/bvals/physics/bfield_bcs.cpp:44-63:
 * ak_bcs_apply on a plan created with nx -> nx/2) and to the fine arrays after the prolongation. */
long ak_smr_coarse_faces(void *plan, int v);
int ak_smr_restrict_fc(void *plan, const double *b1, const double *b2, const double *b3, double *cb1, double *cb2, double *cb3, ak_stream_t s);
int ak_smr_exchange_fc(void *plan, double *b1, double *b2, double *b3, double *cb1, double *cb2, double *cb3, ak_stream_t s);
/* steps: 1 = FillCoarseInBndryFC, 2 = ProlongateFC, 3 = both */
int ak_smr_prolongate_fc_steps(void *plan, double *b1, double *b2, double *b3, double *cb1, double *cb2, double *cb3, int steps, ak_stream_t s);
int ak_smr_emf_correct(void *plan, double *e1, double *e2, double *e3, ak_stream_t s);
/* host-side tables of the plan (no device needed; consumed by tests/emul): which = 0 dims, 1-3 fine maps, 4-6 coarse
 * maps, 7 sources, 8 fill, 9 shared-face, 10 interior boxes, 11 EMF pack, 12 EMF ops, 13 EMF ranges, 14 EMF buffer size */
int ak_smr_fc_tables(void *plan, int which, const void **ptr, long *count);
const char *ak_mesh_last_error(void);
int ak_mesh_info(void *mesh, int *nmb_total, int *root_level);
int ak_mesh_tables(void *mesh, int *lloc /*[nmb_total][4]*/, int *rank_of, int *gids_start, int *nmb_rank);
int ak_mesh_rank_tables(void *mesh, int rank, double *mb_size /*[nmb][9]*/, int *mb_bcs /*[nmb][6]*/,
                        int *nghbr /*[nmb][56][4]*/);
int ak_neighbor_index(int ix, int iy, int iz, int n1, int n2); /* src/mesh/nghbr_index.hpp:28-54 */
/* Synthetic input: spherical blast ICs (src/pgen/fluids/blast.cpp:147-158,236-374), active
 * zones of the rank's blocks, HOST arrays in the device layout.  prm = {pi_amb, di_amb, prat,
 * drat, b_amb, inner_radius, outer_radius}. */
int ak_pgen_blast(void *mesh, int rank, int ng, double gamma, const double *prm, double *u0,
                  double *b1f, double *b2f, double *b3f);

/* ---- C++ host driver (athenak_b200/csrc/host_driver.cpp) -----------------------------------
 * TaskList / Driver / MHD task registration as in the reference (src/tasklist/task_list.hpp,
 * src/driver/driver.cpp:93-130,290-307,380-459, src/mhd/mhd_tasks.cpp:38-84, src/mesh/mesh.cpp:
 * 573-643), owning the device arrays of one MeshBlockPack and sequencing the calls above.
 * Multi-rank: grouped ncclSend/ncclRecv per exchange + ncclAllReduce(min) for dt. */
typedef struct {
  ak_mhd_cfg mhd;
  int integrator;        /* 1 rk1, 2 rk2, 3 rk3 (src/driver/driver.cpp:93-130) */
  double cfl_number, tlim;
  int nlim;              /* <0: no cycle limit */
  int my_rank, nranks, gids_start, nmb_total;
  int fused;             /* 1: ak_mhd_stage_fused + ak_mhd_cons2prim_newdt task list */
} ak_sim_cfg;
int ak_sim_create(const ak_sim_cfg *cfg, const double *mb_size /*host [nmb][9]*/,
                  const int *nghbr /*host [nmb][56][4]*/, void **sim);
int ak_sim_destroy(void *sim);
/* statically refined mesh (hydro, one rank): logical level per block; switches to the multilevel task list
 * (RestrictU, SendU/RecvU across levels, Prolongate, SendFlux/RecvFlux).  After ak_sim_set_bcs, before ak_sim_initialize. */
int ak_sim_set_levels(void *sim, const int *mb_lev);
/* Adaptive mesh refinement inside the C++ host driver, one rank (MeshRefinement::AdaptiveMeshRefinement, src/mesh/
 * mesh_refinement.cpp:145-190, called by Driver::Execute after every cycle, src/driver/driver.cpp:447-450).  mesh = the adaptive host
 * mesh (ak_mesh_create[_smr] + ak_mesh_set_amr) the simulation was created from (with ak_sim_set_levels); crit = ncrit rows of 8
 * doubles {method (0 min_max, 1 slope, 2 second_deriv, 3 location), 1 = primitive density / 0 = conserved, value_min, value_max,
 * location x1, x2, x3, radius}.  ak_sim_adapt = one AdaptiveMeshRefinement call (criteria -> ak_mesh_adapt -> new state owner ->
 * ak_amr_plan / ak_amr_move -> ghost zones + primitives -> ak_amr_repair_fc -> ghost zones + primitives -> new dt);
 * ak_sim_execute_amr = Driver::Execute with it (returns cycles done, -1 on error); ak_sim_amr_status: cnt[4] = MeshBlocks now,
 * mesh changes, blocks created, blocks deleted.  No FOFC / passive scalars on this path. */
int ak_sim_adapt(void *sim, void *mesh, const double *crit, int ncrit, int *changed);
int ak_sim_execute_amr(void *sim, void *mesh, const double *crit, int ncrit, int ncycles);
int ak_sim_amr_status(void *sim, long *cnt);
/* restart-style entries: (time, dt, ncycle) of a state loaded with ak_sim_set_state; ConToPrim over every cell + MHD::NewTimeStep of
 * that state (ghost zones as loaded); Mesh::NewTimeStep */
int ak_sim_set_time(void *sim, double time, double dt, int ncycle);
int ak_sim_prims_newdt(void *sim);
int ak_sim_mesh_new_dt(void *sim);
/* physical BCs (default strictly periodic): mesh_bcs[6], mb_bcs host [nmb][6]; errors -> ak_bcs_last_error(NULL) */
int ak_sim_set_bcs(void *sim, const int *mesh_bcs, const int *mb_bcs);
int ak_nccl_unique_id(void *out128);                         /* rank 0; broadcast by the launcher */
/* collective over all ranks: creates the NCCL communicator, then tries to set up the peer-memory
 * ghost exchange (cudaIpc handles of u0 / face fields / flags all-gathered and mapped; used by the
 * fused task list when EVERY rank succeeded, AKB_PEER_PUSH=0 disables it) */
int ak_sim_comm_init(void *sim, const void *unique_id128);
int ak_sim_set_state(void *sim, const double *u0, const double *b1f, const double *b2f, const double *b3f);
int ak_sim_get_state(void *sim, double *u0, double *b1f, double *b2f, double *b3f);
/* primitives w0 and cell-centred field bcc0 of the current state, ghost zones included, to host buffers (diagnostics) */
int ak_sim_get_prim(void *sim, double *w0, double *bcc0);
int ak_sim_initialize(void *sim);            /* Driver::Initialize: ghosts, c2p, first dt */
int ak_sim_execute(void *sim, int ncycles);  /* Driver::Execute; returns cycles done */
int ak_sim_sync(void *sim);
/* st[3] = time, dt, dtnew; cnt[8] = ncycle, kernel launches, nfloor d/e/t, cudaStream_t,
 * peer-memory ghost exchange active (0/1), exchanges done through it */
int ak_sim_status(void *sim, double *st, long *cnt);
int ak_sim_arrays(void *sim, ak_mhd_arrays *out);            /* the device arrays it owns */
/* x2 / x3 flux sweeps (PPM4, rows of >= 64 columns, even n1) through the TMA bulk-copy kernel instead of per-thread cp.async.
 * Process-wide; returns the previous setting; AKB_MARCH_TMA=0/1 sets the start-up value.  Results are bit-identical. */
int ak_mhd_set_march_tma(int on);
/* a few words device -> host, synchronously, by a kernel storing into mapped pinned memory instead of a D2H copy (which would
 * queue behind any bulk download in flight on the copy engine); bytes % 4 == 0 */
int ak_small_d2h_sync(void *dst_host, const void *src_dev, long bytes, ak_stream_t s);

/* ---- pipelined host-buffer stepping (csrc/host_driver.cpp) -----------------------------------
 * What a caller that keeps its state in HOST memory does per step is ak_sim_set_state (H2D) -> ak_sim_initialize ->
 * ak_sim_execute(n) -> ak_sim_get_state (D2H); in the reference that is the restart-file round trip around Driver::Execute
 * (src/outputs/restart.cpp, src/pgen/pgen.cpp:140-330, src/driver/driver.cpp:380-459).  A pipe runs that sequence for a stream of
 * independent jobs over several engine instances ("slots", created by ak_sim_create on the same device, owned by the caller): one
 * upload stream and one download stream shared by all slots, one worker thread per slot, so the upload of job k+1, the step of
 * job k and the download of job k-1 overlap and both PCIe directions stay busy.  Job k runs on slot k mod nsims and gives bit for
 * bit what the serial sequence gives.  Host buffers must stay valid until ak_pipe_wait returns (pinned memory for full speed). */
int ak_pipe_create(void *const *sims, int nsims, void **pipe);
long ak_pipe_submit(void *pipe, const double *u0, const double *b1f, const double *b2f, const double *b3f,
                    double *out_u0, double *out_b1f, double *out_b2f, double *out_b3f, int ncycles);   /* ticket >= 0, -1 on error */
int ak_pipe_wait(void *pipe, long ticket);                   /* ticket < 0: all jobs submitted so far; non-zero: a job failed */
int ak_pipe_streams(void *pipe, ak_stream_t *h2d, ak_stream_t *d2h);   /* for device-side timing of the pipe */
int ak_pipe_destroy(void *pipe);                             /* drains, joins the workers; the sims stay alive */

#ifdef __cplusplus
}
#endif
#endif /* ATHENAK_B200_H_ */
