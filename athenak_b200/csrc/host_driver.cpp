// host_driver.cpp -- C++ host layer above the C ABI: the reference's operator API for this path.
//
// Mirrors, with the same names and semantics, the pieces of the reference that stay on the host:
//   TaskID / Task / TaskList   src/tasklist/task_list.hpp:37-236 (bitmask dependencies, DoAvailable)
//   Driver                     src/driver/driver.cpp:93-130 (RK tables), :290-307 (ExecuteTaskList),
//                              :314-371,569-653 (Initialize), :380-459 (Execute)
//   Mesh::NewTimeStep          src/mesh/mesh.cpp:573-643
//   MHD (state owner + tasks)  src/mhd/mhd.hpp:87-207, src/mhd/mhd_tasks.cpp:38-84
// Every MHD task body is one call into the C ABI of this library (include/athenak_b200.h);
// the task signature is the reference's `TaskStatus(Driver*, int stage)`.  Inter-rank ghost
// exchange is grouped ncclSend/ncclRecv (one message per peer per exchange) and
// ncclAllReduce(min) for dt; NCCL is loaded at run time (libnccl.so.2) and only when nranks > 1.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include <bitset>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <list>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/athenak_b200.h"

#define AKB_MAX_PUSH_PEERS_HOST 32   // == AKB_MAX_PUSH_PEERS (csrc/halo.h)

namespace akh {

enum class TaskStatus { fail, complete, incomplete };
enum class TaskListStatus { running, stuck, complete, nothing_to_do };
class Driver;

class TaskID {  // task_list.hpp:37-82
 public:
  TaskID() = default;
  explicit TaskID(unsigned int id) { if (id == 0) bitfld_.reset(); else bitfld_.set(--id); }
  void Clear() { bitfld_.reset(); }
  bool CheckDependencies(const TaskID &dep) const { return ((bitfld_ & dep.bitfld_) == dep.bitfld_); }
  void SetComplete(const TaskID &rhs) { bitfld_ |= rhs.bitfld_; }
  bool operator==(const TaskID &rhs) const { return bitfld_ == rhs.bitfld_; }
  TaskID operator|(const TaskID &rhs) const { TaskID r; r.bitfld_ = bitfld_ | rhs.bitfld_; return r; }
 private:
  std::bitset<64> bitfld_;
};

class Task {  // task_list.hpp:85-110
 public:
  Task(TaskID id, TaskID dep, std::function<TaskStatus(Driver *, int)> func) : myid_(id), dep_(dep), func_(func) {}
  TaskStatus operator()(Driver *d, int s) { return func_(d, s); }
  TaskID GetID() { return myid_; }
  TaskID GetDependency() { return dep_; }
  void SetComplete() { complete_ = true; }
  void SetIncomplete() { complete_ = false; }
  bool IsComplete() { return complete_; }
 private:
  TaskID myid_, dep_;
  bool complete_ = false;
  std::function<TaskStatus(Driver *, int)> func_;
};

class TaskList {  // task_list.hpp:116-236
 public:
  bool IsComplete() {
    for (auto &it : task_list_) if (!(tasks_completed_.CheckDependencies(it.GetID()))) return false;
    return true;
  }
  int Size() { return (int)task_list_.size(); }
  static void TraceTask(Driver *d);     // AKB_TRACE=1: an event on the state owner's stream after every task body
  bool Empty() { return task_list_.empty(); }
  void Reset() { tasks_completed_.Clear(); for (auto &it : task_list_) it.SetIncomplete(); }
  TaskListStatus DoAvailable(Driver *d, int s) {
    for (auto &task : task_list_) {
      auto dep = task.GetDependency();
      if (tasks_completed_.CheckDependencies(dep) && !(task.IsComplete())) {
        TaskStatus status = task(d, s);
        TraceTask(d);
        if (status == TaskStatus::complete) { task.SetComplete(); tasks_completed_.SetComplete(task.GetID()); }
      }
    }
    return IsComplete() ? TaskListStatus::complete : TaskListStatus::running;
  }
  template <class F, class T>
  TaskID AddTask(F func, T *obj, TaskID &dep) {
    TaskID id((unsigned)task_list_.size() + 1);
    task_list_.push_back(Task(id, dep, [=](Driver *d, int s) mutable -> TaskStatus { return (obj->*func)(d, s); }));
    return id;
  }
 protected:
  std::list<Task> task_list_;
  TaskID tasks_completed_;
};

// ---- NCCL, loaded lazily -----------------------------------------------------------------
struct Nccl {
  typedef struct { char internal[128]; } UniqueId;
  typedef void *Comm;
  int (*GetUniqueId)(UniqueId *) = nullptr;
  int (*CommInitRank)(Comm *, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(Comm) = nullptr;
  int (*Send)(const void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, Comm, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
  static Nccl &get() {
    static Nccl n;
    if (!n.ok) {
      void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
      if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (h) {
        n.GetUniqueId = (int (*)(UniqueId *))dlsym(h, "ncclGetUniqueId");
        n.CommInitRank = (int (*)(Comm *, int, UniqueId, int))dlsym(h, "ncclCommInitRank");
        n.CommDestroy = (int (*)(Comm))dlsym(h, "ncclCommDestroy");
        n.Send = (int (*)(const void *, size_t, int, int, Comm, cudaStream_t))dlsym(h, "ncclSend");
        n.Recv = (int (*)(void *, size_t, int, int, Comm, cudaStream_t))dlsym(h, "ncclRecv");
        n.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
        n.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
        n.AllReduce = (int (*)(const void *, void *, size_t, int, int, Comm, cudaStream_t))dlsym(h, "ncclAllReduce");
        n.AllGather = (int (*)(const void *, void *, size_t, int, Comm, cudaStream_t))dlsym(h, "ncclAllGather");
        n.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
        n.ok = n.GetUniqueId && n.CommInitRank && n.Send && n.Recv && n.GroupStart && n.GroupEnd && n.AllReduce;
      }
    }
    return n;
  }
};
enum { kNcclFloat64 = 8, kNcclMin = 3, kNcclInt8 = 0 };  // nccl.h: ncclFloat64 = 8, ncclMin = 3, ncclInt8 = 0

[[noreturn]] static void Fatal(const char *file, int line, const std::string &msg) {
  // the reference's fatal-error convention (e.g. src/mhd/mhd.cpp:85-87)
  std::fprintf(stdout, "### FATAL ERROR in %s at line %d\n%s\n", file, line, msg.c_str());
  std::fflush(stdout);
  std::exit(EXIT_FAILURE);
}
#define AK_CHECK(h, call) do { if ((call) != 0) Fatal(__FILE__, __LINE__, ak_last_error(h)); } while (0)
#define CU_CHECK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) Fatal(__FILE__, __LINE__, cudaGetErrorString(e_)); } while (0)

struct MeshState {  // the members of Mesh this path uses (src/mesh/mesh.hpp:120-150)
  double time = 0.0, dt = (double)FLT_MAX, cfl_no = 0.3;
  int ncycle = 0, nmb_total = 1;
};

class MHD;

class Driver {  // driver.hpp / driver.cpp
 public:
  Driver(const std::string &integrator, double tlim_, int nlim_) : tlim(tlim_), nlim(nlim_) {
    if (integrator == "rk1") { nexp_stages = 1; gam0[0] = 0.0; gam1[0] = 1.0; beta[0] = 1.0; }
    else if (integrator == "rk2") {
      nexp_stages = 2; gam0[0] = 0.0; gam1[0] = 1.0; beta[0] = 1.0; gam0[1] = 0.5; gam1[1] = 0.5; beta[1] = 0.5;
    } else if (integrator == "rk3") {
      nexp_stages = 3; gam0[0] = 0.0; gam1[0] = 1.0; beta[0] = 1.0; gam0[1] = 0.25; gam1[1] = 0.75; beta[1] = 0.25;
      gam0[2] = 2.0 / 3.0; gam1[2] = 1.0 / 3.0; beta[2] = 2.0 / 3.0;
    } else {
      Fatal(__FILE__, __LINE__, "integrator must be rk1, rk2 or rk3 on this path");
    }
  }
  double gam0[4], gam1[4], beta[4];
  int nexp_stages = 0;
  double tlim;
  int nlim;
  std::uint64_t nmb_updated_ = 0;
  // AKB_TRACE=1: device time between consecutive task bodies of one cycle (diagnostic; printed to stderr by Execute)
  cudaStream_t trace_stream = nullptr;
  std::vector<cudaEvent_t> trace_events;
  void ExecuteTaskList(std::map<std::string, std::shared_ptr<TaskList>> &tl_map, const std::string &tl, int stage) {
    auto &lst = tl_map[tl];                           // driver.cpp:290-307
    if (lst->Empty()) return;
    lst->Reset();
    while (lst->DoAvailable(this, stage) != TaskListStatus::complete) {}
  }
  void Initialize(MHD *pmhd);
  int Execute(MHD *pmhd, int ncycles);
};

void TaskList::TraceTask(Driver *d) {
  if (!d->trace_stream) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, d->trace_stream);
  d->trace_events.push_back(e);
}

class MHD {  // state owner + tasks
 public:
  ak_mhd_cfg cfg;
  ak_mhd_arrays a;
  void *h = nullptr, *plan = nullptr, *bcs = nullptr;   // bcs == nullptr: strictly periodic mesh
  cudaStream_t st = nullptr;
  MeshState mesh;
  int my_rank, nranks;
  int device = 0;          // CUDA device this state lives on (the current device at construction); ak_sim_* calls select it
  bool fused = true;
  double dtnew = (double)FLT_MAX;
  long nfloor[3] = {0, 0, 0};
  long nfofc = 0;          // EventCounters::nfofc (per-task list only; the fused list leaves the count on the device)
  long launches = 0;
  std::map<std::string, std::shared_ptr<TaskList>> tl_map;
  std::vector<double *> owned;
  struct Peer { int rank; long ncc, nfc; double *scc, *rcc, *sfc, *rfc; long nemf = 0; double *semf = nullptr, *remf = nullptr;
                double *pemf = nullptr; /* my segment inside the peer's EMF receive buffer, mapped over NVLink */ };
  std::vector<Peer> peers;
  Nccl::Comm comm = nullptr;
  double *d_dt = nullptr;
  size_t n_u, n_f1, n_f2, n_f3;
  int nvar = 5;
  bool hydro = false, multi_d = true, three_d = true;
  // peer-memory ghost exchange (ak_halo_push_*): set up by SetupPeerPush() after the communicator exists
  int gids_start_ = 0;
  unsigned long long nemfx = 0;      // EMF exchanges completed through peer memory
  bool push = false;
  bool u_pushed = false;                      // this stage's u0 push is already in flight on the auxiliary stream
  cudaEvent_t ev_upush = nullptr;
  unsigned long long nexch = 0;               // exchanges completed (sequence number of the flags)
  unsigned long long *d_flags = nullptr;      // [2][nranks] counters written by the peers
  std::vector<void *> ipc_opened;
  // static mesh refinement (hydro): multi-level plan, coarse array, coarse-array BC plan (ak_sim_set_levels)
  std::vector<int> nghbr_h;
  std::vector<double> mb_size_h;     // host copy of the block extents (refinement criteria, AMR)
  void *smr = nullptr, *bcs_coarse = nullptr;
  int mesh_bcs_h[6] = {0, 0, 0, 0, 0, 0};
  std::vector<int> mb_bcs_h;
  double *coarse_u0 = nullptr;
  double *coarse_b0[3] = {nullptr, nullptr, nullptr};     // MHD: coarse face fields (mhd.cpp:163-176)
  // Two-kernel stage (ak_mhd_stage_tail): the face field and the primitives are written out of place, so the state owner
  // keeps one spare buffer of each and rotates the pointers in `a` after every stage; at stage 1 the old b0 buffers BECOME b1
  // (no CopyCons pass over the face field).  Peer-memory pushes follow the rotation through per-peer copies of the same
  // three buffer sets (every rank rotates at the same point of the task list).
  // Flux + update fusion (ak_mhd_fluxes_update): the sweeps finish the conserved update themselves; du = their scratch
  bool fluxupd = false;
  double *du = nullptr;
  bool tail = false;
  double *spare_b[3] = {nullptr, nullptr, nullptr}, *spare_w = nullptr, *spare_bcc = nullptr;
  ak_mhd_arrays tail_out;                 // outputs of the stage in flight (u0 in place, b0x?f / w0 / bcc0 = the spares)
  struct PeerFaces { double *b[3][3]; };  // [set][component]: a peer's three face-field buffer sets, in MY rotation order
  std::vector<PeerFaces> peer_faces;      // set 0 = its current b0, 1 = its b1, 2 = its spare

  MHD(const ak_mhd_cfg &c, const double *mb_size, const int *nghbr, int rank, int nranks_, int gids_start, int nmb_total)
      : cfg(c), my_rank(rank), nranks(nranks_), gids_start_(gids_start) {
    mesh.nmb_total = nmb_total;
    CU_CHECK(cudaGetDevice(&device));
    CU_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    hydro = c.rsolver >= AK_RSOLVER_HYD_ADVECT;     // hydro twin: Hydro::Hydro, hydro.cpp:100-120,268-290
    multi_d = c.nx2 > 1; three_d = c.nx3 > 1;       // mesh.cpp:70-79
    const size_t n1 = c.nx1 + 2 * c.ng, n2 = multi_d ? c.nx2 + 2 * c.ng : 1, n3 = three_d ? c.nx3 + 2 * c.ng : 1, nm = c.nmb;
    const size_t cc = n1 * n2 * n3;
    nvar = (c.eos == AK_EOS_ISOTHERMAL ? 4 : 5) + c.nscalars;   // mhd.cpp:148 nmhd, hydro.cpp:75 nhydro
    n_u = nm * nvar * cc; n_f1 = nm * n3 * n2 * (n1 + 1); n_f2 = nm * n3 * (n2 + 1) * n1; n_f3 = nm * (n3 + 1) * n2 * n1;
    auto alloc = [&](size_t n) { double *p; CU_CHECK(cudaMalloc(&p, n * 8)); CU_CHECK(cudaMemsetAsync(p, 0, n * 8, st)); owned.push_back(p); return p; };
    std::memset(&a, 0, sizeof(a));
    a.u0 = alloc(n_u); a.u1 = alloc(n_u); a.w0 = alloc(n_u);
    a.flx1 = alloc(nvar * n_f1); a.flx2 = alloc(nvar * n_f2); a.flx3 = alloc(nvar * n_f3);
    if (!hydro) {
      a.bcc0 = alloc(nm * 3 * cc);   // mhd.cpp:148-161
      a.b0x1f = alloc(n_f1); a.b0x2f = alloc(n_f2); a.b0x3f = alloc(n_f3);
      a.b1x1f = alloc(n_f1); a.b1x2f = alloc(n_f2); a.b1x3f = alloc(n_f3);                       // mhd.cpp:329-347
      a.e3x1 = alloc(nm * cc); a.e2x1 = alloc(nm * cc); a.e1x2 = alloc(nm * cc);
      a.e3x2 = alloc(nm * cc); a.e2x3 = alloc(nm * cc); a.e1x3 = alloc(nm * cc);
      a.e1 = alloc(nm * (n3 + 1) * (n2 + 1) * n1); a.e2 = alloc(nm * (n3 + 1) * n2 * (n1 + 1)); a.e3 = alloc(nm * n3 * (n2 + 1) * (n1 + 1));
    }
    double *ms = alloc(nm * 9);
    CU_CHECK(cudaMemcpyAsync(ms, mb_size, nm * 9 * 8, cudaMemcpyHostToDevice, st));
    a.mb_size = ms;
    d_dt = alloc(1);
    if (hydro ? ak_hydro_create(&cfg, &h) : ak_mhd_create(&cfg, &h)) Fatal(__FILE__, __LINE__, ak_last_error(nullptr));
    {
      const char *env = std::getenv("AKB_TAIL");
      tail = !hydro && !cfg.fofc && ak_mhd_tail_supported(h) && (env && env[0] == '1');      // measured slower than the task kernels so far: opt-in
      const char *envf = std::getenv("AKB_FLUXUPD");
      fluxupd = !tail && !hydro && ak_mhd_fluxes_update_supported(h) && (envf && envf[0] == '1');     // measured slower than the task kernels so far: opt-in
      if (fluxupd) du = alloc(2 * nm * (size_t)(c.eos == AK_EOS_ISOTHERMAL ? 4 : 5) * cc);
      if (tail) {
        spare_b[0] = alloc(n_f1); spare_b[1] = alloc(n_f2); spare_b[2] = alloc(n_f3);
        spare_w = alloc(n_u); spare_bcc = alloc(nm * 3 * cc);
      }
    }
    nghbr_h.assign(nghbr, nghbr + nm * 56 * 4);
    mb_size_h.assign(mb_size, mb_size + nm * 9);
    if (ak_halo_create(&cfg, nghbr, rank, gids_start, &plan)) Fatal(__FILE__, __LINE__, ak_last_error(nullptr));
    for (int p = 0; p < ak_halo_npeers(plan); ++p) {
      Peer pe;
      ak_halo_peer(plan, p, &pe.rank, &pe.ncc, &pe.nfc);
      // one contiguous [U | B] buffer per direction so both exchanges can travel as one message
      pe.scc = alloc(nvar * pe.ncc + pe.nfc); pe.sfc = pe.scc + nvar * pe.ncc;
      pe.rcc = alloc(nvar * pe.ncc + pe.nfc); pe.rfc = pe.rcc + nvar * pe.ncc;
      peers.push_back(pe);
    }
    CU_CHECK(cudaStreamSynchronize(st));
  }
  ~MHD() {
    if (st) cudaStreamSynchronize(st);
    for (void *p : ipc_opened) cudaIpcCloseMemHandle(p);
    if (ev_upush) cudaEventDestroy(ev_upush);
    if (d_flags) cudaFree(d_flags);
    if (comm && Nccl::get().CommDestroy) Nccl::get().CommDestroy(comm);
    ak_halo_destroy(plan);
    ak_bcs_destroy(bcs);
    ak_bcs_destroy(bcs_coarse);
    ak_smr_destroy(smr);
    ak_mhd_destroy(h);
    for (double *p : owned) cudaFree(p);
    if (st) cudaStreamDestroy(st);
  }

  // Peer-memory ghost exchange: every rank publishes cudaIpc handles of its u0 / face fields / flag
  // array (ncclAllGather), maps its neighbours' arrays and registers them with the halo plan.  All
  // ranks must agree: the outcome is min-reduced, and anything short of unanimous success leaves the
  // grouped ncclSend/ncclRecv path in place.  AKB_PEER_PUSH=0 disables it.
  // h[0..3] u0 / face fields, h[4] flags, h[5] EMF receive buffer; emf_rank/emf_off: where each of MY peers'
  // EMF messages lands inside that buffer (doubles from its base)
  struct PeerInfo { cudaIpcMemHandle_t h[12]; int gids_start; int ok; int npeers; int pad;
                    int emf_rank[AKB_MAX_PUSH_PEERS_HOST]; long emf_off[AKB_MAX_PUSH_PEERS_HOST]; };
  void SetupPeerPush(int gids_start) {
    Nccl &n = Nccl::get();
    const char *env = std::getenv("AKB_PEER_PUSH");
    if (nranks < 2 || !comm || !n.AllGather || (env && env[0] == '0')) return;
    int ok = 1;
    if (cudaMalloc(&d_flags, sizeof(unsigned long long) * 4 * nranks) != cudaSuccess) { ok = 0; d_flags = nullptr; }
    if (d_flags) CU_CHECK(cudaMemset(d_flags, 0, sizeof(unsigned long long) * 4 * nranks));
    PeerInfo mine;
    std::memset(&mine, 0, sizeof(mine));
    mine.gids_start = gids_start;
    double *arr[4] = {a.u0, a.b0x1f, a.b0x2f, a.b0x3f};
    for (int q = 0; q < 4 && ok; ++q)
      if (arr[q] && cudaIpcGetMemHandle(&mine.h[q], arr[q]) != cudaSuccess) ok = 0;
    if (tail) {      // the two other face-field buffer sets of the rotation: h[6..8] = b1, h[9..11] = spare
      double *more[6] = {a.b1x1f, a.b1x2f, a.b1x3f, spare_b[0], spare_b[1], spare_b[2]};
      for (int q = 0; q < 6 && ok; ++q)
        if (cudaIpcGetMemHandle(&mine.h[6 + q], more[q]) != cudaSuccess) ok = 0;
    }
    if (ok && cudaIpcGetMemHandle(&mine.h[4], d_flags) != cudaSuccess) ok = 0;
    if ((int)peers.size() > AKB_MAX_PUSH_PEERS_HOST) ok = 0;
    double *emf_base = nullptr;
    if (ok && !hydro) {    // EMF correction (SendE/RecvE) through peer memory as well
      long cnt = 0;
      if (ak_halo_emf_rbuf(plan, &emf_base, &cnt) || !emf_base || cudaIpcGetMemHandle(&mine.h[5], emf_base) != cudaSuccess) ok = 0;
      mine.npeers = (int)peers.size();
      for (size_t p = 0; p < peers.size() && ok; ++p) {
        if (ak_halo_emf_peer(plan, (int)p, &peers[p].nemf, &peers[p].remf)) { ok = 0; break; }
        mine.emf_rank[p] = peers[p].rank;
        mine.emf_off[p] = (long)(peers[p].remf - emf_base);
      }
    }
    mine.ok = ok;
    cudaGetLastError();
    char *d_send = nullptr, *d_recv = nullptr;
    CU_CHECK(cudaMalloc(&d_send, sizeof(PeerInfo)));
    CU_CHECK(cudaMalloc(&d_recv, sizeof(PeerInfo) * nranks));
    CU_CHECK(cudaMemcpyAsync(d_send, &mine, sizeof(PeerInfo), cudaMemcpyHostToDevice, st));
    n.AllGather(d_send, d_recv, sizeof(PeerInfo), kNcclInt8, comm, st);
    std::vector<PeerInfo> all(nranks);
    CU_CHECK(cudaMemcpyAsync(all.data(), d_recv, sizeof(PeerInfo) * nranks, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    for (int r = 0; r < nranks; ++r) ok = ok && all[r].ok;
    std::vector<unsigned long long *> pflags(peers.size(), nullptr);
    for (size_t p = 0; p < peers.size() && ok; ++p) {
      const PeerInfo &pi = all[peers[p].rank];
      void *ptr[12] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
      for (int q = 0; q < 12 && ok; ++q) {
        if ((q >= 1 && q <= 3 || q == 5) && hydro) continue;
        if (q >= 6 && !tail) continue;
        if (cudaIpcOpenMemHandle(&ptr[q], pi.h[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
        ipc_opened.push_back(ptr[q]);
      }
      if (!ok) break;
      if (ak_halo_push_setup(plan, (int)p, pi.gids_start, (double *)ptr[0], (double *)ptr[1], (double *)ptr[2], (double *)ptr[3])) ok = 0;
      pflags[p] = (unsigned long long *)ptr[4];
      if (tail) {
        PeerFaces pf;
        for (int c = 0; c < 3; ++c) { pf.b[0][c] = (double *)ptr[1 + c]; pf.b[1][c] = (double *)ptr[6 + c]; pf.b[2][c] = (double *)ptr[9 + c]; }
        peer_faces.push_back(pf);
      }
      if (!hydro) {      // my segment in the peer's EMF receive buffer
        peers[p].pemf = nullptr;
        for (int q = 0; q < pi.npeers; ++q)
          if (pi.emf_rank[q] == my_rank) peers[p].pemf = (double *)ptr[5] + pi.emf_off[q];
        if (!peers[p].pemf) ok = 0;
      }
    }
    if (ok && ak_halo_push_flags(plan, my_rank, nranks, d_flags, pflags.data())) ok = 0;
    if (ok && ak_halo_push_commit(plan)) ok = 0;
    cudaGetLastError();
    // unanimous?
    double flag = ok ? 1.0 : 0.0;
    CU_CHECK(cudaMemcpyAsync(d_dt, &flag, 8, cudaMemcpyHostToDevice, st));
    n.AllReduce(d_dt, d_dt, 1, kNcclFloat64, kNcclMin, comm, st);
    CU_CHECK(cudaMemcpyAsync(&flag, d_dt, 8, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    cudaFree(d_send); cudaFree(d_recv);
    push = flag > 0.5;
    // the EMF receive segments start out free: first exchange has sequence number 1
    if (push && !hydro) AK_CHECK(nullptr, ak_halo_signal(plan, 2, 1, st));
    if (!push && my_rank == 0) std::fprintf(stderr, "athenak_b200: peer-memory ghost exchange unavailable, using ncclSend/ncclRecv\n");
  }
  // receiver side: nothing of this rank reads its ghost zones any more -> peers may overwrite them
  void SignalFree() {
    if (!push) return;
    AK_CHECK(nullptr, ak_halo_signal(plan, 0, nexch + 1, st));
    launches += 1;
  }

  // MHD::AssembleMHDTasks, mhd_tasks.cpp:38-84 (the tasks that do work on a uniform mesh)
  void AssembleMHDTasks() {
    for (const char *n : {"before_timeintegrator", "before_stagen", "stagen", "after_stagen", "after_timeintegrator"})
      tl_map[n] = std::make_shared<TaskList>();
    TaskID none(0);
    auto &tl = tl_map["stagen"];
    if (hydro && smr) {   // multilevel mesh: SendFlux/RecvFlux between Fluxes and RKUpdate, RestrictU + Prolongate around SendU/RecvU
      TaskID copyu = tl->AddTask(&MHD::CopyCons, this, none);
      TaskID flux = tl->AddTask(&MHD::Fluxes, this, copyu);
      TaskID fc = tl->AddTask(&MHD::FluxCorrect, this, flux);
      TaskID rkupdt = tl->AddTask(&MHD::RKUpdate, this, fc);
      TaskID sendu = tl->AddTask(&MHD::SendU, this, rkupdt);
      TaskID recvu = tl->AddTask(&MHD::RecvU, this, sendu);
      TaskID bc = tl->AddTask(&MHD::ApplyPhysicalBCs, this, recvu);
      if (fused) {
        tl->AddTask(&MHD::ConToPrimNewDt, this, bc);
      } else {
        TaskID c2p = tl->AddTask(&MHD::ConToPrim, this, bc);
        tl->AddTask(&MHD::NewTimeStep, this, c2p);
      }
      return;
    }
    if (smr) {     // MHD on a multilevel mesh, mhd_tasks.cpp:48-75: no fused flux+EMF / update+CT calls (flux correction
                   // sits between Fluxes and RKUpdate); Restrict* ride with the sends, Prolongate after RecvB
      if (fused) {   // corner EMFs + their correction ahead of the update (they read w0 / bcc0 / the corrected fluxes
                     // only), then CopyCons + RKUpdate + CT as one kernel; new dt on the last c2p
        TaskID flux = tl->AddTask(&MHD::Fluxes, this, none);
        TaskID fc = tl->AddTask(&MHD::FluxCorrect, this, flux);
        TaskID efld = tl->AddTask(&MHD::EField, this, fc);
        TaskID sende = tl->AddTask(&MHD::SendE, this, efld);
        TaskID recve = tl->AddTask(&MHD::RecvE, this, sende);
        TaskID up = tl->AddTask(&MHD::StageUpdate, this, recve);
        TaskID sendu = tl->AddTask(&MHD::SendU, this, up);
        TaskID recvu = tl->AddTask(&MHD::RecvU, this, sendu);
        TaskID sendb = tl->AddTask(&MHD::SendB, this, recvu);
        TaskID recvb = tl->AddTask(&MHD::RecvB, this, sendb);
        TaskID prol = tl->AddTask(&MHD::Prolongate, this, recvb);
        TaskID bc = tl->AddTask(&MHD::ApplyPhysicalBCs, this, prol);
        tl->AddTask(&MHD::ConToPrimNewDt, this, bc);
        return;
      }
      TaskID copyu = tl->AddTask(&MHD::CopyCons, this, none);
      TaskID flux = tl->AddTask(&MHD::Fluxes, this, copyu);
      TaskID fc = tl->AddTask(&MHD::FluxCorrect, this, flux);
      TaskID rkupdt = tl->AddTask(&MHD::RKUpdate, this, fc);
      TaskID sendu = tl->AddTask(&MHD::SendU, this, rkupdt);
      TaskID recvu = tl->AddTask(&MHD::RecvU, this, sendu);
      TaskID efld = tl->AddTask(&MHD::EField, this, recvu);
      TaskID sende = tl->AddTask(&MHD::SendE, this, efld);
      TaskID recve = tl->AddTask(&MHD::RecvE, this, sende);
      TaskID ct = tl->AddTask(&MHD::CT, this, recve);
      TaskID sendb = tl->AddTask(&MHD::SendB, this, ct);
      TaskID recvb = tl->AddTask(&MHD::RecvB, this, sendb);
      TaskID prol = tl->AddTask(&MHD::Prolongate, this, recvb);
      TaskID bc = tl->AddTask(&MHD::ApplyPhysicalBCs, this, prol);
      TaskID c2p = tl->AddTask(&MHD::ConToPrim, this, bc);
      tl->AddTask(&MHD::NewTimeStep, this, c2p);
      return;
    }
    if (hydro) {   // Hydro::AssembleHydroTasks, hydro_tasks.cpp:47-79 (uniform mesh, no source terms)
      if (fused) {
        TaskID f = tl->AddTask(&MHD::HydroStageFused, this, none);
        TaskID su = tl->AddTask(&MHD::SendU, this, f);
        TaskID ru = tl->AddTask(&MHD::RecvU, this, su);
        TaskID bc = tl->AddTask(&MHD::ApplyPhysicalBCs, this, ru);
        tl->AddTask(&MHD::ConToPrimNewDt, this, bc);
      } else {
        TaskID copyu = tl->AddTask(&MHD::CopyCons, this, none);
        TaskID flux = tl->AddTask(&MHD::Fluxes, this, copyu);
        TaskID rkupdt = tl->AddTask(&MHD::RKUpdate, this, flux);
        TaskID sendu = tl->AddTask(&MHD::SendU, this, rkupdt);
        TaskID recvu = tl->AddTask(&MHD::RecvU, this, sendu);
        TaskID bc = tl->AddTask(&MHD::ApplyPhysicalBCs, this, recvu);
        TaskID c2p = tl->AddTask(&MHD::ConToPrim, this, bc);
        tl->AddTask(&MHD::NewTimeStep, this, c2p);
      }
      return;
    }
    if (fused && tail) {     // two-kernel stage: flux sweeps + ONE tail kernel; the outermost layer waits for the EMF correction
      TaskID f = tl->AddTask(&MHD::StageFluxes, this, none);                // task 2
      TaskID t = tl->AddTask(&MHD::StageTail, this, f);                     // tasks 14, 1, 5, 17, 27 (28) for everything but the outermost layer
      TaskID se = tl->AddTask(&MHD::SendE, this, t);                        // tasks 15/16 on the surface EMFs the tail kernel left in HBM
      TaskID re = tl->AddTask(&MHD::RecvE, this, se);
      TaskID sc = tl->AddTask(&MHD::ShellCT, this, re);                     // task 17 for the faces that touch the block surface
      TaskID su = tl->AddTask(&MHD::SendUB, this, sc);
      TaskID ru = tl->AddTask(&MHD::RecvUB, this, su);
      TaskID bc = tl->AddTask(&MHD::ApplyPhysicalBCs, this, ru);
      tl->AddTask(&MHD::ShellC2PNewDt, this, bc);                           // tasks 27 (28) for the ghost zones + the outermost layer
    } else if (fused && fluxupd) {     // flux + update fusion: no flux arrays, no update kernel
      TaskID f = tl->AddTask(&MHD::StageFluxUpdate, this, none);           // tasks 2 + 1 + 5
      TaskID ef = tl->AddTask(&MHD::EField, this, f);                       // task 14
      TaskID se = tl->AddTask(&MHD::SendE, this, ef);
      TaskID re = tl->AddTask(&MHD::RecvE, this, se);
      TaskID ct = tl->AddTask(&MHD::CTSave, this, re);                      // task 17 (+ CopyCons of b at stage 1)
      TaskID su = tl->AddTask(&MHD::SendUB, this, ct);
      TaskID ru = tl->AddTask(&MHD::RecvUB, this, su);
      TaskID bc = tl->AddTask(&MHD::ApplyPhysicalBCs, this, ru);
      tl->AddTask(&MHD::ConToPrimNewDt, this, bc);
    } else if (fused) {
      TaskID f = tl->AddTask(&MHD::StageFluxEmf, this, none);
      TaskID se = tl->AddTask(&MHD::SendE, this, f);
      TaskID re = tl->AddTask(&MHD::RecvE, this, se);
      TaskID up = tl->AddTask(&MHD::StageUpdate, this, re);
      TaskID su = tl->AddTask(&MHD::SendUB, this, up);     // tasks 10+21 as one message per peer
      TaskID ru = tl->AddTask(&MHD::RecvUB, this, su);     // tasks 11+22
      TaskID bc = tl->AddTask(&MHD::ApplyPhysicalBCs, this, ru);
      tl->AddTask(&MHD::ConToPrimNewDt, this, bc);
    } else {
      TaskID copyu = tl->AddTask(&MHD::CopyCons, this, none);
      TaskID flux = tl->AddTask(&MHD::Fluxes, this, copyu);
      TaskID rkupdt = tl->AddTask(&MHD::RKUpdate, this, flux);
      TaskID sendu = tl->AddTask(&MHD::SendU, this, rkupdt);
      TaskID recvu = tl->AddTask(&MHD::RecvU, this, sendu);
      TaskID efld = tl->AddTask(&MHD::EField, this, recvu);
      TaskID sende = tl->AddTask(&MHD::SendE, this, efld);
      TaskID recve = tl->AddTask(&MHD::RecvE, this, sende);
      TaskID ct = tl->AddTask(&MHD::CT, this, recve);
      TaskID sendb = tl->AddTask(&MHD::SendB, this, ct);
      TaskID recvb = tl->AddTask(&MHD::RecvB, this, sendb);
      TaskID bc = tl->AddTask(&MHD::ApplyPhysicalBCs, this, recvb);
      TaskID c2p = tl->AddTask(&MHD::ConToPrim, this, bc);
      tl->AddTask(&MHD::NewTimeStep, this, c2p);
    }
  }

  // ---- tasks ------------------------------------------------------------------------------
  int nsweeps() const { return 1 + (multi_d ? 1 : 0) + (three_d ? 1 : 0); }
  TaskStatus CopyCons(Driver *pd, int stage) {
    if (stage == 1) {
      if (hydro) { AK_CHECK(h, ak_hydro_copy_cons(h, &a, st)); launches += 1; }
      else { AK_CHECK(h, ak_mhd_copy_cons(h, &a, st)); launches += 4; }
    }
    return TaskStatus::complete;
  }
  TaskStatus Fluxes(Driver *pd, int stage) {
    if (hydro) AK_CHECK(h, ak_hydro_fluxes(h, &a, st)); else AK_CHECK(h, ak_mhd_fluxes(h, &a, st));
    launches += nsweeps();
    if (cfg.fofc) {     // mhd_tasks.cpp:208-211
      int n = 0;
      AK_CHECK(h, ak_mhd_fofc(h, &a, pd->gam0[stage - 1], pd->gam1[stage - 1], pd->beta[stage - 1] * mesh.dt, &n, st));
      nfofc += n;
      launches += 1 + nsweeps();
    }
    return TaskStatus::complete;
  }
  TaskStatus RKUpdate(Driver *pd, int stage) {
    const double g0 = pd->gam0[stage - 1], g1 = pd->gam1[stage - 1], bdt = pd->beta[stage - 1] * mesh.dt;
    if (hydro) AK_CHECK(h, ak_hydro_rk_update(h, &a, g0, g1, bdt, st)); else AK_CHECK(h, ak_mhd_rk_update(h, &a, g0, g1, bdt, st));
    launches += 1;
    return TaskStatus::complete;
  }
  TaskStatus HydroStageFused(Driver *pd, int stage) {   // CopyCons + Fluxes + RKUpdate
    SignalFree();    // u0 ghost zones were last read by the previous c2p
    AK_CHECK(h, ak_hydro_stage_fused(h, &a, stage, pd->gam0[stage - 1], pd->gam1[stage - 1], pd->beta[stage - 1] * mesh.dt, st));
    launches += nsweeps() + 1;
    return TaskStatus::complete;
  }
  TaskStatus EField(Driver *pd, int stage) { AK_CHECK(h, ak_mhd_corner_e(h, &a, st)); launches += 1; return TaskStatus::complete; }
  TaskStatus CT(Driver *pd, int stage) {
    AK_CHECK(h, ak_mhd_ct(h, &a, pd->gam0[stage - 1], pd->gam1[stage - 1], pd->beta[stage - 1] * mesh.dt, st));
    launches += 1;
    return TaskStatus::complete;
  }
  TaskStatus StageFluxEmf(Driver *pd, int stage) {   // tasks 2 + 14
    if (cfg.fofc) {     // the FOFC estimate reads u1/b1 in the ghost zones: whole-array CopyCons, RK coefficients
      CopyCons(pd, stage);
      launches += 1 + nsweeps();
    }
    // RK coefficients for FOFC + the U half of the update overlapped with the corner-EMF kernel
    AK_CHECK(h, ak_mhd_stage_begin(h, stage, pd->gam0[stage - 1], pd->gam1[stage - 1], pd->beta[stage - 1] * mesh.dt));
    AK_CHECK(h, ak_mhd_stage_flux_emf(h, &a, st));
    static const bool overlap_u = !(std::getenv("AKB_OVERLAP_U") && std::getenv("AKB_OVERLAP_U")[0] == '0');
    launches += nsweeps() + 1 + (overlap_u ? 1 : 0);    // sweeps, corner EMFs, (early U update)
    // Overlapped ghost exchange: the U update already ran on the library's auxiliary stream beside the corner-EMF kernel, so the
    // push of u0 (5 of the 8 exchanged fields) is queued right behind it there -- "my ghost zones are free", wait for the
    // neighbours' same signal, push -- and overlaps corner EMFs, SendE/RecvE and CT; SendUB then pushes the face field only and
    // publishes "filled" once both pushes are done.  AKB_OVERLAP_PUSH=0 keeps the single push after CT.
    static const bool overlap_push = !(std::getenv("AKB_OVERLAP_PUSH") && std::getenv("AKB_OVERLAP_PUSH")[0] == '0');
    cudaStream_t aux = nullptr;
    if (UsePush() && overlap_push) AK_CHECK(h, ak_mhd_aux_stream(h, reinterpret_cast<ak_stream_t *>(&aux)));
    if (aux) {
      if (!ev_upush) CU_CHECK(cudaEventCreateWithFlags(&ev_upush, cudaEventDisableTiming));
      const unsigned long long seq = nexch + 1;
      AK_CHECK(nullptr, ak_halo_signal(plan, 0, seq, aux));      // the sweeps (joined before the fork) were the last readers of my ghost zones
      AK_CHECK(nullptr, ak_halo_wait(plan, 0, seq, aux));
      AK_CHECK(nullptr, ak_halo_push(plan, a.u0, nvar, nullptr, nullptr, nullptr, aux));
      CU_CHECK(cudaEventRecord(ev_upush, aux));
      u_pushed = true;
      launches += 3;
    } else {
      SignalFree();    // the flux sweeps were the last readers of the b0 ghost faces (u0 ghosts: c2p of the previous stage)
    }
    return TaskStatus::complete;
  }
  TaskStatus StageUpdate(Driver *pd, int stage) {    // tasks 1 + 5 + 17
    AK_CHECK(h, ak_mhd_stage_update(h, &a, stage, pd->gam0[stage - 1], pd->gam1[stage - 1], pd->beta[stage - 1] * mesh.dt, st));
    launches += 1;
    return TaskStatus::complete;
  }
  // ---- flux + update fusion ---------------------------------------------------------------
  TaskStatus StageFluxUpdate(Driver *pd, int stage) {
    AK_CHECK(h, ak_mhd_fluxes_update(h, &a, du, stage, pd->gam0[stage - 1], pd->gam1[stage - 1], pd->beta[stage - 1] * mesh.dt, st));
    launches += nsweeps();
    SignalFree();    // the sweeps were the last readers of the b0 ghost faces (u0 ghosts: the c2p of the previous stage)
    return TaskStatus::complete;
  }
  TaskStatus CTSave(Driver *pd, int stage) {
    AK_CHECK(h, ak_mhd_ct_save(h, &a, stage, pd->gam0[stage - 1], pd->gam1[stage - 1], pd->beta[stage - 1] * mesh.dt, st));
    launches += 1;
    return TaskStatus::complete;
  }
  // ---- two-kernel stage -------------------------------------------------------------------
  TaskStatus StageFluxes(Driver *pd, int stage) {
    AK_CHECK(h, ak_mhd_fluxes(h, &a, st));
    launches += nsweeps();
    SignalFree();    // the sweeps were the last readers of the b0 ghost faces (u0 ghosts: the shell c2p of the previous stage)
    return TaskStatus::complete;
  }
  TaskStatus StageTail(Driver *pd, int stage) {
    tail_out = a;
    tail_out.b0x1f = spare_b[0]; tail_out.b0x2f = spare_b[1]; tail_out.b0x3f = spare_b[2];
    tail_out.w0 = spare_w; tail_out.bcc0 = spare_bcc;
    AK_CHECK(h, ak_mhd_stage_tail(h, &a, &tail_out, stage, pd->gam0[stage - 1], pd->gam1[stage - 1], pd->beta[stage - 1] * mesh.dt,
                                  stage == pd->nexp_stages ? 1 : 0, st));
    launches += 1;
    return TaskStatus::complete;
  }
  TaskStatus ShellCT(Driver *pd, int stage) {
    AK_CHECK(h, ak_mhd_stage_shell_ct(h, &a, &tail_out, stage, pd->gam0[stage - 1], pd->gam1[stage - 1], pd->beta[stage - 1] * mesh.dt, st));
    launches += 1;
    // rotate: the outputs become the state; at stage 1 the old b0 buffers become the register b1 (CopyCons without a copy)
    double **cur[3] = {&a.b0x1f, &a.b0x2f, &a.b0x3f}, **reg[3] = {&a.b1x1f, &a.b1x2f, &a.b1x3f};
    for (int c = 0; c < 3; ++c) {
      double *oc = *cur[c], *orr = *reg[c], *osp = spare_b[c];
      *cur[c] = osp;
      if (stage == 1) { *reg[c] = oc; spare_b[c] = orr; } else { spare_b[c] = oc; }
      for (auto &pf : peer_faces) {       // the same permutation on every neighbour's buffer sets
        double *pc = pf.b[0][c], *pr = pf.b[1][c], *ps = pf.b[2][c];
        pf.b[0][c] = ps;
        if (stage == 1) { pf.b[1][c] = pc; pf.b[2][c] = pr; } else { pf.b[2][c] = pc; }
      }
    }
    std::swap(a.w0, spare_w);
    std::swap(a.bcc0, spare_bcc);
    for (size_t p = 0; p < peer_faces.size(); ++p)
      AK_CHECK(nullptr, ak_halo_push_set_faces(plan, (int)p, peer_faces[p].b[0][0], peer_faces[p].b[0][1], peer_faces[p].b[0][2]));
    return TaskStatus::complete;
  }
  TaskStatus ShellC2PNewDt(Driver *pd, int stage) {
    if (stage != pd->nexp_stages) {
      AK_CHECK(h, ak_mhd_stage_shell_c2p(h, &a, nullptr, nullptr, st));
    } else {
      int nf[3];
      double dt3[3];
      AK_CHECK(h, ak_mhd_stage_shell_c2p(h, &a, nf, dt3, st));
      for (int q = 0; q < 3; ++q) nfloor[q] += nf[q];
      dtnew = MinDt(dt3);
    }
    launches += 1;
    return TaskStatus::complete;
  }
  // SendE/RecvE: flux_correct_fc.cpp:29-437 (EMF flux correction, always executed as in the reference)
  TaskStatus SendE(Driver *pd, int stage) {
    if (smr) {       // coarser / finer neighbours: restricted pack + gather-form sum / zero / average (flux_correct_fc.cpp)
      if (ak_smr_emf_correct(smr, a.e1, a.e2, a.e3, st)) Fatal(__FILE__, __LINE__, ak_smr_last_error(smr));
      launches += 2;
      return TaskStatus::complete;
    }
    Nccl &n = Nccl::get();
    if (UsePush()) {
      // peer-memory variant: the pack kernel of each peer stores straight into my segment of that peer's
      // receive buffer over NVLink (no send buffer, no message); handshake kinds 2 ("segment free") / 3 ("filled")
      const unsigned long long seq = nemfx + 1;
      AK_CHECK(nullptr, ak_halo_wait(plan, 2, seq, st));
      for (size_t p = 0; p < peers.size(); ++p) AK_CHECK(nullptr, ak_halo_emf_pack(plan, a.e1, a.e2, a.e3, (int)p, peers[p].pemf, st));
      AK_CHECK(nullptr, ak_halo_signal(plan, 3, seq, st));
      AK_CHECK(nullptr, ak_halo_emf_pack_local(plan, a.e1, a.e2, a.e3, st));
      launches += 3 + (long)peers.size();
      return TaskStatus::complete;
    }
    if (!peers.empty()) {
      if (!comm) Fatal(__FILE__, __LINE__, "MeshBlocks on other ranks but no NCCL communicator (call ak_sim_comm_init)");
      for (size_t p = 0; p < peers.size(); ++p) {
        if (!peers[p].semf) {
          AK_CHECK(nullptr, ak_halo_emf_peer(plan, (int)p, &peers[p].nemf, &peers[p].remf));
          CU_CHECK(cudaMalloc(&peers[p].semf, (size_t)peers[p].nemf * 8));
          owned.push_back(peers[p].semf);
        }
        AK_CHECK(nullptr, ak_halo_emf_pack(plan, a.e1, a.e2, a.e3, (int)p, peers[p].semf, st));
        launches += 1;
      }
      n.GroupStart();
      for (auto &pe : peers) {
        n.Send(pe.semf, (size_t)pe.nemf, kNcclFloat64, pe.rank, comm, st);
        n.Recv(pe.remf, (size_t)pe.nemf, kNcclFloat64, pe.rank, comm, st);   // straight into the plan's segment
      }
      n.GroupEnd();
    }
    AK_CHECK(nullptr, ak_halo_emf_pack_local(plan, a.e1, a.e2, a.e3, st));
    launches += 1;
    return TaskStatus::complete;
  }
  TaskStatus RecvE(Driver *pd, int stage) {
    if (smr) return TaskStatus::complete;
    if (UsePush()) { AK_CHECK(nullptr, ak_halo_wait(plan, 3, nemfx + 1, st)); launches += 1; }
    AK_CHECK(nullptr, ak_halo_emf_sumavg(plan, a.e1, a.e2, a.e3, st));
    launches += 1;
    if (UsePush()) {      // segments consumed: peers may fill them for the next exchange
      nemfx += 1;
      AK_CHECK(nullptr, ak_halo_signal(plan, 2, nemfx + 1, st));
      launches += 1;
    }
    return TaskStatus::complete;
  }
  // SendU/RecvU, SendB/RecvB: bvals_cc.cpp:42-447, bvals_fc.cpp:63-436
  void Exchange(bool cc_vars) {
    Nccl &n = Nccl::get();
    if (!peers.empty()) {
      if (!comm) Fatal(__FILE__, __LINE__, "MeshBlocks on other ranks but no NCCL communicator (call ak_sim_comm_init)");
      for (size_t p = 0; p < peers.size(); ++p) {
        if (cc_vars) AK_CHECK(nullptr, ak_halo_pack_cc(plan, a.u0, nvar, (int)p, peers[p].scc, st));
        else AK_CHECK(nullptr, ak_halo_pack_fc(plan, a.b0x1f, a.b0x2f, a.b0x3f, (int)p, peers[p].sfc, st));
        launches += 1;
      }
      n.GroupStart();
      for (auto &pe : peers) {
        const size_t cnt = cc_vars ? (size_t)nvar * pe.ncc : (size_t)pe.nfc;
        n.Send(cc_vars ? pe.scc : pe.sfc, cnt, kNcclFloat64, pe.rank, comm, st);
        n.Recv(cc_vars ? pe.rcc : pe.rfc, cnt, kNcclFloat64, pe.rank, comm, st);
      }
      n.GroupEnd();
    }
    if (cc_vars) AK_CHECK(nullptr, ak_halo_local_cc(plan, a.u0, nvar, st));
    else AK_CHECK(nullptr, ak_halo_local_fc(plan, a.b0x1f, a.b0x2f, a.b0x3f, st));
    launches += 1;
  }
  void Unpack(bool cc_vars) {
    for (size_t p = 0; p < peers.size(); ++p) {
      if (cc_vars) AK_CHECK(nullptr, ak_halo_unpack_cc(plan, a.u0, nvar, (int)p, peers[p].rcc, st));
      else AK_CHECK(nullptr, ak_halo_unpack_fc(plan, a.b0x1f, a.b0x2f, a.b0x3f, (int)p, peers[p].rfc, st));
      launches += 1;
    }
  }
  // U and B ghost exchange in one round: legal in the fused list because nothing runs between
  // RecvU and SendB there (CornerE/CT already happened); one ncclSend/ncclRecv pair per peer.
  // Peer-memory variant of one exchange (sequence number nexch+1): wait until every neighbour has
  // released its ghost zones, write my boundary slabs into them, publish "filled".
  void PushExchange(bool with_b) {
    const unsigned long long seq = nexch + 1;
    AK_CHECK(nullptr, ak_halo_wait(plan, 0, seq, st));
    if (u_pushed) {        // u0 went out early on the auxiliary stream (StageFluxEmf): face field only, then join
      AK_CHECK(nullptr, ak_halo_push(plan, nullptr, nvar, a.b0x1f, a.b0x2f, a.b0x3f, st));
      CU_CHECK(cudaStreamWaitEvent(st, ev_upush, 0));
      u_pushed = false;
    } else {
      AK_CHECK(nullptr, ak_halo_push(plan, a.u0, nvar, with_b ? a.b0x1f : nullptr, a.b0x2f, a.b0x3f, st));
    }
    AK_CHECK(nullptr, ak_halo_signal(plan, 1, seq, st));
    launches += 3;
  }
  void PushComplete() {   // receiver side: my ghost zones have been filled by every neighbour
    AK_CHECK(nullptr, ak_halo_wait(plan, 1, nexch + 1, st));
    launches += 1;
    nexch += 1;
  }
  bool UsePush() const { return push && fused && !peers.empty(); }
  TaskStatus SendUB(Driver *pd, int stage) {
    Nccl &n = Nccl::get();
    if (UsePush()) {
      PushExchange(true);
      AK_CHECK(nullptr, ak_halo_local_cc(plan, a.u0, nvar, st));
      AK_CHECK(nullptr, ak_halo_local_fc(plan, a.b0x1f, a.b0x2f, a.b0x3f, st));
      launches += 2;
      return TaskStatus::complete;
    }
    if (!peers.empty()) {
      if (!comm) Fatal(__FILE__, __LINE__, "MeshBlocks on other ranks but no NCCL communicator (call ak_sim_comm_init)");
      for (size_t p = 0; p < peers.size(); ++p) {
        AK_CHECK(nullptr, ak_halo_pack_cc(plan, a.u0, nvar, (int)p, peers[p].scc, st));
        AK_CHECK(nullptr, ak_halo_pack_fc(plan, a.b0x1f, a.b0x2f, a.b0x3f, (int)p, peers[p].sfc, st));
        launches += 2;
      }
      n.GroupStart();
      for (auto &pe : peers) {
        const size_t cnt = (size_t)nvar * pe.ncc + (size_t)pe.nfc;
        n.Send(pe.scc, cnt, kNcclFloat64, pe.rank, comm, st);
        n.Recv(pe.rcc, cnt, kNcclFloat64, pe.rank, comm, st);
      }
      n.GroupEnd();
    }
    AK_CHECK(nullptr, ak_halo_local_cc(plan, a.u0, nvar, st));
    AK_CHECK(nullptr, ak_halo_local_fc(plan, a.b0x1f, a.b0x2f, a.b0x3f, st));
    launches += 2;
    return TaskStatus::complete;
  }
  TaskStatus RecvUB(Driver *pd, int stage) {
    if (UsePush()) { PushComplete(); return TaskStatus::complete; }
    Unpack(true); Unpack(false);
    return TaskStatus::complete;
  }
  // SendFlux/RecvFlux (hydro_tasks.cpp:58-59 -> flux_correct_cc.cpp)
  TaskStatus FluxCorrect(Driver *pd, int stage) {
    if (ak_smr_flux_correct_cc(smr, a.flx1, a.flx2, a.flx3, nvar, st)) Fatal(__FILE__, __LINE__, ak_smr_last_error(smr));
    launches += 1;
    return TaskStatus::complete;
  }
  // RestrictU + SendU/RecvU across levels + Prolongate (hydro_tasks.cpp:63-68,296-398)
  void SmrGhosts() {
    auto ck = [&](int rc) { if (rc) Fatal(__FILE__, __LINE__, ak_smr_last_error(smr)); };
    ck(ak_smr_restrict_cc(smr, a.u0, coarse_u0, nvar, st));
    ck(ak_smr_exchange_cc(smr, a.u0, coarse_u0, nvar, st));
    if (!bcs) {
      ck(ak_smr_prolongate_cc(smr, a.u0, coarse_u0, nvar, st));
    } else {   // coarse-array BCs between the coarse fill and the prolongation (HydroBCsCoarse)
      ck(ak_smr_prolongate_cc_steps(smr, a.u0, coarse_u0, nvar, 1, st));
      if (ak_bcs_hydro(bcs_coarse, coarse_u0, nvar, st)) Fatal(__FILE__, __LINE__, ak_bcs_last_error(bcs_coarse));
      ck(ak_smr_prolongate_cc_steps(smr, a.u0, coarse_u0, nvar, 2, st));
      launches += nbc_launches;
    }
    launches += 4;
  }
  // MHD on a multilevel mesh: RestrictU + SendU/RecvU and RestrictB + SendB/RecvB across levels (mhd_tasks.cpp:55-57,
  // 63-65), then Prolongate (:527-551): coarse fills of U and B first, then both prolongations
  void SmrExchangeU() {
    auto ck = [&](int rc) { if (rc) Fatal(__FILE__, __LINE__, ak_smr_last_error(smr)); };
    ck(ak_smr_restrict_cc(smr, a.u0, coarse_u0, nvar, st));
    ck(ak_smr_exchange_cc(smr, a.u0, coarse_u0, nvar, st));
    launches += 2;
  }
  void SmrExchangeB() {
    auto ck = [&](int rc) { if (rc) Fatal(__FILE__, __LINE__, ak_smr_last_error(smr)); };
    ck(ak_smr_restrict_fc(smr, a.b0x1f, a.b0x2f, a.b0x3f, coarse_b0[0], coarse_b0[1], coarse_b0[2], st));
    ck(ak_smr_exchange_fc(smr, a.b0x1f, a.b0x2f, a.b0x3f, coarse_b0[0], coarse_b0[1], coarse_b0[2], st));
    launches += 2;
  }
  TaskStatus Prolongate(Driver *pd, int stage) {
    auto ck = [&](int rc) { if (rc) Fatal(__FILE__, __LINE__, ak_smr_last_error(smr)); };
    ck(ak_smr_prolongate_cc_steps(smr, a.u0, coarse_u0, nvar, 1, st));
    ck(ak_smr_prolongate_fc_steps(smr, a.b0x1f, a.b0x2f, a.b0x3f, coarse_b0[0], coarse_b0[1], coarse_b0[2], 1, st));
    if (bcs) {     // HydroBCsCoarse + BFieldBCsCoarse before the prolongation reads the coarse arrays (mhd_tasks.cpp:532-537)
      if (ak_bcs_apply(bcs_coarse, coarse_u0, nvar, coarse_b0[0], coarse_b0[1], coarse_b0[2], st)) Fatal(__FILE__, __LINE__, ak_bcs_last_error(bcs_coarse));
      launches += nbc_launches;
    }
    ck(ak_smr_prolongate_cc_steps(smr, a.u0, coarse_u0, nvar, 2, st));
    ck(ak_smr_prolongate_fc_steps(smr, a.b0x1f, a.b0x2f, a.b0x3f, coarse_b0[0], coarse_b0[1], coarse_b0[2], 2, st));
    launches += 5;
    return TaskStatus::complete;
  }
  TaskStatus SendU(Driver *pd, int stage) {
    if (smr && !hydro) { SmrExchangeU(); return TaskStatus::complete; }
    if (smr) { SmrGhosts(); return TaskStatus::complete; }
    if (hydro && UsePush()) {     // hydro twin, fused list: u0 is the only exchanged array
      PushExchange(false);
      AK_CHECK(nullptr, ak_halo_local_cc(plan, a.u0, nvar, st));
      launches += 1;
      return TaskStatus::complete;
    }
    Exchange(true);
    return TaskStatus::complete;
  }
  TaskStatus RecvU(Driver *pd, int stage) {   // stream-ordered: never incomplete
    if (smr) return TaskStatus::complete;
    if (hydro && UsePush()) { PushComplete(); return TaskStatus::complete; }
    Unpack(true);
    return TaskStatus::complete;
  }
  TaskStatus SendB(Driver *pd, int stage) {
    if (smr) { SmrExchangeB(); return TaskStatus::complete; }
    Exchange(false);
    return TaskStatus::complete;
  }
  TaskStatus RecvB(Driver *pd, int stage) {
    if (smr) return TaskStatus::complete;
    Unpack(false);
    return TaskStatus::complete;
  }
  // task 26, mhd_tasks.cpp:504-520: nothing to do on a strictly periodic mesh
  TaskStatus ApplyPhysicalBCs(Driver *pd, int stage) {
    if (!bcs) return TaskStatus::complete;
    if (hydro ? ak_bcs_hydro(bcs, a.u0, nvar, st) : ak_bcs_apply(bcs, a.u0, nvar, a.b0x1f, a.b0x2f, a.b0x3f, st))
      Fatal(__FILE__, __LINE__, ak_bcs_last_error(bcs));
    launches += nbc_launches;
    return TaskStatus::complete;
  }
  int nbc_launches = 0;
  // new dt over the active dimensions only (mhd_newdt.cpp:152-155, hydro_newdt.cpp:122-125)
  double MinDt(const double *dt3) const {
    double d = dt3[0];
    if (multi_d) d = std::min(d, dt3[1]);
    if (three_d) d = std::min(d, dt3[2]);
    return d;
  }
  TaskStatus ConToPrim(Driver *pd, int stage) {
    const int n1 = cfg.nx1 + 2 * cfg.ng, n2 = multi_d ? cfg.nx2 + 2 * cfg.ng : 1, n3 = three_d ? cfg.nx3 + 2 * cfg.ng : 1;
    int nf[3];
    if (hydro) AK_CHECK(h, ak_hydro_cons2prim(h, &a, 0, n1 - 1, 0, n2 - 1, 0, n3 - 1, nf, st));
    else AK_CHECK(h, ak_mhd_cons2prim(h, &a, 0, n1 - 1, 0, n2 - 1, 0, n3 - 1, nf, st));
    for (int q = 0; q < 3; ++q) nfloor[q] += nf[q];
    launches += 1;
    return TaskStatus::complete;
  }
  TaskStatus NewTimeStep(Driver *pd, int stage) {
    if (stage != pd->nexp_stages) return TaskStatus::complete;    // mhd_newdt.cpp:32-34
    double dt3[3];
    if (hydro) AK_CHECK(h, ak_hydro_new_dt(h, &a, dt3, st)); else AK_CHECK(h, ak_mhd_new_dt(h, &a, dt3, st));
    dtnew = MinDt(dt3);
    launches += 1;
    return TaskStatus::complete;
  }
  TaskStatus ConToPrimNewDt(Driver *pd, int stage) {              // tasks 27+28 fused on the last stage
    const int n1 = cfg.nx1 + 2 * cfg.ng, n2 = multi_d ? cfg.nx2 + 2 * cfg.ng : 1, n3 = three_d ? cfg.nx3 + 2 * cfg.ng : 1;
    if (stage != pd->nexp_stages) {
      if (hydro) AK_CHECK(h, ak_hydro_cons2prim(h, &a, 0, n1 - 1, 0, n2 - 1, 0, n3 - 1, nullptr, st));
      else AK_CHECK(h, ak_mhd_cons2prim(h, &a, 0, n1 - 1, 0, n2 - 1, 0, n3 - 1, nullptr, st));
    } else {
      int nf[3];
      double dt3[3];
      if (hydro) AK_CHECK(h, ak_hydro_cons2prim_newdt(h, &a, nf, dt3, st)); else AK_CHECK(h, ak_mhd_cons2prim_newdt(h, &a, nf, dt3, st));
      for (int q = 0; q < 3; ++q) nfloor[q] += nf[q];
      dtnew = MinDt(dt3);
    }
    launches += 1;
    return TaskStatus::complete;
  }

  // Mesh::NewTimeStep, mesh.cpp:573-643
  void MeshNewTimeStep(double tlim) {
    double dt = 2.0 * mesh.dt;
    dt = std::min(dt, mesh.cfl_no * dtnew);
    if (nranks > 1) {
      Nccl &n = Nccl::get();
      CU_CHECK(cudaMemcpyAsync(d_dt, &dt, 8, cudaMemcpyHostToDevice, st));
      n.AllReduce(d_dt, d_dt, 1, kNcclFloat64, kNcclMin, comm, st);
      AK_CHECK(nullptr, ak_small_d2h_sync(&dt, d_dt, 8, st));     // not cudaMemcpyAsync: see small_d2h in capi.cu
    }
    if ((mesh.time < tlim) && ((mesh.time + dt) > tlim)) dt = tlim - mesh.time;
    mesh.dt = dt;
  }
};

void Driver::Initialize(MHD *pmhd) {   // driver.cpp:314-371 -> InitBoundaryValuesAndPrimitives :569-653
  if (pmhd->UsePush()) {          // peer-memory exchange: same protocol as inside a stage
    pmhd->SignalFree();
    if (pmhd->hydro) { pmhd->SendU(this, 0); pmhd->RecvU(this, 0); }
    else { pmhd->SendUB(this, 0); pmhd->RecvUB(this, 0); }
  } else {
    pmhd->SendU(this, 0); pmhd->RecvU(this, 0);
    if (!pmhd->hydro) { pmhd->SendB(this, 0); pmhd->RecvB(this, 0); }
    if (!pmhd->hydro && pmhd->smr) pmhd->Prolongate(this, 0);     // driver.cpp:607-653 on a multilevel mesh
  }
  pmhd->ApplyPhysicalBCs(this, 0);                 // driver.cpp:626-627
  pmhd->ConToPrim(this, 0);
  pmhd->NewTimeStep(this, nexp_stages);
  pmhd->MeshNewTimeStep(tlim);
}

int Driver::Execute(MHD *pmhd, int ncycles) {   // driver.cpp:380-459
  int done = 0;
  auto &m = pmhd->mesh;
  static const bool trace = std::getenv("AKB_TRACE") && std::getenv("AKB_TRACE")[0] == '1';
  while ((m.time < tlim) && (m.ncycle < nlim || nlim < 0) && (ncycles < 0 || done < ncycles)) {
    if (trace && done == ncycles - 1) {       // last cycle of this call: event after every task body
      cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, pmhd->st);
      trace_events.assign(1, e);
      trace_stream = pmhd->st;
    }
    ExecuteTaskList(pmhd->tl_map, "before_timeintegrator", 0);
    for (int stage = 1; stage <= nexp_stages; ++stage) {
      ExecuteTaskList(pmhd->tl_map, "before_stagen", stage);
      ExecuteTaskList(pmhd->tl_map, "stagen", stage);
      ExecuteTaskList(pmhd->tl_map, "after_stagen", stage);
    }
    ExecuteTaskList(pmhd->tl_map, "after_timeintegrator", 1);
    if (trace_stream) {
      cudaStreamSynchronize(trace_stream);
      std::fprintf(stderr, "AKB_TRACE cycle %d: ms after each task body (stage-major):", m.ncycle);
      float tot = 0.f;
      for (size_t q = 1; q < trace_events.size(); ++q) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, trace_events[q - 1], trace_events[q]);
        tot += ms;
        std::fprintf(stderr, " %.3f", ms);
      }
      std::fprintf(stderr, "  total %.3f\n", tot);
      for (auto e : trace_events) cudaEventDestroy(e);
      trace_events.clear();
      trace_stream = nullptr;
    }
    m.time = m.time + m.dt;
    m.ncycle++;
    nmb_updated_ += m.nmb_total;
    done++;
    pmhd->MeshNewTimeStep(tlim);
  }
  return done;
}

struct Sim {
  std::unique_ptr<MHD> pmhd;
  std::unique_ptr<Driver> pdriver;
  ak_sim_cfg cfg;                  // as given to ak_sim_create: an AMR mesh change re-creates the state owner from it
  long nadapt = 0, ncreated = 0, ndeleted = 0;
};

}  // namespace akh

extern "C" {

int ak_sim_create(const ak_sim_cfg *c, const double *mb_size, const int *nghbr, void **sim) {
  if (!c || !sim || !mb_size || !nghbr) return 1;
  auto *s = new akh::Sim();
  s->cfg = *c;
  s->pmhd.reset(new akh::MHD(c->mhd, mb_size, nghbr, c->my_rank, c->nranks, c->gids_start, c->nmb_total));
  s->pmhd->fused = c->fused != 0;
  s->pmhd->mesh.cfl_no = c->cfl_number;
  s->pmhd->AssembleMHDTasks();
  s->pdriver.reset(new akh::Driver(c->integrator == 1 ? "rk1" : c->integrator == 2 ? "rk2" : "rk3", c->tlim, c->nlim));
  *sim = s;
  return 0;
}
/* physical boundary flags of the mesh and of this rank's blocks (default: strictly periodic) */
int ak_sim_set_bcs(void *sim, const int *mesh_bcs, const int *mb_bcs) {
  auto *m = static_cast<akh::Sim *>(sim)->pmhd.get();
  ak_bcs_destroy(m->bcs);
  m->bcs = nullptr; m->nbc_launches = 0;
  void *b = nullptr;
  if (ak_bcs_create(&m->cfg, mesh_bcs, mb_bcs, &b)) return 1;
  if (ak_bcs_strictly_periodic(b)) { ak_bcs_destroy(b); return 0; }
  m->bcs = b;
  std::memcpy(m->mesh_bcs_h, mesh_bcs, sizeof(m->mesh_bcs_h));
  m->mb_bcs_h.assign(mb_bcs, mb_bcs + (size_t)m->cfg.nmb * 6);
  const bool act[3] = {true, m->multi_d, m->three_d};
  for (int d = 0; d < 3; ++d) m->nbc_launches += (act[d] && mesh_bcs[2 * d] != AK_BC_PERIODIC);
  return 0;
}
/* static mesh refinement (hydro and MHD, one rank): logical level of every block of the pack (ak_mesh_create_smr).  Builds the
 * multi-level plan from the neighbour table given to ak_sim_create, allocates coarse_u0 and switches to the multilevel
 * task list.  Call before ak_sim_initialize; with physical boundaries call ak_sim_set_bcs first. */
int ak_sim_set_levels(void *sim, const int *mb_lev) {
  auto *m = static_cast<akh::Sim *>(sim)->pmhd.get();
  if (m->nranks != 1 || !mb_lev || m->smr) return 1;
  if (!m->hydro && (m->cfg.fofc || m->cfg.nscalars)) return 1;
  if (ak_smr_create(&m->cfg, m->nghbr_h.data(), mb_lev, m->my_rank, m->gids_start_, &m->smr)) return 1;
  const size_t n = (size_t)m->cfg.nmb * m->nvar * ak_smr_coarse_cells(m->smr);
  if (cudaMalloc(&m->coarse_u0, n * 8) != cudaSuccess || cudaMemsetAsync(m->coarse_u0, 0, n * 8, m->st) != cudaSuccess) return 1;
  m->owned.push_back(m->coarse_u0);
  if (!m->hydro)
    for (int v = 0; v < 3; ++v) {
      const size_t nf = (size_t)m->cfg.nmb * ak_smr_coarse_faces(m->smr, v);
      if (cudaMalloc(&m->coarse_b0[v], nf * 8) != cudaSuccess || cudaMemsetAsync(m->coarse_b0[v], 0, nf * 8, m->st) != cudaSuccess) return 1;
      m->owned.push_back(m->coarse_b0[v]);
    }
  if (m->bcs) {      // HydroBCsCoarse: the same flags on the coarse index space
    ak_mhd_cfg cc = m->cfg;
    cc.nx1 = m->cfg.nx1 / 2; cc.nx2 = std::max(1, m->cfg.nx2 / 2); cc.nx3 = std::max(1, m->cfg.nx3 / 2);
    if (ak_bcs_create(&cc, m->mesh_bcs_h, m->mb_bcs_h.data(), &m->bcs_coarse)) return 1;
  }
  m->tl_map.clear();
  m->AssembleMHDTasks();
  return 0;
}

/* ---- adaptive mesh refinement inside the C++ host driver (one rank) ----------------------------------------------------
 * MeshRefinement::AdaptiveMeshRefinement (src/mesh/mesh_refinement.cpp:145-190) as Driver::Execute calls it after every
 * cycle (src/driver/driver.cpp:447-450): criteria on the device (ak_amr_criterion) -> ak_mesh_adapt on the host mesh ->
 * a new state owner for the new block list (arrays, ghost-exchange and multilevel plans) -> ak_amr_plan + ak_amr_move
 * (every element of the new arrays gathered from the old ones) -> ghost zones + primitives -> ak_amr_repair_fc -> ghost
 * zones + primitives again -> new dt.  mesh = the adaptive host mesh (ak_mesh_create[_smr] + ak_mesh_set_amr) whose tables
 * the simulation was created from; crit = ncrit rows {method, use_primitive, value_min, value_max, loc x1, x2, x3, rad}. */
static int sim_init_bvp(akh::Sim *s) {       // Driver::InitBoundaryValuesAndPrimitives, driver.cpp:569-653
  akh::MHD *m = s->pmhd.get();
  akh::Driver *d = s->pdriver.get();
  m->SendU(d, 0); m->RecvU(d, 0);
  if (!m->hydro) { m->SendB(d, 0); m->RecvB(d, 0); if (m->smr) m->Prolongate(d, 0); }
  m->ApplyPhysicalBCs(d, 0);
  m->ConToPrim(d, 0);
  return 0;
}

int ak_sim_adapt(void *sim, void *mesh, const double *crit, int ncrit, int *changed) {
  auto *s = static_cast<akh::Sim *>(sim);
  akh::MHD *old = s->pmhd.get();
  if (changed) *changed = 0;
  if (!mesh || (ncrit > 0 && !crit)) return 1;
  if (old->nranks != 1) return 1;                       // block migration across ranks is not built
  cudaSetDevice(old->device);
  const int onmb = old->cfg.nmb;
  std::vector<int> flag(onmb, 0);
  for (int q = 0; q < ncrit; ++q) {
    const double *c = crit + 8 * q;
    const double *field = (c[1] != 0.0) ? old->a.w0 : old->a.u0;
    if (ak_amr_criterion(&old->cfg, (int)c[0], field, old->nvar, 0, c[2], c[3], c + 4, old->mb_size_h.data(), flag.data(), old->st)) return 1;
    old->launches += 1;
  }
  int nnew = 0, ndel = 0;
  if (ak_mesh_adapt(mesh, old->mesh.ncycle, flag.data(), &nnew, &ndel)) return 1;
  if (!nnew && !ndel) return 0;
  // tables of the NEW mesh
  int nmb = 0, root = 0;
  if (ak_mesh_info(mesh, &nmb, &root)) return 1;
  std::vector<int> lloc(4 * (size_t)nmb), rank_of(nmb), gs(1), nr(1);
  if (ak_mesh_tables(mesh, lloc.data(), rank_of.data(), gs.data(), nr.data())) return 1;
  std::vector<double> mbs(9 * (size_t)nmb);
  std::vector<int> bcs(6 * (size_t)nmb), ngh(56 * 4 * (size_t)nmb);
  if (ak_mesh_rank_tables(mesh, 0, mbs.data(), bcs.data(), ngh.data())) return 1;
  std::vector<int> n2o(nmb), o2n(onmb), oflag(onmb);
  int old_nmb = 0;
  if (ak_mesh_amr_maps(mesh, n2o.data(), o2n.data(), oflag.data(), &old_nmb) || old_nmb != onmb) return 1;
  // new state owner
  ak_sim_cfg nc = s->cfg;
  nc.mhd.nmb = nmb; nc.nmb_total = nmb;
  std::unique_ptr<akh::MHD> nw(new akh::MHD(nc.mhd, mbs.data(), ngh.data(), nc.my_rank, nc.nranks, nc.gids_start, nmb));
  nw->fused = old->fused;
  nw->mesh = old->mesh;
  nw->mesh.nmb_total = nmb;
  nw->launches = old->launches;
  for (int q = 0; q < 3; ++q) nw->nfloor[q] = old->nfloor[q];
  nw->nfofc = old->nfofc;
  nw->dtnew = old->dtnew;
  std::unique_ptr<akh::MHD> keep = std::move(s->pmhd);
  s->pmhd = std::move(nw);
  akh::MHD *m = s->pmhd.get();
  if (keep->bcs && ak_sim_set_bcs(sim, keep->mesh_bcs_h, bcs.data())) return 1;
  std::vector<int> lev(nmb);
  for (int q = 0; q < nmb; ++q) lev[q] = lloc[4 * q + 3];
  if (ak_sim_set_levels(sim, lev.data())) return 1;
  // data movement
  std::vector<int> plan(5 * (size_t)nmb);
  if (ak_amr_plan(&m->cfg, onmb, n2o.data(), oflag.data(), lloc.data(), plan.data())) return 1;
  if (ak_amr_move(&m->cfg, plan.data(), keep->a.u0, m->a.u0, m->nvar, keep->a.b0x1f, keep->a.b0x2f, keep->a.b0x3f,
                  m->a.b0x1f, m->a.b0x2f, m->a.b0x3f, m->st)) return 1;
  m->launches += m->hydro ? 1 : 4;
  if (cudaStreamSynchronize(m->st) != cudaSuccess) return 1;       // the old arrays go away with `keep`
  keep.reset();
  sim_init_bvp(s);
  if (!m->hydro) {      // mesh_refinement.cpp:164-171: MHD repairs the interior faces of new fine blocks, then rebuilds the ghost zones
    std::vector<int> rep(nmb);
    for (int q = 0; q < nmb; ++q) rep[q] = oflag[n2o[q]] != 0;
    if (ak_amr_repair_fc(&m->cfg, rep.data(), m->a.b0x1f, m->a.b0x2f, m->a.b0x3f, m->st)) return 1;
    m->launches += 1;
    sim_init_bvp(s);
  }
  m->NewTimeStep(s->pdriver.get(), s->pdriver->nexp_stages);
  s->nadapt += 1; s->ncreated += nnew; s->ndeleted += ndel;
  if (changed) *changed = 1;
  return 0;
}

/* Driver::Execute with AMR (driver.cpp:380-459): cycle -> AdaptiveMeshRefinement -> Mesh::NewTimeStep.  Returns the cycles done. */
int ak_sim_execute_amr(void *sim, void *mesh, const double *crit, int ncrit, int ncycles) {
  auto *s = static_cast<akh::Sim *>(sim);
  akh::Driver *d = s->pdriver.get();
  cudaSetDevice(s->pmhd->device);
  int done = 0;
  while ((s->pmhd->mesh.time < d->tlim) && (s->pmhd->mesh.ncycle < d->nlim || d->nlim < 0) && (ncycles < 0 || done < ncycles)) {
    akh::MHD *m = s->pmhd.get();
    for (int stage = 1; stage <= d->nexp_stages; ++stage) {
      d->ExecuteTaskList(m->tl_map, "before_stagen", stage);
      d->ExecuteTaskList(m->tl_map, "stagen", stage);
      d->ExecuteTaskList(m->tl_map, "after_stagen", stage);
    }
    m->mesh.time = m->mesh.time + m->mesh.dt;
    m->mesh.ncycle++;
    done++;
    int changed = 0;
    if (ak_sim_adapt(sim, mesh, crit, ncrit, &changed)) return -1;
    s->pmhd->MeshNewTimeStep(d->tlim);
  }
  return done;
}
/* restart-style entry: the state just loaded with ak_sim_set_state belongs to (time, dt, ncycle) */
int ak_sim_set_time(void *sim, double time, double dt, int ncycle) {
  auto *m = static_cast<akh::Sim *>(sim)->pmhd.get();
  m->mesh.time = time; m->mesh.dt = dt; m->mesh.ncycle = ncycle;
  return 0;
}
/* ConToPrim over every cell of the loaded state (ghost zones as loaded) + MHD::NewTimeStep: what the last stage of a cycle leaves */
int ak_sim_prims_newdt(void *sim) {
  auto *s = static_cast<akh::Sim *>(sim);
  cudaSetDevice(s->pmhd->device);
  s->pmhd->ConToPrim(s->pdriver.get(), 0);
  s->pmhd->NewTimeStep(s->pdriver.get(), s->pdriver->nexp_stages);
  return 0;
}
/* Mesh::NewTimeStep (mesh.cpp:573-643): dt = min(2 dt, cfl * dtnew), clipped at tlim */
int ak_sim_mesh_new_dt(void *sim) {
  auto *s = static_cast<akh::Sim *>(sim);
  s->pmhd->MeshNewTimeStep(s->pdriver->tlim);
  return 0;
}
/* cnt[0..3] = MeshBlocks now, mesh changes so far, blocks created, blocks deleted by refinement */
int ak_sim_amr_status(void *sim, long *cnt) {
  auto *s = static_cast<akh::Sim *>(sim);
  cnt[0] = s->pmhd->cfg.nmb; cnt[1] = s->nadapt; cnt[2] = s->ncreated; cnt[3] = s->ndeleted;
  return 0;
}
int ak_sim_destroy(void *sim) { delete static_cast<akh::Sim *>(sim); return 0; }

int ak_nccl_unique_id(void *out128) {
  akh::Nccl &n = akh::Nccl::get();
  if (!n.ok) return 1;
  return n.GetUniqueId(static_cast<akh::Nccl::UniqueId *>(out128));
}
int ak_sim_comm_init(void *sim, const void *unique_id128) {
  auto *s = static_cast<akh::Sim *>(sim);
  akh::Nccl &n = akh::Nccl::get();
  if (!n.ok) return 1;
  akh::Nccl::UniqueId id;
  std::memcpy(&id, unique_id128, sizeof(id));
  if (n.CommInitRank(&s->pmhd->comm, s->pmhd->nranks, id, s->pmhd->my_rank)) return 1;
  s->pmhd->SetupPeerPush(s->pmhd->gids_start_);
  return 0;
}
int ak_sim_set_state(void *sim, const double *u0, const double *b1f, const double *b2f, const double *b3f) {
  auto *m = static_cast<akh::Sim *>(sim)->pmhd.get();
  cudaSetDevice(m->device);      // a host thread other than the creating one may drive this handle
  cudaError_t e = cudaMemcpyAsync(m->a.u0, u0, m->n_u * 8, cudaMemcpyHostToDevice, m->st);
  if (m->hydro) return e == cudaSuccess ? 0 : 1;     // hydro twin: u0 only, the face-field pointers are ignored
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->a.b0x1f, b1f, m->n_f1 * 8, cudaMemcpyHostToDevice, m->st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->a.b0x2f, b2f, m->n_f2 * 8, cudaMemcpyHostToDevice, m->st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->a.b0x3f, b3f, m->n_f3 * 8, cudaMemcpyHostToDevice, m->st);
  return e == cudaSuccess ? 0 : 1;
}
int ak_sim_get_state(void *sim, double *u0, double *b1f, double *b2f, double *b3f) {
  auto *m = static_cast<akh::Sim *>(sim)->pmhd.get();
  cudaSetDevice(m->device);
  cudaError_t e = cudaMemcpyAsync(u0, m->a.u0, m->n_u * 8, cudaMemcpyDeviceToHost, m->st);
  if (m->hydro) { if (e == cudaSuccess) e = cudaStreamSynchronize(m->st); return e == cudaSuccess ? 0 : 1; }
  if (e == cudaSuccess) e = cudaMemcpyAsync(b1f, m->a.b0x1f, m->n_f1 * 8, cudaMemcpyDeviceToHost, m->st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(b2f, m->a.b0x2f, m->n_f2 * 8, cudaMemcpyDeviceToHost, m->st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(b3f, m->a.b0x3f, m->n_f3 * 8, cudaMemcpyDeviceToHost, m->st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(m->st);
  return e == cudaSuccess ? 0 : 1;
}
/* primitives and cell-centred field of the current state, ghost zones included (test / diagnostic read-back) */
int ak_sim_get_prim(void *sim, double *w0, double *bcc0) {
  auto *m = static_cast<akh::Sim *>(sim)->pmhd.get();
  cudaSetDevice(m->device);
  cudaError_t e = cudaMemcpyAsync(w0, m->a.w0, m->n_u * 8, cudaMemcpyDeviceToHost, m->st);
  if (e == cudaSuccess && !m->hydro && bcc0)
    e = cudaMemcpyAsync(bcc0, m->a.bcc0, (m->n_u / m->nvar) * 3 * 8, cudaMemcpyDeviceToHost, m->st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(m->st);
  return e == cudaSuccess ? 0 : 1;
}
int ak_sim_initialize(void *sim) {
  auto *s = static_cast<akh::Sim *>(sim);
  cudaSetDevice(s->pmhd->device);
  s->pmhd->mesh.time = 0.0; s->pmhd->mesh.ncycle = 0; s->pmhd->mesh.dt = (double)FLT_MAX; s->pmhd->dtnew = (double)FLT_MAX;
  s->pdriver->Initialize(s->pmhd.get());
  return 0;
}
int ak_sim_execute(void *sim, int ncycles) {
  auto *s = static_cast<akh::Sim *>(sim);
  cudaSetDevice(s->pmhd->device);
  return s->pdriver->Execute(s->pmhd.get(), ncycles);
}
int ak_sim_sync(void *sim) {
  cudaSetDevice(static_cast<akh::Sim *>(sim)->pmhd->device);
  return cudaStreamSynchronize(static_cast<akh::Sim *>(sim)->pmhd->st) == cudaSuccess ? 0 : 1;
}
/* st[0..2] = time, dt, dtnew; cnt[0..7] = ncycle, launches, nfloor d/e/t, stream handle (as integer),
 * peer-memory ghost exchange active (0/1), exchanges done through it */
int ak_sim_status(void *sim, double *st, long *cnt) {
  auto *m = static_cast<akh::Sim *>(sim)->pmhd.get();
  st[0] = m->mesh.time; st[1] = m->mesh.dt; st[2] = m->dtnew;
  cnt[0] = m->mesh.ncycle; cnt[1] = m->launches; cnt[2] = m->nfloor[0]; cnt[3] = m->nfloor[1]; cnt[4] = m->nfloor[2];
  cnt[5] = (long)(intptr_t)m->st;
  cnt[6] = m->push ? 1 : 0;
  cnt[7] = (long)m->nexch;
  return 0;
}
int ak_sim_arrays(void *sim, ak_mhd_arrays *out) {
  *out = static_cast<akh::Sim *>(sim)->pmhd->a;
  return 0;
}

}  // extern "C"

/* ---- pipelined host-buffer stepping -----------------------------------------------------------------------------------------
 * The reference-facing call with HOST buffers is  set_state (H2D) -> initialize -> execute(n) -> get_state (D2H).  Issued back to
 * back on one engine instance the three phases serialise (PCIe idle while the SMs work and the other way round); several
 * instances driven by independent threads drift into lockstep (all uploading, then all computing, then all downloading) and share
 * the link instead of overlapping.  The pipe owns ONE upload stream and ONE download stream for all its engine instances (slots):
 * copies of one direction queue behind each other at full link speed, uploads, steps and downloads of consecutive jobs overlap,
 * and both PCIe directions stay busy.  A job is exactly the serial sequence above on its slot; jobs are independent of each other
 * (job k runs on slot k mod nslots, each slot has a worker thread because Driver::Execute reads dt back every cycle). */
namespace akh {
struct Pipe {
  struct Job { const double *in[4]; double *out[4]; int ncycles; long ticket; };
  struct Slot {
    Sim *sim = nullptr;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Job> q;
    bool stop = false;
    long completed = -1;        // last ticket of this slot whose result is in its host buffers
    int rc = 0;                 // sticky error
    cudaEvent_t up = nullptr, comp = nullptr, down = nullptr;
    cudaEvent_t t_up0 = nullptr, t_down0 = nullptr;      // AKB_PIPE_TRACE=1: when each copy phase actually started
  };
  std::vector<std::unique_ptr<Slot>> slots;
  cudaStream_t h2d = nullptr, d2h = nullptr;
  int device = 0;
  long next_ticket = 0;
  bool trace = false;
  cudaEvent_t epoch = nullptr;
  std::mutex tmu;
  // Turnstiles: uploads are queued and steps are run strictly in job order.  Steps of different slots left to run side by
  // side only time-slice the SMs (every job finishes late, every slot stays occupied longer); and on several ranks the same
  // job must be stepping everywhere at the same time, because its ghost exchange pairs the slots of equal index.
  struct Turn {
    std::mutex mu;
    std::condition_variable cv;
    long serving = 0;
    void enter(long t) { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return serving == t; }); }
    void leave() { { std::lock_guard<std::mutex> lk(mu); ++serving; } cv.notify_all(); }
  } up_turn, step_turn;

  static bool ok(cudaError_t e) { return e == cudaSuccess; }
  int RunJob(Slot &sl, const Job &j) {
    MHD *m = sl.sim->pmhd.get();
    const size_t nb[4] = {(size_t)m->n_u * 8, (size_t)m->n_f1 * 8, (size_t)m->n_f2 * 8, (size_t)m->n_f3 * 8};
    double *dev[4] = {m->a.u0, m->a.b0x1f, m->a.b0x2f, m->a.b0x3f};
    const int nf = m->hydro ? 1 : 4;
    // upload: the slot's previous download has completed (the worker waited for it), so the state arrays are free
    int rc = 0;
    up_turn.enter(j.ticket);
    if (trace) cudaEventRecord(sl.t_up0, h2d);
    for (int q = 0; q < nf && !rc; ++q)
      if (!ok(cudaMemcpyAsync(dev[q], j.in[q], nb[q], cudaMemcpyHostToDevice, h2d))) rc = 1;
    if (!rc && (!ok(cudaEventRecord(sl.up, h2d)) || !ok(cudaStreamWaitEvent(m->st, sl.up, 0)))) rc = 1;
    up_turn.leave();
    step_turn.enter(j.ticket);          // a failed job still takes and passes its turns, or every later job would wait for ever
    if (!rc && ak_sim_initialize(sl.sim)) rc = 1;
    if (!rc && ak_sim_execute(sl.sim, j.ncycles) != j.ncycles) rc = 1;
    step_turn.leave();
    if (rc) return rc;
    if (!ok(cudaEventRecord(sl.comp, m->st)) || !ok(cudaStreamWaitEvent(d2h, sl.comp, 0))) return 1;
    if (trace) cudaEventRecord(sl.t_down0, d2h);
    for (int q = 0; q < nf; ++q)
      if (!ok(cudaMemcpyAsync(j.out[q], dev[q], nb[q], cudaMemcpyDeviceToHost, d2h))) return 1;
    if (!ok(cudaEventRecord(sl.down, d2h)) || !ok(cudaEventSynchronize(sl.down))) return 1;
    if (trace) {
      float t[5] = {0, 0, 0, 0, 0};
      cudaEvent_t ev[5] = {sl.t_up0, sl.up, sl.comp, sl.t_down0, sl.down};
      for (int q = 0; q < 5; ++q) cudaEventElapsedTime(&t[q], epoch, ev[q]);
      std::lock_guard<std::mutex> lk(tmu);
      std::fprintf(stderr, "AKB_PIPE_TRACE job %ld slot %ld: h2d %.2f-%.2f  step done %.2f  d2h %.2f-%.2f ms\n", j.ticket,
                   j.ticket % (long)slots.size(), t[0], t[1], t[2], t[3], t[4]);
    }
    return 0;
  }
  void Work(Slot *sl) {
    cudaSetDevice(device);
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> lk(sl->mu);
        sl->cv.wait(lk, [&] { return sl->stop || !sl->q.empty(); });
        if (sl->q.empty()) return;
        j = sl->q.front();
      }
      const int rc = RunJob(*sl, j);
      {
        std::lock_guard<std::mutex> lk(sl->mu);
        sl->q.pop_front();
        sl->completed = j.ticket;
        if (rc) sl->rc = rc;
      }
      sl->cv.notify_all();
    }
  }
  ~Pipe() {
    for (auto &sl : slots) {
      { std::lock_guard<std::mutex> lk(sl->mu); sl->stop = true; }
      sl->cv.notify_all();
      if (sl->th.joinable()) sl->th.join();       // drains the slot's queue first
      cudaSetDevice(device);
      for (cudaEvent_t e : {sl->up, sl->comp, sl->down, sl->t_up0, sl->t_down0}) if (e) cudaEventDestroy(e);
    }
    if (h2d) cudaStreamDestroy(h2d);
    if (d2h) cudaStreamDestroy(d2h);
  }
};
}  // namespace akh

extern "C" {

int ak_pipe_create(void *const *sims, int nsims, void **pipe) {
  if (!sims || nsims < 1 || !pipe) return 1;
  std::unique_ptr<akh::Pipe> p(new akh::Pipe());
  p->device = static_cast<akh::Sim *>(sims[0])->pmhd->device;
  for (int q = 0; q < nsims; ++q)
    if (!sims[q] || static_cast<akh::Sim *>(sims[q])->pmhd->device != p->device) return 1;
  cudaSetDevice(p->device);
  if (cudaStreamCreateWithFlags(&p->h2d, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&p->d2h, cudaStreamNonBlocking) != cudaSuccess) return 1;
  p->trace = std::getenv("AKB_PIPE_TRACE") && std::getenv("AKB_PIPE_TRACE")[0] == '1';
  if (p->trace) { cudaEventCreate(&p->epoch); cudaEventRecord(p->epoch, p->h2d); }
  for (int q = 0; q < nsims; ++q) {
    std::unique_ptr<akh::Pipe::Slot> sl(new akh::Pipe::Slot());
    sl->sim = static_cast<akh::Sim *>(sims[q]);
    for (cudaEvent_t *e : {&sl->up, &sl->comp, &sl->down, &sl->t_up0, &sl->t_down0})
      if (cudaEventCreateWithFlags(e, p->trace ? cudaEventDefault : cudaEventDisableTiming) != cudaSuccess) return 1;
    p->slots.push_back(std::move(sl));
  }
  for (auto &sl : p->slots) sl->th = std::thread(&akh::Pipe::Work, p.get(), sl.get());
  *pipe = p.release();
  return 0;
}
/* One job: the HOST arrays of (u0, b1f, b2f, b3f) in, n cycles from time 0 (ak_sim_initialize + ak_sim_execute), the HOST arrays
 * of the result out.  Buffers must stay valid (and should be pinned) until ak_pipe_wait(ticket) returns.  Returns the ticket. */
long ak_pipe_submit(void *pipe, const double *u0, const double *b1f, const double *b2f, const double *b3f,
                    double *ou0, double *ob1f, double *ob2f, double *ob3f, int ncycles) {
  auto *p = static_cast<akh::Pipe *>(pipe);
  if (!p || !u0 || !ou0 || ncycles < 0) return -1;
  const long t = p->next_ticket++;
  auto &sl = *p->slots[(size_t)(t % (long)p->slots.size())];
  {
    std::lock_guard<std::mutex> lk(sl.mu);
    sl.q.push_back(akh::Pipe::Job{{u0, b1f, b2f, b3f}, {ou0, ob1f, ob2f, ob3f}, ncycles, t});
  }
  sl.cv.notify_all();
  return t;
}
/* blocks until the job's result is in its host buffers; ticket < 0: every job submitted so far */
int ak_pipe_wait(void *pipe, long ticket) {
  auto *p = static_cast<akh::Pipe *>(pipe);
  if (!p) return 1;
  int rc = 0;
  const long ns = (long)p->slots.size();
  for (long s = 0; s < ns; ++s) {
    long want;
    if (ticket >= 0) { if (ticket % ns != s) continue; want = ticket; }
    else { want = p->next_ticket - 1; while (want >= 0 && want % ns != s) --want; }
    auto &sl = *p->slots[(size_t)s];
    std::unique_lock<std::mutex> lk(sl.mu);
    sl.cv.wait(lk, [&] { return sl.completed >= want; });
    rc |= sl.rc;
  }
  return rc;
}
int ak_pipe_streams(void *pipe, ak_stream_t *h2d, ak_stream_t *d2h) {
  auto *p = static_cast<akh::Pipe *>(pipe);
  if (h2d) *h2d = p->h2d;
  if (d2h) *d2h = p->d2h;
  return 0;
}
int ak_pipe_destroy(void *pipe) { delete static_cast<akh::Pipe *>(pipe); return 0; }

}  // extern "C"
