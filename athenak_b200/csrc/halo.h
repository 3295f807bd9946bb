// halo.h -- internal declarations of the ghost-zone exchange plan (see halo.cu).
#pragma once
#include <cuda_runtime.h>
#include <map>
#include <string>
#include <utility>
#include <vector>
#include "../../include/athenak_b200.h"

namespace akb {

enum { HALO_LOCAL = 0, HALO_PACK = 1, HALO_UNPACK = 2 };

struct HaloDesc {
  int m, src_m, comp;      // block written/read; source block (local gather); -1 = CC, 0..2 = x1f..x3f
  int lo[3], hi[3];        // i,j,k box in the array named by comp
  int shift[3];            // local gather: src index = dst index + shift
  long boff;               // message offset (CC: per variable, scaled by nvar in the kernel)
};

struct HaloArrays {
  double *p[3];
  long ext1[3], ext2[3], bstride[3];  // inner, middle extents and per-block stride
};

// EMF flux correction (SendE/RecvE): one descriptor per region of boundary edges that is
// summed and scaled together (a face interior, or a block-edge line)
struct EmfSumDesc {
  int m, v;                 // block, edge-field component (0 x1e, 1 x2e, 2 x3e)
  int lo[3], hi[3];         // i,j,k box of the region in the edge array
  int nsrc;                 // contributing receive segments, in increasing slot order
  long soff[3];             // their offsets in the receive buffer
  int slo[3][3], sni[3], snj[3];   // their boxes: low corner and i/j extents (linearisation)
  int count;                // 1 + nsrc (nflx)
  int face;                 // 1: face interior (x0.5), 0: block edge (/count)
};

// Peer-memory exchange ("push"): one boundary slab of one of MY blocks written straight into the
// ghost zones of the neighbouring block in the PEER GPU's array (no pack buffer, no message, no
// unpack pass).  Source box = the pack box, destination box = the peer's unpack box.
struct PushDesc {
  int m, comp, peer;       // my block; -1 = CC, 0..2 = x1f..x3f; index into the peer pointer table
  int dst_gid, dst_m;      // neighbour's gid; its index in the peer's pack (set by push setup)
  int lo[3], hi[3];        // i,j,k source box
  int dlo[3];              // low corner of the destination box (same extents)
};
#define AKB_MAX_PUSH_PEERS 32
struct PushPeerPtrs { double *p[AKB_MAX_PUSH_PEERS][4]; };   // [peer][u0, x1f, x2f, x3f] mapped peer arrays

struct HaloPeer {
  int rank;
  long cc_count, fc_count;  // doubles per message: CC per variable, FC total
  std::vector<HaloDesc> pack_cc, unpack_cc, pack_fc, unpack_fc;
  HaloDesc *d_pack_cc = nullptr, *d_unpack_cc = nullptr, *d_pack_fc = nullptr, *d_unpack_fc = nullptr;
  std::vector<HaloDesc> pack_emf;      // my boundary EMFs -> message to this peer
  HaloDesc *d_pack_emf = nullptr;
  long emf_count = 0, emf_recv_base = 0;   // message length; where the peer's message lands in emf_rbuf
};

struct HaloPlan {
  ak_mhd_cfg cfg;
  std::vector<HaloDesc> local_cc, local_fc;
  HaloDesc *d_local_cc = nullptr, *d_local_fc = nullptr;
  std::vector<HaloPeer> peers;
  std::vector<HaloDesc> emf_pack_local;   // same-rank: straight into the receiver's segment of emf_rbuf
  std::vector<EmfSumDesc> emf_sum;
  HaloDesc *d_emf_pack_local = nullptr;
  EmfSumDesc *d_emf_sum = nullptr;
  long emf_rbuf_count = 0;
  double *d_emf_rbuf = nullptr;
  // peer-memory exchange state
  std::vector<PushDesc> push;             // all peers, CC entries first then FC
  PushDesc *d_push = nullptr;
  long n_push_cc = 0;
  PushPeerPtrs push_ptrs{};
  bool push_ready = false;
  unsigned long long *flags_local = nullptr;                 // [2][nranks]: free / ready counters written by peers
  unsigned long long **d_sig[4] = {nullptr, nullptr, nullptr, nullptr};   // per flag kind: device array of remote slot addresses
  int *d_wait_slots = nullptr;            // peer ranks to wait for
  int my_rank = 0, nranks = 1;
  std::string err;
  bool uploaded = false;   // descriptor lists are copied to the device on first use
  // Launch tables: a descriptor list is launched as ONE compact 1-D grid of tiles {descriptor, chunk of HALO_TILE elements},
  // built on first use per (list, nvar).  (Launching max-box-size CTAs for every descriptor, as round 1 did, starts hundreds of
  // thousands of empty CTAs when a pack holds many small blocks: 1.0 ms per face-field exchange for 64 blocks of 64^3.)
  struct TileTab { int2 *d = nullptr; int n = 0; };
  std::map<std::pair<const void *, int>, TileTab> tiles;
};
#define HALO_TILE 2048

int neighbor_index(int ix, int iy, int iz, int n1, int n2);
HaloPlan *halo_build(const ak_mhd_cfg &cfg, const int *nghbr, int my_rank, int gids_start, std::string &err);
void halo_free(HaloPlan *P);
cudaError_t halo_ensure_uploaded(HaloPlan *P);
cudaError_t halo_launch(HaloPlan *P, const std::vector<HaloDesc> &host, const HaloDesc *dev, const HaloArrays &A, int nvar,
                        long varstride, double *buf, int mode, cudaStream_t st);
HaloArrays halo_arrays_cc(const ak_mhd_cfg &c, double *a, int nvar);
HaloArrays halo_arrays_fc(const ak_mhd_cfg &c, double *x1f, double *x2f, double *x3f);
HaloArrays halo_arrays_ec(const ak_mhd_cfg &c, double *x1e, double *x2e, double *x3e);
cudaError_t emf_sumavg_launch(HaloPlan *P, const HaloArrays &A, cudaStream_t st);
cudaError_t halo_push_launch(HaloPlan *P, const HaloArrays &cc, int nvar, long varstride, const HaloArrays &fc, bool do_cc,
                             bool do_fc, cudaStream_t st);
cudaError_t peer_signal_launch(unsigned long long **slots, int n, unsigned long long value, cudaStream_t st);
cudaError_t peer_wait_launch(const unsigned long long *flags, const int *slots, int n, unsigned long long value, cudaStream_t st);

}  // namespace akb
