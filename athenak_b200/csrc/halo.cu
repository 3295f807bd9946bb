// halo.cu -- same-level ghost-zone exchange for cell-centred (CC) and face-centred (FC) data.
//
// Reference: src/bvals/bvals_cc.cpp:42-447, bvals_fc.cpp:63-436 and the index tables in
// buffs_cc.cpp:28-196, buffs_fc.cpp:29-99,388-456.  The reference packs every neighbour slab
// into a per-slot buffer and unpacks it in a second kernel, even for neighbours on the same
// rank.  Here the plan is resolved once on the host into flat copy descriptors:
//   * same-rank neighbours: ONE gather kernel writes ghost zones straight from the
//     neighbour's active zones (src index = dst index - offset*nx), no buffer round trip;
//   * off-rank neighbours: pack into / unpack from one contiguous message per peer rank
//     (entry layout as the reference's rank-packed buffers, bvals_cc.cpp:127,
//     bvals_fc.cpp:150-156), moved by NCCL send/recv or peer stores.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/athenak_b200.h"
#include "halo.h"

namespace akb {

// src/mesh/nghbr_index.hpp:28-54
int neighbor_index(int ix, int iy, int iz, int n1, int n2) {
  if ((std::abs(ix) + std::abs(iy) + std::abs(iz)) == 0) return -1;
  if (std::abs(ix * iy * iz) > 1) return -1;
  if (iz == 0) {
    if (ix * iy == 0) return std::abs(ix) * 2 * (ix + 1) + std::abs(iy) * 2 * (iy + 5) + (n1 + 2 * n2);
    return 16 + (ix + 1) + 2 * (iy + 1) + n1;
  }
  if (ix * iy == 0) return 24 + std::abs(ix) * (ix + 9) + std::abs(iy) * (iy + 17) + 2 * (iz + 1) + (n1 + 2 * n2);
  return 48 + (ix + 1) / 2 + (iy + 1) + 2 * (iz + 1);
}

struct Box { int lo[3], hi[3]; long count() const { return (long)(hi[0]-lo[0]+1)*(hi[1]-lo[1]+1)*(hi[2]-lo[2]+1); } };

// comp = -1: cell-centred; 0,1,2: x1f,x2f,x3f.  d: 0=i,1=j,2=k.  Same-level boxes.
static Box halo_box(const ak_mhd_cfg &c, int comp, const int o[3], bool send) {
  const int ng = c.ng, ng1 = ng - 1;
  // inactive dimensions (nx = 1) have no ghost cells: s = e = 0 (mesh.cpp:285-333)
  const int s[3] = {ng, c.nx2 > 1 ? ng : 0, c.nx3 > 1 ? ng : 0};
  const int e[3] = {s[0] + c.nx1 - 1, s[1] + c.nx2 - 1, s[2] + c.nx3 - 1};
  Box b;
  for (int d = 0; d < 3; ++d) {
    const int f = (comp == d) ? 1 : 0;  // this component is staggered along d
    if (send) {                          // buffs_cc.cpp:36-46, buffs_fc.cpp:44-83
      if (o[d] == 0) { b.lo[d] = s[d]; b.hi[d] = e[d] + f; }
      else if (o[d] > 0) { b.lo[d] = e[d] - ng1; b.hi[d] = e[d]; }
      else { b.lo[d] = s[d] + f; b.hi[d] = s[d] + ng1 + f; }
    } else {                             // buffs_cc.cpp:170-196, buffs_fc.cpp:402-441
      if (o[d] == 0) { b.lo[d] = s[d]; b.hi[d] = e[d] + f; }
      else if (o[d] > 0) { b.lo[d] = e[d] + 1 + f; b.hi[d] = e[d] + ng + f; }
      else { b.lo[d] = s[d] - ng; b.hi[d] = s[d] - 1; }
    }
  }
  return b;
}

__global__ void __launch_bounds__(256)
halo_copy_kernel(const HaloDesc *__restrict__ descs, const int2 *__restrict__ tiles, HaloArrays A, int nvar, long varstride,
                 double *__restrict__ buf, int mode) {
  const int2 tl = tiles[blockIdx.x];        // {descriptor, chunk}: one CTA per HALO_TILE elements of one descriptor
  const HaloDesc d = descs[tl.x];
  const int ni = d.hi[0] - d.lo[0] + 1, nj = d.hi[1] - d.lo[1] + 1, nk = d.hi[2] - d.lo[2] + 1;
  const long per = (long)ni * nj * nk;
  const long total = per * nvar;
  const int c = d.comp < 0 ? 0 : d.comp;
  double *arr = A.p[c];
  const long e1 = A.ext1[c], e2 = A.ext2[c], bs = A.bstride[c];
  const long qend = ((long)(tl.y + 1) * HALO_TILE < total) ? (long)(tl.y + 1) * HALO_TILE : total;
  // 32-bit index decode: a slab holds far fewer than 2^31 elements (checked by the launcher); the 64-bit divisions of the first
  // version were most of the kernel's instructions
  for (unsigned q = (unsigned)tl.y * HALO_TILE + threadIdx.x; q < (unsigned)qend; q += blockDim.x) {
    unsigned t = q;
    const int i = (int)(t % (unsigned)ni); t /= (unsigned)ni;
    const int j = (int)(t % (unsigned)nj); t /= (unsigned)nj;
    const int k = (int)(t % (unsigned)nk);
    const int v = (int)(t / (unsigned)nk);
    const long a_idx = ((long)(d.lo[2] + k) * e2 + (d.lo[1] + j)) * e1 + (d.lo[0] + i) + v * varstride;
    const long base = (d.comp < 0) ? d.boff * nvar + v * per : d.boff;
    if (mode == HALO_LOCAL) {
      const long s_idx = ((long)(d.lo[2] + k + d.shift[2]) * e2 + (d.lo[1] + j + d.shift[1])) * e1 +
                         (d.lo[0] + i + d.shift[0]) + v * varstride;
      arr[(long)d.m * bs + a_idx] = arr[(long)d.src_m * bs + s_idx];
    } else if (mode == HALO_PACK) {
      buf[base + i + (long)ni * (j + (long)nj * k)] = arr[(long)d.m * bs + a_idx];
    } else {
      arr[(long)d.m * bs + a_idx] = buf[base + i + (long)ni * (j + (long)nj * k)];
    }
  }
}

// EMF (edge-field) boundary boxes for flux correction, same level: buffs_fc.cpp:327-368 (send)
// and :892-935 (recv) -- identical formulae.  v: 0 x1e, 1 x2e, 2 x3e.
static Box emf_box(const ak_mhd_cfg &c, int v, const int o[3]) {
  const int s[3] = {c.ng, c.nx2 > 1 ? c.ng : 0, c.nx3 > 1 ? c.ng : 0};
  const int e[3] = {s[0] + c.nx1 - 1, s[1] + c.nx2 - 1, s[2] + c.nx3 - 1};
  Box b;
  for (int d = 0; d < 3; ++d) {
    if (o[d] == 0) { b.lo[d] = s[d]; b.hi[d] = e[d] + (v == d ? 0 : 1); }
    else if (o[d] > 0) { b.lo[d] = e[d] + 1; b.hi[d] = e[d] + 1; }
    else { b.lo[d] = s[d]; b.hi[d] = s[d]; }
  }
  return b;
}
// faces carry the two tangential components, block edges the component along the edge
// (flux_correct_fc.cpp:517-635); corners carry nothing
static bool emf_slot_has(const int o[3], int v) {
  const int nz = (o[0] != 0) + (o[1] != 0) + (o[2] != 0);
  return (nz == 1 || nz == 2) && o[v] == 0;
}

__global__ void __launch_bounds__(128)
emf_sumavg_kernel(const EmfSumDesc *__restrict__ descs, const int2 *__restrict__ tiles, HaloArrays A, const double *__restrict__ rbuf) {
  const int2 tl = tiles[blockIdx.x];
  const EmfSumDesc d = descs[tl.x];
  const int ni = d.hi[0] - d.lo[0] + 1, nj = d.hi[1] - d.lo[1] + 1, nk = d.hi[2] - d.lo[2] + 1;
  const long total = (long)ni * nj * nk;
  double *arr = A.p[d.v];
  const long e1 = A.ext1[d.v], e2 = A.ext2[d.v], bs = A.bstride[d.v];
  const long qend = ((long)(tl.y + 1) * HALO_TILE < total) ? (long)(tl.y + 1) * HALO_TILE : total;
  for (long q = (long)tl.y * HALO_TILE + threadIdx.x; q < qend; q += blockDim.x) {
    long t = q;
    const int i = d.lo[0] + (int)(t % ni); t /= ni;
    const int j = d.lo[1] + (int)(t % nj);
    const int k = d.lo[2] + (int)(t / nj);
    const long idx = (long)d.m * bs + ((long)k * e2 + j) * e1 + i;
    double val = arr[idx];
    for (int s = 0; s < d.nsrc; ++s)   // SumBoundaryFluxes: increasing slot order
      val += rbuf[d.soff[s] + (i - d.slo[s][0]) + (long)d.sni[s] * ((j - d.slo[s][1]) + (long)d.snj[s] * (k - d.slo[s][2]))];
    if (d.face) val *= 0.5;                       // AverageBoundaryFluxes, face interiors
    else val /= static_cast<double>(d.count);     // block edges: / nflx
    arr[idx] = val;
  }
}

static const int kOff[26][3] = {
    {-1,0,0},{1,0,0},{0,-1,0},{0,1,0},{-1,-1,0},{1,-1,0},{-1,1,0},{1,1,0},{0,0,-1},{0,0,1},
    {-1,0,-1},{1,0,-1},{-1,0,1},{1,0,1},{0,-1,-1},{0,1,-1},{0,-1,1},{0,1,1},
    {-1,-1,-1},{1,-1,-1},{-1,1,-1},{1,1,-1},{-1,-1,1},{1,-1,1},{-1,1,1},{1,1,1}};

HaloPlan *halo_build(const ak_mhd_cfg &cfg, const int *nghbr, int my_rank, int gids_start, std::string &err) {
  // hydro twin: cell-centred data only; MHD adds the face fields and the EMF correction.
  // 1-D/2-D/3-D: only the active dimensions have neighbours, ghost cells and end lines.
  const bool cc_only = cfg.rsolver >= AK_RSOLVER_HYD_ADVECT;
  const bool act[3] = {true, cfg.nx2 > 1, cfg.nx3 > 1};
  if (cfg.nx1 < cfg.ng || (act[1] && cfg.nx2 < cfg.ng) || (act[2] && cfg.nx3 < cfg.ng)) { err = "halo: block smaller than nghost"; return nullptr; }
  auto *P = new HaloPlan();
  P->cfg = cfg;
  const int nx[3] = {cfg.nx1, cfg.nx2, cfg.nx3};
  struct Entry { int gid_src, n_src, m, slot; int o[3]; };
  std::map<int, std::vector<Entry>> send, recv;   // per peer rank
  for (int m = 0; m < cfg.nmb; ++m) {
    for (int q = 0; q < 26; ++q) {
      const int *o = kOff[q];
      if ((o[1] && !act[1]) || (o[2] && !act[2])) continue;   // meshblock.cpp:170-425: neighbours only in active dims
      const int slot = neighbor_index(o[0], o[1], o[2], 0, 0);
      const int *nb = nghbr + ((long)m * 56 + slot) * 4;
      const int gid = nb[0], rank = nb[2];
      if (gid < 0) continue;
      const int no[3] = {-o[0], -o[1], -o[2]};
      if (rank == my_rank) {
        // ghost box of m on side o is filled from block gid's active zones
        const int src_m = gid - gids_start;
        if (src_m < 0 || src_m >= cfg.nmb) { err = "halo: same-rank neighbour gid outside this pack"; delete P; return nullptr; }
        for (int comp = -1; comp < (cc_only ? 0 : 3); ++comp) {
          Box b = halo_box(cfg, comp, o, false);
          HaloDesc d{};
          d.m = m; d.src_m = src_m; d.comp = comp; d.boff = 0;
          for (int a = 0; a < 3; ++a) { d.lo[a] = b.lo[a]; d.hi[a] = b.hi[a]; d.shift[a] = -o[a] * nx[a]; }
          (comp < 0 ? P->local_cc : P->local_fc).push_back(d);
        }
      } else {
        // I receive into (m, slot) what the peer sends from (gid, slot(-o)); and I send my
        // box on side o to the peer's block gid, which receives it in ITS slot(-o).
        Entry r{gid, neighbor_index(no[0], no[1], no[2], 0, 0), m, slot, {o[0], o[1], o[2]}};
        recv[rank].push_back(r);
        Entry s{gids_start + m, slot, m, slot, {o[0], o[1], o[2]}};
        send[rank].push_back(s);
      }
    }
  }
  auto key_lt = [](const Entry &a, const Entry &b) { return a.gid_src != b.gid_src ? a.gid_src < b.gid_src : a.n_src < b.n_src; };
  for (auto &kv : send) {
    const int rank = kv.first;
    HaloPeer peer;
    peer.rank = rank;
    auto &sv = kv.second;
    auto &rv = recv[rank];
    std::sort(sv.begin(), sv.end(), key_lt);
    std::sort(rv.begin(), rv.end(), key_lt);
    long cc_off = 0, fc_off = 0;
    for (const Entry &e : sv) {
      Box b = halo_box(cfg, -1, e.o, true);
      HaloDesc d{}; d.m = e.m; d.src_m = e.m; d.comp = -1; d.boff = cc_off;
      for (int a = 0; a < 3; ++a) { d.lo[a] = b.lo[a]; d.hi[a] = b.hi[a]; }
      peer.pack_cc.push_back(d);
      {  // peer-memory path: my send box on side o lands in the neighbour's ghost box on ITS side -o
        const int no[3] = {-e.o[0], -e.o[1], -e.o[2]};
        const int ngid = nghbr[((long)e.m * 56 + e.slot) * 4];
        for (int comp = -1; comp < (cc_only ? 0 : 3); ++comp) {
          Box sb = halo_box(cfg, comp, e.o, true), rb = halo_box(cfg, comp, no, false);
          PushDesc pd{};
          pd.m = e.m; pd.comp = comp; pd.peer = (int)P->peers.size(); pd.dst_gid = ngid; pd.dst_m = -1;
          for (int a = 0; a < 3; ++a) { pd.lo[a] = sb.lo[a]; pd.hi[a] = sb.hi[a]; pd.dlo[a] = rb.lo[a]; }
          if (sb.count() != rb.count()) { err = "halo: push source/destination boxes differ"; delete P; return nullptr; }
          P->push.push_back(pd);
        }
      }
      cc_off += b.count();   // per variable; scaled by nvar at launch
      if (cc_only) continue;
      long ndat = 0;
      for (int comp = 0; comp < 3; ++comp) ndat = std::max(ndat, halo_box(cfg, comp, e.o, true).count());
      for (int comp = 0; comp < 3; ++comp) {
        Box bb = halo_box(cfg, comp, e.o, true);
        HaloDesc f{}; f.m = e.m; f.src_m = e.m; f.comp = comp; f.boff = fc_off + ndat * comp;
        for (int a = 0; a < 3; ++a) { f.lo[a] = bb.lo[a]; f.hi[a] = bb.hi[a]; }
        peer.pack_fc.push_back(f);
      }
      fc_off += 3 * ndat;
    }
    peer.cc_count = cc_off; peer.fc_count = fc_off;
    cc_off = 0; fc_off = 0;
    for (const Entry &e : rv) {
      // sender's box is its send box on side -o; same extents as my recv box on side o
      const int so[3] = {-e.o[0], -e.o[1], -e.o[2]};
      Box b = halo_box(cfg, -1, e.o, false);
      HaloDesc d{}; d.m = e.m; d.src_m = e.m; d.comp = -1; d.boff = cc_off;
      for (int a = 0; a < 3; ++a) { d.lo[a] = b.lo[a]; d.hi[a] = b.hi[a]; }
      peer.unpack_cc.push_back(d);
      cc_off += b.count();
      if (cc_only) continue;
      long ndat = 0;
      for (int comp = 0; comp < 3; ++comp) ndat = std::max(ndat, halo_box(cfg, comp, so, true).count());
      for (int comp = 0; comp < 3; ++comp) {
        Box bb = halo_box(cfg, comp, e.o, false);
        HaloDesc f{}; f.m = e.m; f.src_m = e.m; f.comp = comp; f.boff = fc_off + ndat * comp;
        for (int a = 0; a < 3; ++a) { f.lo[a] = bb.lo[a]; f.hi[a] = bb.hi[a]; }
        peer.unpack_fc.push_back(f);
      }
      fc_off += 3 * ndat;
    }
    if (cc_off != peer.cc_count || fc_off != peer.fc_count) { err = "halo: send/recv message sizes differ"; delete P; return nullptr; }
    P->peers.push_back(peer);
  }
  {  // push descriptors: CC entries first, then FC (one launch can take either range)
    auto mid = std::stable_partition(P->push.begin(), P->push.end(), [](const PushDesc &d) { return d.comp < 0; });
    P->n_push_cc = (long)(mid - P->push.begin());
  }
  // ---- EMF flux correction plan -----------------------------------------------------------
  if (!cc_only) {
    // receive segments: (m, slot, v) -> offset in emf_rbuf.  Same-rank sources first, then one
    // contiguous range per peer ordered by (source gid, source slot, v) = the sender's order.
    std::map<long, long> roff;   // key = (m*48 + slot)*3 + v
    auto key = [](int m, int slot, int v) { return ((long)m * 48 + slot) * 3 + v; };
    long off = 0;
    struct RSeg { int gid_src, n_src, v, m, slot; long cnt; };
    std::map<int, std::vector<RSeg>> rsegs;   // per peer
    for (int m = 0; m < cfg.nmb; ++m)
      for (int q = 0; q < 26; ++q) {
        const int *o = kOff[q];
        const int nz = (o[0] != 0) + (o[1] != 0) + (o[2] != 0);
        if (nz == 3) continue;
        if ((o[1] && !act[1]) || (o[2] && !act[2])) continue;   // flux_correct_fc.cpp:898,953,963,1013,1023
        const int slot = neighbor_index(o[0], o[1], o[2], 0, 0);
        const int *nb = nghbr + ((long)m * 56 + slot) * 4;
        if (nb[0] < 0) continue;
        for (int v = 0; v < 3; ++v) {
          if (!emf_slot_has(o, v)) continue;
          const long cnt = emf_box(cfg, v, o).count();
          if (nb[2] == my_rank) { roff[key(m, slot, v)] = off; off += cnt; }
          else rsegs[nb[2]].push_back(RSeg{nb[0], neighbor_index(-o[0], -o[1], -o[2], 0, 0), v, m, slot, cnt});
        }
      }
    for (auto &peer : P->peers) {
      auto &v = rsegs[peer.rank];
      std::sort(v.begin(), v.end(), [](const RSeg &a, const RSeg &b) {
        if (a.gid_src != b.gid_src) return a.gid_src < b.gid_src;
        if (a.n_src != b.n_src) return a.n_src < b.n_src;
        return a.v < b.v; });
      peer.emf_recv_base = off;
      for (auto &r : v) { roff[key(r.m, r.slot, r.v)] = off; off += r.cnt; }
      peer.emf_count = off - peer.emf_recv_base;
    }
    P->emf_rbuf_count = off;
    // pack descriptors: my box on side o goes to the neighbour's segment (its slot(-o))
    std::map<int, std::vector<std::pair<RSeg, HaloDesc>>> psend;
    for (int m = 0; m < cfg.nmb; ++m)
      for (int q = 0; q < 26; ++q) {
        const int *o = kOff[q];
        const int nz = (o[0] != 0) + (o[1] != 0) + (o[2] != 0);
        if (nz == 3) continue;
        if ((o[1] && !act[1]) || (o[2] && !act[2])) continue;   // flux_correct_fc.cpp:898,953,963,1013,1023
        const int slot = neighbor_index(o[0], o[1], o[2], 0, 0);
        const int *nb = nghbr + ((long)m * 56 + slot) * 4;
        if (nb[0] < 0) continue;
        const int dslot = neighbor_index(-o[0], -o[1], -o[2], 0, 0);
        for (int v = 0; v < 3; ++v) {
          if (!emf_slot_has(o, v)) continue;
          Box b = emf_box(cfg, v, o);
          HaloDesc d{};
          d.m = m; d.src_m = m; d.comp = v;
          for (int a = 0; a < 3; ++a) { d.lo[a] = b.lo[a]; d.hi[a] = b.hi[a]; }
          if (nb[2] == my_rank) {
            d.boff = roff.at(key(nb[0] - gids_start, dslot, v));
            P->emf_pack_local.push_back(d);
          } else {
            psend[nb[2]].push_back({RSeg{gids_start + m, slot, v, m, slot, b.count()}, d});
          }
        }
      }
    for (auto &peer : P->peers) {
      auto &v = psend[peer.rank];
      std::sort(v.begin(), v.end(), [](const std::pair<RSeg, HaloDesc> &a, const std::pair<RSeg, HaloDesc> &b) {
        if (a.first.gid_src != b.first.gid_src) return a.first.gid_src < b.first.gid_src;
        if (a.first.n_src != b.first.n_src) return a.first.n_src < b.first.n_src;
        return a.first.v < b.first.v; });
      long so = 0;
      for (auto &pr : v) { pr.second.boff = so; so += pr.first.cnt; peer.pack_emf.push_back(pr.second); }
      if (so != peer.emf_count) { err = "halo: EMF send/recv message sizes differ"; delete P; return nullptr; }
    }
    // sum/scale regions
    for (int m = 0; m < cfg.nmb; ++m)
      for (int q = 0; q < 26; ++q) {
        const int *o = kOff[q];
        const int nz = (o[0] != 0) + (o[1] != 0) + (o[2] != 0);
        if (nz == 3) continue;
        if ((o[1] && !act[1]) || (o[2] && !act[2])) continue;   // flux_correct_fc.cpp:898,953,963,1013,1023
        const int slot = neighbor_index(o[0], o[1], o[2], 0, 0);
        for (int v = 0; v < 3; ++v) {
          if (!emf_slot_has(o, v)) continue;
          EmfSumDesc d{};
          d.m = m; d.v = v;
          Box b = emf_box(cfg, v, o);
          std::vector<std::pair<int, const int *>> src;   // (slot, offsets) of contributing neighbours
          int fo[2][3];
          if (nz == 1) {
            if (nghbr[((long)m * 56 + slot) * 4] < 0) continue;   // physical boundary: nothing to average
            d.face = 1;
            for (int a = 0; a < 3; ++a) if (o[a] == 0 && a != v && act[a]) { b.lo[a] += 1; b.hi[a] -= 1; }   // drop the end lines (active dims only, :846,873)
            if (b.count() <= 0) continue;
            src.push_back({slot, o});
          } else {
            d.face = 0;
            int nf = 0;
            for (int a = 0; a < 3; ++a) if (o[a] != 0) { fo[nf][0] = fo[nf][1] = fo[nf][2] = 0; fo[nf][a] = o[a]; ++nf; }
            for (int f = 0; f < 2; ++f) {
              const int fs = neighbor_index(fo[f][0], fo[f][1], fo[f][2], 0, 0);
              if (nghbr[((long)m * 56 + fs) * 4] >= 0) src.push_back({fs, fo[f]});
            }
            if (nghbr[((long)m * 56 + slot) * 4] >= 0) src.push_back({slot, o});
            if (src.empty()) continue;
            std::sort(src.begin(), src.end(), [](const std::pair<int, const int *> &a, const std::pair<int, const int *> &b) { return a.first < b.first; });
          }
          for (int a = 0; a < 3; ++a) { d.lo[a] = b.lo[a]; d.hi[a] = b.hi[a]; }
          d.nsrc = (int)src.size();
          d.count = 1 + d.nsrc;
          for (int s = 0; s < d.nsrc; ++s) {
            Box sb = emf_box(cfg, v, src[s].second);
            d.soff[s] = roff.at(key(m, src[s].first, v));
            for (int a = 0; a < 3; ++a) d.slo[s][a] = sb.lo[a];
            d.sni[s] = sb.hi[0] - sb.lo[0] + 1;
            d.snj[s] = sb.hi[1] - sb.lo[1] + 1;
          }
          P->emf_sum.push_back(d);
        }
      }
  }
  return P;
}

cudaError_t halo_ensure_uploaded(HaloPlan *P) {
  if (P->uploaded) return cudaSuccess;
  auto up = [&](const std::vector<HaloDesc> &v, HaloDesc **dptr) -> cudaError_t {
    *dptr = nullptr;
    if (v.empty()) return cudaSuccess;
    cudaError_t e = cudaMalloc(dptr, v.size() * sizeof(HaloDesc));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dptr, v.data(), v.size() * sizeof(HaloDesc), cudaMemcpyHostToDevice);
  };
  cudaError_t e;
  if ((e = up(P->local_cc, &P->d_local_cc)) != cudaSuccess) return e;
  if ((e = up(P->local_fc, &P->d_local_fc)) != cudaSuccess) return e;
  for (auto &p : P->peers) {
    if ((e = up(p.pack_cc, &p.d_pack_cc)) != cudaSuccess) return e;
    if ((e = up(p.unpack_cc, &p.d_unpack_cc)) != cudaSuccess) return e;
    if ((e = up(p.pack_fc, &p.d_pack_fc)) != cudaSuccess) return e;
    if ((e = up(p.unpack_fc, &p.d_unpack_fc)) != cudaSuccess) return e;
  }
  if ((e = up(P->emf_pack_local, &P->d_emf_pack_local)) != cudaSuccess) return e;
  for (auto &p : P->peers)
    if ((e = up(p.pack_emf, &p.d_pack_emf)) != cudaSuccess) return e;
  if (!P->emf_sum.empty()) {
    if ((e = cudaMalloc(&P->d_emf_sum, P->emf_sum.size() * sizeof(EmfSumDesc))) != cudaSuccess) return e;
    if ((e = cudaMemcpy(P->d_emf_sum, P->emf_sum.data(), P->emf_sum.size() * sizeof(EmfSumDesc), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
  }
  if (P->emf_rbuf_count > 0 && (e = cudaMalloc(&P->d_emf_rbuf, P->emf_rbuf_count * sizeof(double))) != cudaSuccess) return e;
  P->uploaded = true;
  return cudaSuccess;
}

// tile table of a descriptor list: built and uploaded on first use, cached per (list, nvar)
template <class Desc, class CountFn>
static cudaError_t tile_table(HaloPlan *P, const std::vector<Desc> &host, const void *key, int nvar, CountFn count, HaloPlan::TileTab &out) {
  auto it = P->tiles.find({key, nvar});
  if (it != P->tiles.end()) { out = it->second; return cudaSuccess; }
  std::vector<int2> t;
  for (size_t q = 0; q < host.size(); ++q) {
    const long total = count(host[q]);
    if (total >= 0x7fffffffL) return cudaErrorInvalidConfiguration;      // the kernels decode element indices in 32 bits
    for (long c = 0; c * HALO_TILE < total; ++c) t.push_back(make_int2((int)q, (int)c));
  }
  HaloPlan::TileTab tab;
  tab.n = (int)t.size();
  if (tab.n) {
    cudaError_t e = cudaMalloc(&tab.d, t.size() * sizeof(int2));
    if (e == cudaSuccess) e = cudaMemcpy(tab.d, t.data(), t.size() * sizeof(int2), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
  }
  P->tiles[{key, nvar}] = tab;
  out = tab;
  return cudaSuccess;
}

cudaError_t emf_sumavg_launch(HaloPlan *P, const HaloArrays &A, cudaStream_t st) {
  if (P->emf_sum.empty()) return cudaSuccess;
  HaloPlan::TileTab tab;
  cudaError_t e = tile_table(P, P->emf_sum, P->d_emf_sum, 1, [](const EmfSumDesc &d) {
    return (long)(d.hi[0] - d.lo[0] + 1) * (d.hi[1] - d.lo[1] + 1) * (d.hi[2] - d.lo[2] + 1); }, tab);
  if (e != cudaSuccess) return e;
  if (tab.n) emf_sumavg_kernel<<<tab.n, 128, 0, st>>>(P->d_emf_sum, tab.d, A, P->d_emf_rbuf);
  return cudaGetLastError();
}

void halo_free(HaloPlan *P) {
  if (!P) return;
  cudaFree(P->d_local_cc); cudaFree(P->d_local_fc); cudaFree(P->d_emf_pack_local); cudaFree(P->d_emf_sum); cudaFree(P->d_emf_rbuf);
  for (auto &p : P->peers) { cudaFree(p.d_pack_cc); cudaFree(p.d_unpack_cc); cudaFree(p.d_pack_fc); cudaFree(p.d_unpack_fc); cudaFree(p.d_pack_emf); }
  cudaFree(P->d_push); for (auto *q : P->d_sig) cudaFree(q); cudaFree(P->d_wait_slots);
  for (auto &kv : P->tiles) cudaFree(kv.second.d);
  delete P;
}

cudaError_t halo_launch(HaloPlan *P, const std::vector<HaloDesc> &host, const HaloDesc *dev, const HaloArrays &A, int nvar,
                        long varstride, double *buf, int mode, cudaStream_t st) {
  if (host.empty()) return cudaSuccess;
  HaloPlan::TileTab tab;
  cudaError_t e = tile_table(P, host, dev, nvar, [nvar](const HaloDesc &d) {
    return (long)(d.hi[0] - d.lo[0] + 1) * (d.hi[1] - d.lo[1] + 1) * (d.hi[2] - d.lo[2] + 1) * nvar; }, tab);
  if (e != cudaSuccess) return e;
  if (tab.n) halo_copy_kernel<<<tab.n, 256, 0, st>>>(dev, tab.d, A, nvar, varstride, buf, mode);
  return cudaGetLastError();
}

// ---- peer-memory exchange ------------------------------------------------------------------
// One launch moves every off-rank slab of this rank (all peers, CC and/or FC): coalesced reads of
// my arrays, stores straight into the peer GPUs' arrays over NVLink (mapped with cudaIpc by the
// host layer).  grid.y = descriptor.
__global__ void __launch_bounds__(256)
halo_push_kernel(const PushDesc *__restrict__ descs, const int2 *__restrict__ tiles, HaloArrays CCa, int nvar, long varstride, HaloArrays FCa,
                 PushPeerPtrs ptrs) {
  const int2 tl = tiles[blockIdx.x];
  const PushDesc d = descs[tl.x];
  const int ni = d.hi[0] - d.lo[0] + 1, nj = d.hi[1] - d.lo[1] + 1, nk = d.hi[2] - d.lo[2] + 1;
  const long per = (long)ni * nj * nk;
  const bool cc = d.comp < 0;
  const int c = cc ? 0 : d.comp;
  const HaloArrays &A = cc ? CCa : FCa;
  const double *src = A.p[c] + (long)d.m * A.bstride[c];
  double *dst = ptrs.p[d.peer][cc ? 0 : 1 + d.comp] + (long)d.dst_m * A.bstride[c];
  const long e1 = A.ext1[c], e2 = A.ext2[c];
  const long total = per * (cc ? nvar : 1);
  const long qend = ((long)(tl.y + 1) * HALO_TILE < total) ? (long)(tl.y + 1) * HALO_TILE : total;
  for (unsigned q = (unsigned)tl.y * HALO_TILE + threadIdx.x; q < (unsigned)qend; q += blockDim.x) {     // 32-bit decode, see halo_copy_kernel
    unsigned t = q;
    const int i = (int)(t % (unsigned)ni); t /= (unsigned)ni;
    const int j = (int)(t % (unsigned)nj); t /= (unsigned)nj;
    const int k = (int)(t % (unsigned)nk);
    const long v = (long)(t / (unsigned)nk) * (cc ? varstride : 0);
    dst[((long)(d.dlo[2] + k) * e2 + (d.dlo[1] + j)) * e1 + (d.dlo[0] + i) + v] =
        src[((long)(d.lo[2] + k) * e2 + (d.lo[1] + j)) * e1 + (d.lo[0] + i) + v];
  }
}

cudaError_t halo_push_launch(HaloPlan *P, const HaloArrays &cc, int nvar, long varstride, const HaloArrays &fc, bool do_cc,
                             bool do_fc, cudaStream_t st) {
  if (P->push.empty()) return cudaSuccess;
  const long first = do_cc ? 0 : P->n_push_cc;
  const long last = do_fc ? (long)P->push.size() : P->n_push_cc;
  if (last <= first) return cudaSuccess;
  // one tile table per (range, nvar): the sub-list is addressed through its own base pointer
  const std::vector<PushDesc> sub(P->push.begin() + first, P->push.begin() + last);
  HaloPlan::TileTab tab;
  const void *key = reinterpret_cast<const char *>(P->d_push + first) + (do_cc ? 1 : 0) + (do_fc ? 2 : 0);   // distinct key per range
  cudaError_t e = tile_table(P, sub, key, nvar, [nvar](const PushDesc &d) {
    return (long)(d.hi[0] - d.lo[0] + 1) * (d.hi[1] - d.lo[1] + 1) * (d.hi[2] - d.lo[2] + 1) * (d.comp < 0 ? nvar : 1); }, tab);
  if (e != cudaSuccess) return e;
  if (tab.n) halo_push_kernel<<<tab.n, 256, 0, st>>>(P->d_push + first, tab.d, cc, nvar, varstride, fc, P->push_ptrs);
  return cudaGetLastError();
}

// Cross-GPU flags: monotone counters in each rank's memory, one slot per (kind, writer rank).
// signal: runs after the producing kernel in stream order (its stores are complete and visible
// system-wide at the kernel boundary), then publishes the counter in every peer's memory.
__global__ void peer_signal_kernel(unsigned long long **slots, int n, unsigned long long value) {
  const int p = threadIdx.x;
  if (p < n) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(slots[p]) = value;
    __threadfence_system();
  }
}
// wait: spins (one thread per awaited peer) until every peer's counter has reached `value`
__global__ void peer_wait_kernel(const unsigned long long *flags, const int *slots, int n, unsigned long long value) {
  const int p = threadIdx.x;
  if (p < n) {
    const volatile unsigned long long *f = flags + slots[p];
    while (*f < value) __nanosleep(200);
    __threadfence_system();
  }
}
cudaError_t peer_signal_launch(unsigned long long **slots, int n, unsigned long long value, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  peer_signal_kernel<<<1, 32 * ((n + 31) / 32), 0, st>>>(slots, n, value);
  return cudaGetLastError();
}
cudaError_t peer_wait_launch(const unsigned long long *flags, const int *slots, int n, unsigned long long value, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  peer_wait_kernel<<<1, 32 * ((n + 31) / 32), 0, st>>>(flags, slots, n, value);
  return cudaGetLastError();
}

HaloArrays halo_arrays_cc(const ak_mhd_cfg &c, double *a, int nvar) {
  const long n1 = c.nx1 + 2 * c.ng, n2 = c.nx2 > 1 ? c.nx2 + 2 * c.ng : 1, n3 = c.nx3 > 1 ? c.nx3 + 2 * c.ng : 1;
  HaloArrays A{};
  A.p[0] = a; A.ext1[0] = n1; A.ext2[0] = n2; A.bstride[0] = (long)nvar * n1 * n2 * n3;
  return A;
}

HaloArrays halo_arrays_fc(const ak_mhd_cfg &c, double *x1f, double *x2f, double *x3f) {
  const long n1 = c.nx1 + 2 * c.ng, n2 = c.nx2 > 1 ? c.nx2 + 2 * c.ng : 1, n3 = c.nx3 > 1 ? c.nx3 + 2 * c.ng : 1;
  HaloArrays A{};
  A.p[0] = x1f; A.ext1[0] = n1 + 1; A.ext2[0] = n2;     A.bstride[0] = (n1 + 1) * n2 * n3;
  A.p[1] = x2f; A.ext1[1] = n1;     A.ext2[1] = n2 + 1; A.bstride[1] = n1 * (n2 + 1) * n3;
  A.p[2] = x3f; A.ext1[2] = n1;     A.ext2[2] = n2;     A.bstride[2] = n1 * n2 * (n3 + 1);
  return A;
}

HaloArrays halo_arrays_ec(const ak_mhd_cfg &c, double *x1e, double *x2e, double *x3e) {
  const long n1 = c.nx1 + 2 * c.ng, n2 = c.nx2 > 1 ? c.nx2 + 2 * c.ng : 1, n3 = c.nx3 > 1 ? c.nx3 + 2 * c.ng : 1;
  HaloArrays A{};
  A.p[0] = x1e; A.ext1[0] = n1;     A.ext2[0] = n2 + 1; A.bstride[0] = n1 * (n2 + 1) * (n3 + 1);
  A.p[1] = x2e; A.ext1[1] = n1 + 1; A.ext2[1] = n2;     A.bstride[1] = (n1 + 1) * n2 * (n3 + 1);
  A.p[2] = x3e; A.ext1[2] = n1 + 1; A.ext2[2] = n2 + 1; A.bstride[2] = (n1 + 1) * (n2 + 1) * n3;
  return A;
}

}  // namespace akb
