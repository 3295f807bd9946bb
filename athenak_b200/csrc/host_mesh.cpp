// host_mesh.cpp -- host-side (CPU, plain C++17) mirror of the reference's uniform-mesh metadata:
// MeshBlockTree Z-ordering, LoadBalance, MeshBlock sizes/BC flags and the 56-slot neighbour
// table, plus the blast problem generator that provides the synthetic inputs of the benchmark.
//
// Reference: src/mesh/build_tree.cpp:32-307 (BuildTreeFromScratch), meshblock_tree.cpp:60-90
// (CreateRootGrid), :336-350 (Z-order gids), :355-430 (FindNeighbor), load_balance.cpp:38-88,
// meshblock.cpp:25-131 (sizes, BC flags), :142-430 (SetNeighbors, same-level branch),
// nghbr_index.hpp:28-54, coordinates/cell_locations.hpp:24-39, pgen/fluids/blast.cpp:134-402.
// The integer maps must be bit-exact with the reference (tests/test_host_mesh.py).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/athenak_b200.h"
#include "halo.h"

namespace {

struct LLoc { int32_t lx1, lx2, lx3, level; };

struct HostMesh {
  int nmesh[3], nblock[3], nroot[3];
  double xmin[3], xmax[3];
  int bcs[6];          // per mesh face: enum ak_bc_flag
  int nranks;
  int root_level, nmb_total;
  int max_level = 0;
  bool multilevel = false;     // <mesh_refinement>/refinement != none: neighbour sub-block slots are in use
  std::vector<LLoc> lloc;            // Z-ordered
  std::vector<int> rank_of, gids_start, nmb_rank;
  std::vector<double> mb_size;       // [nmb_total][9]
  std::vector<int> mb_bcs;           // [nmb_total][6]
  std::vector<int> nghbr;            // [nmb_total][56][4]
  std::string err;
  struct Tree *tree = nullptr;       // kept for adaptive refinement (ak_mesh_adapt)
  // MeshRefinement state (mesh_refinement.cpp:46-131)
  int ncyc_check_amr = 1, refinement_interval = 5;
  bool adaptive = false;
  std::vector<int> ncyc_since_ref;   // [nmb_total]
  std::vector<int> newtoold, oldtonew, old_flag;      // maps of the last ak_mesh_adapt that changed the mesh
};

thread_local std::string g_mesh_err;

// coordinates/cell_locations.hpp:24-39
inline double LeftEdgeX(int ith, int n, double xmin, double xmax) {
  double x = (static_cast<double>(ith)) / (static_cast<double>(n));
  return (x * xmax - x * xmin) - (0.5 * xmax - 0.5 * xmin) + (0.5 * xmin + 0.5 * xmax);
}
inline double CellCenterX(int ith, int n, double xmin, double xmax) {
  double x = (static_cast<double>(ith) + 0.5) / (static_cast<double>(n));
  return (x * xmax - x * xmin) - (0.5 * xmax - 0.5 * xmin) + (0.5 * xmin + 0.5 * xmax);
}

// MeshBlockTree (meshblock_tree.cpp:28-479): oct-/quad-/binary tree over logical locations.  The root
// grid can be incomplete; refining a node also refines its neighbours as needed so that adjacent
// leaves differ by at most one level (Refine -> AddNode on the 26 neighbours of the refined node).
struct Tree {
  struct Node {
    LLoc l{0, 0, 0, 0};
    int gid = -1;
    Node *leaf[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool split = false;          // pleaf_ != nullptr
    ~Node() { for (Node *c : leaf) delete c; }
  };
  const int *nroot; const int *bcs; int root_level; bool multi_d, three_d; int nleaf;
  Node *root = nullptr;
  Tree(const int *nroot_, const int *bcs_, int root_level_, bool multi_d_, bool three_d_)
      : nroot(nroot_), bcs(bcs_), root_level(root_level_), multi_d(multi_d_), three_d(three_d_) {
    nleaf = three_d ? 8 : (multi_d ? 4 : 2);          // meshblock_tree.cpp:64-69
    root = new Node();
    CreateRootGrid(root);
  }
  ~Tree() { delete root; }
  static Node *child(Node *p, int i, int j, int k) {
    Node *c = new Node();
    c->l = LLoc{(p->l.lx1 << 1) + i, (p->l.lx2 << 1) + j, (p->l.lx3 << 1) + k, p->l.level + 1};   // meshblock_tree.cpp:38-47
    c->gid = p->gid;            // a new leaf carries its parent's gid until the next Z-ordering (AMR: new -> old map)
    return c;
  }
  void CreateRootGrid(Node *t) {                       // meshblock_tree.cpp:64-91
    if (t->l.level == root_level) return;
    t->split = true;
    const int levfac = 1 << (root_level - t->l.level - 1);
    for (int n = 0; n < nleaf; ++n) {
      const int i = n & 1, j = (n >> 1) & 1, k = (n >> 2) & 1;
      if ((t->l.lx3 * 2 + k) * levfac < nroot[2] && (t->l.lx2 * 2 + j) * levfac < nroot[1] && (t->l.lx1 * 2 + i) * levfac < nroot[0]) {
        t->leaf[n] = child(t, i, j, k);
        CreateRootGrid(t->leaf[n]);
      }
    }
  }
  static int leaf_index(const LLoc &r, int level) {    // leaf of a node at `level` on the way to r
    const int sh = r.level - level - 1;
    return ((r.lx1 >> sh) & 1) + (((r.lx2 >> sh) & 1) << 1) + (((r.lx3 >> sh) & 1) << 2);
  }
  void AddNode(Node *t, const LLoc &r, int &nnew) {    // meshblock_tree.cpp:97-114
    if (t->l.level == r.level) return;
    if (!t->split) Refine(t, nnew);
    AddNode(t->leaf[leaf_index(r, t->l.level)], r, nnew);
  }
  bool periodic(int face) const { return bcs[face] == AK_BC_PERIODIC || (face < 2 && bcs[face] == AK_BC_SHEAR_PERIODIC); }
  void Refine(Node *t, int &nnew) {                    // meshblock_tree.cpp:151-249
    if (t->split) return;
    t->split = true;
    for (int n = 0; n < nleaf; ++n) t->leaf[n] = child(t, n & 1, (n >> 1) & 1, (n >> 2) & 1);
    const int sh = t->l.level - root_level;
    const int nmax[3] = {nroot[0] << sh, multi_d ? nroot[1] << sh : 1, three_d ? nroot[2] << sh : 1};
    const int oy0 = multi_d ? -1 : 0, oz0 = three_d ? -1 : 0;
    for (int oz = oz0; oz <= -oz0; ++oz) {
      int lz = t->l.lx3 + oz;
      if (lz < 0) { if (!periodic(4)) continue; lz = nmax[2] - 1; }
      if (lz >= nmax[2]) { if (!periodic(5)) continue; lz = 0; }
      for (int oy = oy0; oy <= -oy0; ++oy) {
        int ly = t->l.lx2 + oy;
        if (ly < 0) { if (!periodic(2)) continue; ly = nmax[1] - 1; }
        if (ly >= nmax[1]) { if (!periodic(3)) continue; ly = 0; }
        for (int ox = -1; ox <= 1; ++ox) {
          if (ox == 0 && oy == 0 && oz == 0) continue;
          int lx = t->l.lx1 + ox;
          if (lx < 0) { if (!periodic(0)) continue; lx = nmax[0] - 1; }
          if (lx >= nmax[0]) { if (!periodic(1)) continue; lx = 0; }
          AddNode(root, LLoc{lx, ly, lz, t->l.level}, nnew);
        }
      }
    }
    t->gid = -1;
    nnew += nleaf - 1;
  }
  // meshblock_tree.cpp:361-448: the block at (myloc + offset) if it is a leaf at the same level, the
  // coarser leaf that contains it, or -- when that position is refined -- the node whose leaves are finer
  Node *FindNeighbor(const LLoc &my, int ox1, int ox2, int ox3) const {
    int lx = my.lx1 + ox1, ly = my.lx2 + ox2, lz = my.lx3 + ox3;
    const int ll = my.level, sh = ll - root_level;
    if (lx < 0) { if (!periodic(0)) return nullptr; lx = (nroot[0] << sh) - 1; }
    if (lx >= (nroot[0] << sh)) { if (!periodic(1)) return nullptr; lx = 0; }
    if (ly < 0) { if (!periodic(2)) return nullptr; ly = (nroot[1] << sh) - 1; }
    if (ly >= (nroot[1] << sh)) { if (!periodic(3)) return nullptr; ly = 0; }
    if (lz < 0) { if (!periodic(4)) return nullptr; lz = (nroot[2] << sh) - 1; }
    if (lz >= (nroot[2] << sh)) { if (!periodic(5)) return nullptr; lz = 0; }
    if (ll < 1) return root;
    Node *bt = root;
    for (int level = 0; level < ll; ++level) {
      if (!bt->split) return bt;                       // coarser leaf (the reference insists level == ll-1)
      const int s = ll - level - 1;
      bt = bt->leaf[((lx >> s) & 1) + (((ly >> s) & 1) << 1) + (((lz >> s) & 1) << 2)];
      if (!bt) return nullptr;
    }
    return bt;     // leaf at the same level, or a split node: its leaves are one level finer
  }
  void ZOrder(Node *t, std::vector<LLoc> &out, std::vector<int> *oldgid = nullptr) {       // meshblock_tree.cpp:336-350
    if (!t->split) { if (oldgid) oldgid->push_back(t->gid); t->gid = (int)out.size(); out.push_back(t->l); return; }
    for (int n = 0; n < nleaf; ++n) if (t->leaf[n]) ZOrder(t->leaf[n], out, oldgid);
  }
  Node *FindMeshBlock(const LLoc &r) const {          // meshblock_tree.cpp:453-479
    Node *bt = root;
    for (int level = 0; level < r.level; ++level) {
      if (!bt || !bt->split) return nullptr;
      bt = bt->leaf[leaf_index(r, level)];
    }
    return bt;
  }
  // meshblock_tree.cpp:256-304: make this node a leaf again unless a neighbouring leaf that touches it is itself
  // refined (adjacent leaves may differ by one level only)
  void Derefine(Node *t, int &ndel) {
    const int s2 = multi_d ? -1 : 0, s3 = three_d ? -1 : 0;
    for (int ox3 = s3; ox3 <= -s3; ++ox3) for (int ox2 = s2; ox2 <= -s2; ++ox2) for (int ox1 = -1; ox1 <= 1; ++ox1) {
      Node *bt = FindNeighbor(t->l, ox1, ox2, ox3);
      if (!bt || !bt->split) continue;
      int lis, lie, ljs, lje, lks, lke;
      if (ox1 == -1) lis = lie = 1; else if (ox1 == 1) lis = lie = 0; else { lis = 0; lie = 1; }
      if (multi_d) { if (ox2 == -1) ljs = lje = 1; else if (ox2 == 1) ljs = lje = 0; else { ljs = 0; lje = 1; } } else ljs = lje = 0;
      if (three_d) { if (ox3 == -1) lks = lke = 1; else if (ox3 == 1) lks = lke = 0; else { lks = 0; lke = 1; } } else lks = lke = 0;
      for (int lk = lks; lk <= lke; ++lk) for (int lj = ljs; lj <= lje; ++lj) for (int li = lis; li <= lie; ++li)
        if (bt->leaf[li + (lj << 1) + (lk << 2)]->split) return;
    }
    t->gid = t->leaf[0]->gid;      // inherits the first leaf's gid
    for (Node *&c : t->leaf) { delete c; c = nullptr; }
    t->split = false;
    ndel += nleaf - 1;
  }
};

// load_balance.cpp:38-88 with unit costs
void load_balance(HostMesh &M) {
  const int nb = M.nmb_total, nr = M.nranks;
  std::vector<float> clist(nb, 1.0f);
  M.rank_of.assign(nb, 0); M.gids_start.assign(nr, 0); M.nmb_rank.assign(nr, 0);
  float totalcost = 0.0f;
  for (int i = 0; i < nb; i++) totalcost += clist[i];
  int j = nr - 1;
  float targetcost = totalcost / nr;
  float mycost = 0.0f;
  for (int i = nb - 1; i >= 0; i--) {
    mycost += clist[i];
    M.rank_of[i] = j;
    if (mycost >= targetcost && j > 0) {
      j--;
      totalcost -= mycost;
      mycost = 0.0f;
      targetcost = totalcost / (j + 1);
    }
  }
  M.gids_start[0] = 0;
  j = 0;
  for (int i = 1; i < nb; i++) {
    if (M.rank_of[i] != M.rank_of[i - 1]) {
      M.nmb_rank[j] = i - M.gids_start[j];
      M.gids_start[++j] = i;
    }
  }
  M.nmb_rank[j] = nb - M.gids_start[j];
}

}  // namespace

extern "C" {


const char *ak_mesh_last_error() { return g_mesh_err.c_str(); }

// Fills the per-block tables from the Z-ordered leaf list: sizes / BC flags (meshblock.cpp:25-131) and the
// 56-slot neighbour table with levels and sub-block indices (meshblock.cpp:142-430).
static void fill_tables(HostMesh *M, Tree &T) {
  const int nb = M->nmb_total;
  const bool multi_d = M->nmesh[1] > 1, three_d = M->nmesh[2] > 1;
  M->mb_size.assign((size_t)nb * 9, 0.0);
  M->mb_bcs.assign((size_t)nb * 6, 0);
  M->nghbr.assign((size_t)nb * 56 * 4, -1);
  const int nfx = M->multilevel ? 2 : 1, nfy = (M->multilevel && multi_d) ? 2 : 1, nfz = (M->multilevel && three_d) ? 2 : 1;
  (void)nfx;
  for (int g = 0; g < nb; ++g) {
    const LLoc &ll = M->lloc[g];
    const int l[3] = {ll.lx1, ll.lx2, ll.lx3};
    const bool act[3] = {true, multi_d, three_d};
    for (int d = 0; d < 3; ++d) {
      if (!act[d]) {      // meshblock.cpp:63-67,90-94: the whole mesh extent, mesh flags
        M->mb_size[g * 9 + d] = M->xmin[d]; M->mb_size[g * 9 + 3 + d] = M->xmax[d];
        M->mb_bcs[g * 6 + 2 * d] = M->bcs[2 * d]; M->mb_bcs[g * 6 + 2 * d + 1] = M->bcs[2 * d + 1];
      } else {
        const int n = M->nroot[d] << (ll.level - M->root_level);
        double lo, hi;
        if (l[d] == 0) { lo = M->xmin[d]; M->mb_bcs[g * 6 + 2 * d] = M->bcs[2 * d]; }
        else { lo = LeftEdgeX(l[d], n, M->xmin[d], M->xmax[d]); M->mb_bcs[g * 6 + 2 * d] = AK_BC_BLOCK; }
        if (l[d] == n - 1) { hi = M->xmax[d]; M->mb_bcs[g * 6 + 2 * d + 1] = M->bcs[2 * d + 1]; }
        else { hi = LeftEdgeX(l[d] + 1, n, M->xmin[d], M->xmax[d]); M->mb_bcs[g * 6 + 2 * d + 1] = AK_BC_BLOCK; }
        M->mb_size[g * 9 + d] = lo;
        M->mb_size[g * 9 + 3 + d] = hi;
      }
      M->mb_size[g * 9 + 6 + d] = (M->mb_size[g * 9 + 3 + d] - M->mb_size[g * 9 + d]) / static_cast<double>(M->nblock[d]);
    }
    // position of this block inside its parent
    const int myfx1 = ll.lx1 & 1, myfx2 = ll.lx2 & 1, myfx3 = ll.lx3 & 1;
    const int myox1 = myfx1 * 2 - 1, myox2 = multi_d ? myfx2 * 2 - 1 : 0, myox3 = three_d ? myfx3 * 2 - 1 : 0;
    auto set = [&](int slot, const Tree::Node *nt, int dest) {
      int *e = &M->nghbr[((size_t)g * 56 + slot) * 4];
      e[0] = nt->gid; e[1] = nt->l.level; e[2] = M->rank_of[nt->gid]; e[3] = dest;
    };
    using akb::neighbor_index;
    for (int oz = -1; oz <= 1; ++oz) for (int oy = -1; oy <= 1; ++oy) for (int ox = -1; ox <= 1; ++ox) {
      if (!ox && !oy && !oz) continue;
      if ((oy && !multi_d) || (oz && !three_d)) continue;
      Tree::Node *nt = T.FindNeighbor(ll, ox, oy, oz);
      if (!nt) continue;
      const int nzero = (ox == 0) + (oy == 0) + (oz == 0);      // 2: face, 1: edge, 0: corner
      if (nt->split) {   // finer neighbour(s): the leaves of nt that touch this block
        const int ffx = 1 - (ox + 1) / 2, ffy = 1 - (oy + 1) / 2, ffz = 1 - (oz + 1) / 2;   // 0 toward +, 1 toward -
        if (nzero == 2) {
          // sub-block indices (f1,f2) run over the two transverse directions in the order (x1|x2, x2|x3)
          for (int f2 = 0; f2 < ((ox || oy) ? nfz : nfy); ++f2)
            for (int f1 = 0; f1 < (ox ? nfy : nfx); ++f1) {
              const Tree::Node *nf = ox ? nt->leaf[ffx + 2 * f1 + 4 * f2] : oy ? nt->leaf[f1 + 2 * ffy + 4 * f2] : nt->leaf[f1 + 2 * f2 + 4 * ffz];
              set(neighbor_index(ox, oy, oz, f1, f2), nf, neighbor_index(-ox, -oy, -oz, f1, f2));
            }
        } else if (nzero == 1) {
          const int nf1 = (oz == 0) ? nfz : (oy == 0) ? nfy : nfy;   // meshblock.cpp:373 loops fx over nfy
          for (int f1 = 0; f1 < nf1; ++f1) {
            const Tree::Node *nf = (oz == 0) ? nt->leaf[ffx + 2 * ffy + 4 * f1] : (oy == 0) ? nt->leaf[ffx + 2 * f1 + 4 * ffz] : nt->leaf[f1 + 2 * ffy + 4 * ffz];
            set(neighbor_index(ox, oy, oz, f1, 0), nf, neighbor_index(-ox, -oy, -oz, f1, 0));
          }
        } else {
          set(neighbor_index(ox, oy, oz, 0, 0), nt->leaf[ffx + 2 * ffy + 4 * ffz], neighbor_index(-ox, -oy, -oz, 0, 0));
        }
      } else if (nt->l.level == ll.level) {   // same level
        set(neighbor_index(ox, oy, oz, 0, 0), nt, neighbor_index(-ox, -oy, -oz, 0, 0));
      } else {                                // coarser: slot and destination carry MY sub-block position
        if (nzero == 2) {
          const int f1 = ox ? myfx2 : myfx1, f2 = (ox || oy) ? myfx3 : myfx2;
          set(neighbor_index(ox, oy, oz, f1, f2), nt, neighbor_index(-ox, -oy, -oz, f1, f2));
        } else if (nzero == 1) {
          // only the exterior edges of the coarser block's face (meshblock.cpp:263,333,383)
          const int f1 = (oz == 0) ? myfx3 : (oy == 0) ? myfx2 : myfx1;
          const bool exterior = (oz == 0) ? (myox1 == ox && myox2 == oy) : (oy == 0) ? (myox1 == ox && myox3 == oz) : (myox2 == oy && myox3 == oz);
          if (exterior) set(neighbor_index(ox, oy, oz, f1, 0), nt, neighbor_index(-ox, -oy, -oz, f1, 0));
        } else if (myox1 == ox && myox2 == oy && myox3 == oz) {   // :410
          set(neighbor_index(ox, oy, oz, 0, 0), nt, neighbor_index(-ox, -oy, -oz, 0, 0));
        }
      }
    }
  }
}

static void free_mesh(HostMesh *M) { if (M) { delete M->tree; delete M; } }

// regions: [nregions][7] = physical level (>= 1), x1min, x1max, x2min, x2max, x3min, x3max of each
// <refined_region> block (build_tree.cpp:79-230); multilevel = <mesh_refinement>/refinement != none.
static int mesh_build(const int *nmesh, const int *nblock, const double *xmin, const double *xmax, const int *bcs,
                      int nranks, int multilevel, int nregions, const double *regions, void **mesh) {
  if (!mesh) { g_mesh_err = "ak_mesh_create: null output"; return 1; }
  *mesh = nullptr;
  auto *M = new HostMesh();
  for (int d = 0; d < 3; ++d) {
    M->nmesh[d] = nmesh[d]; M->nblock[d] = nblock[d]; M->xmin[d] = xmin[d]; M->xmax[d] = xmax[d];
    if (nblock[d] < 1 || nmesh[d] < nblock[d] || nmesh[d] % nblock[d] != 0) {
      g_mesh_err = "ak_mesh_create: mesh/nx must be a positive multiple of meshblock/nx"; free_mesh(M); return 1;
    }
    M->nroot[d] = nmesh[d] / nblock[d];                       // build_tree.cpp:34-36
    if ((bcs[2 * d] == AK_BC_PERIODIC) != (bcs[2 * d + 1] == AK_BC_PERIODIC)) {
      g_mesh_err = "ak_mesh_create: both inner and outer bcs must be periodic"; free_mesh(M); return 1;  // mesh.cpp:86-93
    }
  }
  for (int q = 0; q < 6; ++q) M->bcs[q] = bcs[q];
  M->nranks = nranks;
  M->multilevel = multilevel != 0 || nregions > 0;
  const bool multi_d = nmesh[1] > 1, three_d = nmesh[2] > 1;
  if (M->multilevel && (nblock[0] % 2 || (multi_d && nblock[1] % 2) || (three_d && nblock[2] % 2))) {
    g_mesh_err = "ak_mesh_create: number of cells in MeshBlock must be divisible by 2 with SMR or AMR"; free_mesh(M); return 1;   // build_tree.cpp:68-76
  }
  int nmbmax = std::max(M->nroot[0], std::max(M->nroot[1], M->nroot[2]));
  for (M->root_level = 0; ((1 << M->root_level) < nmbmax); M->root_level++) {}   // build_tree.cpp:42-45
  M->tree = new Tree(M->nroot, M->bcs, M->root_level, multi_d, three_d);
  Tree &T = *M->tree;
  M->max_level = M->root_level;
  for (int r = 0; r < nregions; ++r) {        // build_tree.cpp:79-230
    const double *R = regions + 7 * r;
    const int phy = (int)R[0], logl = phy + M->root_level;
    double rmin[3] = {R[1], multi_d ? R[3] : xmin[1], three_d ? R[5] : xmin[2]};
    double rmax[3] = {R[2], multi_d ? R[4] : xmax[1], three_d ? R[6] : xmax[2]};
    if (phy < 1) { g_mesh_err = "ak_mesh_create: <refined_region> level must be larger than 0"; free_mesh(M); return 1; }
    for (int d = 0; d < 3; ++d)
      if (rmin[d] > rmax[d] || rmin[d] < xmin[d] || rmax[d] > xmax[d]) {
        g_mesh_err = "ak_mesh_create: <refined_region> must be valid and fully contained within the root mesh"; free_mesh(M); return 1;
      }
    if (logl > M->max_level) M->max_level = logl;
    int lmin[3] = {0, 0, 0}, lmax[3] = {0, 0, 0};
    const bool act[3] = {true, multi_d, three_d};
    for (int d = 0; d < 3; ++d) {
      if (!act[d]) continue;
      const int lxmax = M->nroot[d] * (1 << phy);
      for (lmin[d] = 0; lmin[d] < lxmax; lmin[d]++) if (LeftEdgeX(lmin[d] + 1, lxmax, xmin[d], xmax[d]) > rmin[d]) break;
      for (lmax[d] = lmin[d]; lmax[d] < lxmax; lmax[d]++) if (LeftEdgeX(lmax[d] + 1, lxmax, xmin[d], xmax[d]) >= rmax[d]) break;
      if (lmin[d] % 2 == 1) lmin[d]--;
      if (lmax[d] % 2 == 0) lmax[d]++;
    }
    int nnew = 0;
    for (int k = lmin[2]; k < (three_d ? lmax[2] : lmin[2] + 1); k += 2)
      for (int j = lmin[1]; j < (multi_d ? lmax[1] : lmin[1] + 1); j += 2)
        for (int i = lmin[0]; i < lmax[0]; i += 2) T.AddNode(T.root, LLoc{i, multi_d ? j : 0, three_d ? k : 0, logl}, nnew);
  }
  T.ZOrder(T.root, M->lloc);
  M->nmb_total = (int)M->lloc.size();
  if (M->nmb_total < nranks) { g_mesh_err = "ak_mesh_create: fewer MeshBlocks than ranks"; free_mesh(M); return 1; }
  load_balance(*M);
  fill_tables(M, T);
  *mesh = M;
  return 0;
}

// nmesh/nblock: cells per direction in the mesh / per MeshBlock; bcs[6]: ix1,ox1,ix2,ox2,ix3,ox3.
int ak_mesh_create(const int *nmesh, const int *nblock, const double *xmin, const double *xmax, const int *bcs,
                   int nranks, void **mesh) {
  return mesh_build(nmesh, nblock, xmin, xmax, bcs, nranks, 0, 0, nullptr, mesh);
}

// Static mesh refinement (host maps only): the same tables for a mesh with
// <refined_region> blocks; NeighborBlock.lev and the sub-block slots / destinations are filled as in
// meshblock.cpp:142-430.  lloc[.][3] (ak_mesh_tables) is the logical level of each block.
int ak_mesh_create_smr(const int *nmesh, const int *nblock, const double *xmin, const double *xmax, const int *bcs,
                       int nranks, int nregions, const double *regions, void **mesh) {
  return mesh_build(nmesh, nblock, xmin, xmax, bcs, nranks, 1, nregions, regions, mesh);
}

int ak_mesh_destroy(void *mesh) { free_mesh(static_cast<HostMesh *>(mesh)); return 0; }

int ak_mesh_info(void *mesh, int *nmb_total, int *root_level) {
  auto *M = static_cast<HostMesh *>(mesh);
  if (!M) return 1;
  if (nmb_total) *nmb_total = M->nmb_total;
  if (root_level) *root_level = M->root_level;
  return 0;
}

// lloc [nmb_total][4], rank_of [nmb_total], gids_start/nmb_rank [nranks] (any may be NULL)
int ak_mesh_tables(void *mesh, int *lloc, int *rank_of, int *gids_start, int *nmb_rank) {
  auto *M = static_cast<HostMesh *>(mesh);
  if (!M) return 1;
  if (lloc) std::memcpy(lloc, M->lloc.data(), sizeof(LLoc) * M->lloc.size());
  if (rank_of) std::memcpy(rank_of, M->rank_of.data(), sizeof(int) * M->rank_of.size());
  if (gids_start) std::memcpy(gids_start, M->gids_start.data(), sizeof(int) * M->nranks);
  if (nmb_rank) std::memcpy(nmb_rank, M->nmb_rank.data(), sizeof(int) * M->nranks);
  return 0;
}

// per-rank slices: mb_size [nmb][9], mb_bcs [nmb][6], nghbr [nmb][56][4]
int ak_mesh_rank_tables(void *mesh, int rank, double *mb_size, int *mb_bcs, int *nghbr) {
  auto *M = static_cast<HostMesh *>(mesh);
  if (!M || rank < 0 || rank >= M->nranks) return 1;
  const size_t g0 = M->gids_start[rank], n = M->nmb_rank[rank];
  if (mb_size) std::memcpy(mb_size, &M->mb_size[g0 * 9], sizeof(double) * n * 9);
  if (mb_bcs) std::memcpy(mb_bcs, &M->mb_bcs[g0 * 6], sizeof(int) * n * 6);
  if (nghbr) std::memcpy(nghbr, &M->nghbr[g0 * 56 * 4], sizeof(int) * n * 56 * 4);
  return 0;
}

// ---- adaptive mesh refinement, host maps: mesh_refinement.cpp:46-131 (parameters),
// :196-266 (CheckForRefinement: cycle counters, level limits, refinement interval), :274-419 (UpdateMeshBlockTree:
// refine flagged leaves with the 2:1 rule, derefine complete sibling sets that the tree allows), :427-510
// (new Z-ordered list, new -> old and old -> new gid maps, per-old-block flags) and :611-656 (housekeeping).
// One rank's view of the whole mesh, as every rank of the reference holds it.
int ak_mesh_set_amr(void *mesh, int num_levels, int refinement_interval, int ncycle_check) {
  auto *M = static_cast<HostMesh *>(mesh);
  if (!M || !M->tree || num_levels < 1 || refinement_interval < 0 || ncycle_check < 1) return 1;
  M->adaptive = true;
  M->multilevel = true;
  M->max_level = std::max(M->max_level, M->root_level + num_levels - 1);      // build_tree.cpp:232-246
  M->refinement_interval = refinement_interval;
  M->ncyc_check_amr = ncycle_check;
  M->ncyc_since_ref.assign(M->nmb_total, 0);
  fill_tables(M, *M->tree);       // neighbour sub-block slots are in use on a multilevel mesh
  return 0;
}

// refine_flag [nmb_total]: in = the criteria's verdict per block (+1 refine, -1 derefine, 0), ignored on cycles that
// are not check cycles; out = what happened to each OLD block (+1 refined, -nleaf derefined, 0).  ncycle = the mesh's
// cycle counter when the reference calls AdaptiveMeshRefinement (after it was incremented).  Returns the number of
// blocks created / deleted; when either is non-zero the mesh tables are those of the NEW mesh afterwards.
int ak_mesh_adapt(void *mesh, int ncycle, int *refine_flag, int *nnew_out, int *ndel_out) {
  auto *M = static_cast<HostMesh *>(mesh);
  if (!M || !M->tree || !M->adaptive || !refine_flag) { g_mesh_err = "ak_mesh_adapt: mesh is not adaptive (ak_mesh_set_amr)"; return 1; }
  Tree &T = *M->tree;
  const int nb = M->nmb_total;
  const bool multi_d = M->nmesh[1] > 1, three_d = M->nmesh[2] > 1;
  const int nleaf = T.nleaf;
  if (nnew_out) *nnew_out = 0;
  if (ndel_out) *ndel_out = 0;
  for (int m = 0; m < nb; ++m) M->ncyc_since_ref[m] += 1;
  if (ncycle % M->ncyc_check_amr != 0) { std::fill(refine_flag, refine_flag + nb, 0); return 0; }
  for (int m = 0; m < nb; ++m) {
    if (M->lloc[m].level == M->max_level && refine_flag[m] > 0) refine_flag[m] = 0;
    if (M->lloc[m].level == M->root_level && refine_flag[m] < 0) refine_flag[m] = 0;
    if (M->ncyc_since_ref[m] < M->refinement_interval) refine_flag[m] = 0;
  }
  // UpdateMeshBlockTree
  std::vector<LLoc> llref, llderef, cllderef;
  for (int m = 0; m < nb; ++m) if (refine_flag[m] == 1) llref.push_back(M->lloc[m]);
  int tnderef = 0;
  for (int m = 0; m < nb; ++m) if (refine_flag[m] == -1) tnderef++;
  int nnew = 0, ndel = 0;
  if (!(llref.empty() && tnderef < nleaf)) {
    if (tnderef >= nleaf) for (int m = 0; m < nb; ++m) if (refine_flag[m] == -1) llderef.push_back(M->lloc[m]);
    const int lj = multi_d ? 1 : 0, lk = three_d ? 1 : 0;
    for (int n = 0; n < (int)llderef.size(); ++n) {
      if ((llderef[n].lx1 & 1) || (llderef[n].lx2 & 1) || (llderef[n].lx3 & 1)) continue;
      int r = n, rr = 0;
      for (int k = 0; k <= lk; ++k) for (int j = 0; j <= lj; ++j) for (int i = 0; i <= 1; ++i) {
        if (r < (int)llderef.size()) {
          if (llderef[n].lx1 + i == llderef[r].lx1 && llderef[n].lx2 + j == llderef[r].lx2 && llderef[n].lx3 + k == llderef[r].lx3 &&
              llderef[n].level == llderef[r].level) rr++;
          r++;
        }
      }
      if (rr == nleaf) cllderef.push_back(LLoc{llderef[n].lx1 >> 1, llderef[n].lx2 >> 1, llderef[n].lx3 >> 1, llderef[n].level - 1});
    }
    // mesh_refinement.cpp:393-395 sorts [first, last-1) by decreasing level: the last entry stays where it is
    if (cllderef.size() > 1) std::sort(cllderef.begin(), cllderef.end() - 1, [](const LLoc &a, const LLoc &b) { return a.level > b.level; });
    for (const LLoc &l : llref) { Tree::Node *bt = T.FindMeshBlock(l); if (bt) T.Refine(bt, nnew); }
    for (const LLoc &l : cllderef) { Tree::Node *bt = T.FindMeshBlock(l); if (bt && bt->split) T.Derefine(bt, ndel); }
  }
  if (nnew == 0 && ndel == 0) { std::fill(refine_flag, refine_flag + nb, 0); return 0; }
  // RedistAndRefineMeshBlocks, steps 1-3 and 10
  std::vector<LLoc> new_lloc;
  std::vector<int> newtoold;
  T.ZOrder(T.root, new_lloc, &newtoold);
  const int new_nmb = (int)new_lloc.size();
  if (new_nmb != nb + nnew - ndel) { g_mesh_err = "ak_mesh_adapt: number of MeshBlocks in the new tree differs from the expected value"; return 1; }
  std::vector<int> oldtonew(nb, 0);
  int mb_idx = 1;
  for (int n = 1; n < new_nmb; ++n) {
    if (newtoold[n] == newtoold[n - 1] + 1) oldtonew[mb_idx++] = n;
    else if (newtoold[n] == newtoold[n - 1] + nleaf) { for (int j = 0; j < nleaf - 1; ++j) oldtonew[mb_idx++] = n - 1; oldtonew[mb_idx++] = n; }
  }
  while (mb_idx < nb) oldtonew[mb_idx++] = new_nmb - 1;
  for (int oldm = 0; oldm < nb; ++oldm) {
    const int ol = M->lloc[oldm].level, nl = new_lloc[oldtonew[oldm]].level;
    refine_flag[oldm] = (ol > nl) ? -nleaf : (ol < nl) ? 1 : 0;
  }
  std::vector<int> ncyc(new_nmb);
  for (int m = 0; m < new_nmb; ++m) ncyc[m] = refine_flag[newtoold[m]] != 0 ? 0 : M->ncyc_since_ref[newtoold[m]];
  M->old_flag.assign(refine_flag, refine_flag + nb);
  M->newtoold = newtoold; M->oldtonew = oldtonew;
  M->ncyc_since_ref = ncyc;
  M->lloc = new_lloc;
  M->nmb_total = new_nmb;
  load_balance(*M);
  fill_tables(M, T);
  if (nnew_out) *nnew_out = nnew;
  if (ndel_out) *ndel_out = ndel;
  return 0;
}

// maps of the last ak_mesh_adapt that changed the mesh: newtoold [new nmb_total], oldtonew / old_flag [old nmb_total]
int ak_mesh_amr_maps(void *mesh, int *newtoold, int *oldtonew, int *old_flag, int *old_nmb) {
  auto *M = static_cast<HostMesh *>(mesh);
  if (!M) return 1;
  if (newtoold) std::memcpy(newtoold, M->newtoold.data(), sizeof(int) * M->newtoold.size());
  if (oldtonew) std::memcpy(oldtonew, M->oldtonew.data(), sizeof(int) * M->oldtonew.size());
  if (old_flag) std::memcpy(old_flag, M->old_flag.data(), sizeof(int) * M->old_flag.size());
  if (old_nmb) *old_nmb = (int)M->oldtonew.size();
  return 0;
}

int ak_neighbor_index(int ix, int iy, int iz, int n1, int n2) { return akb::neighbor_index(ix, iy, iz, n1, n2); }

// Blast problem generator, pgen/fluids/blast.cpp:147-158,236-374 (MHD branch, Cartesian).
// Fills the ACTIVE zones of the rank's blocks in host arrays laid out like the device arrays
// (ghosts are left to the halo exchange, as in the reference).  prm = {pi_amb, di_amb, prat,
// drat, b_amb, inner_radius, outer_radius}.
int ak_pgen_blast(void *mesh, int rank, int ng, double gamma, const double *prm, double *u0, double *b1f,
                  double *b2f, double *b3f) {
  auto *M = static_cast<HostMesh *>(mesh);
  if (!M || rank < 0 || rank >= M->nranks) return 1;
  const double pi_amb = prm[0], di_amb = prm[1], prat = prm[2], drat = prm[3], b_amb = prm[4];
  const double rout = prm[6], rin = rout - prm[5];
  const int nx1 = M->nblock[0], nx2 = M->nblock[1], nx3 = M->nblock[2];
  const int n1 = nx1 + 2 * ng, n2 = nx2 + 2 * ng, n3 = nx3 + 2 * ng;
  const int is = ng, ie = ng + nx1 - 1, js = ng, je = ng + nx2 - 1, ks = ng, ke = ng + nx3 - 1;
  const size_t cc = (size_t)n3 * n2 * n1, f1 = (size_t)n3 * n2 * (n1 + 1), f2 = (size_t)n3 * (n2 + 1) * n1, f3 = (size_t)(n3 + 1) * n2 * n1;
  const double gm1 = gamma - 1.0;
  const int nmb = M->nmb_rank[rank], g0 = M->gids_start[rank];
  std::vector<double> a3((size_t)n3 * n2 * n1);
  for (int m = 0; m < nmb; ++m) {
    const double *sz = &M->mb_size[(size_t)(g0 + m) * 9];
    const double x1min = sz[0], x2min = sz[1], x3min = sz[2], x1max = sz[3], x2max = sz[4], x3max = sz[5];
    const double dx1 = sz[6], dx2 = sz[7];
    double *U = u0 + (size_t)m * 5 * cc;
    std::vector<double> bcx(cc, 0.0), bcy(cc, 0.0), bcz(cc, 0.0), wd(cc, 0.0), we(cc, 0.0);
    auto C = [&](int k, int j, int i) { return ((size_t)k * n2 + j) * n1 + i; };
    for (int k = ks; k <= ke; ++k) for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) {
      double x1v = CellCenterX(i - is, nx1, x1min, x1max);
      double x2v = CellCenterX(j - js, nx2, x2min, x2max);
      double x3v = CellCenterX(k - ks, nx3, x3min, x3max);
      double rad = std::sqrt(x1v * x1v + x2v * x2v + x3v * x3v);
      double den = di_amb, pres = pi_amb;
      if (rad < rout) {
        if (rad < rin) { den *= drat; pres *= prat; }
        else {
          double f = (rad - rin) / (rout - rin);
          double log_den = (1.0 - f) * std::log(drat * di_amb) + f * std::log(di_amb);
          den = std::exp(log_den);
          double log_pres = (1.0 - f) * std::log(prat * pi_amb) + f * std::log(pi_amb);
          pres = std::exp(log_pres);
        }
      }
      wd[C(k, j, i)] = den;
      we[C(k, j, i)] = pres / gm1;
    }
    // vector potential a3 = b_amb * y on faces (blast.cpp:318-351)
    for (int k = ks; k <= ke + 1; ++k) for (int j = js; j <= je + 1; ++j) for (int i = is; i <= ie + 1; ++i) {
      double x2f = LeftEdgeX(j - js, nx2, x2min, x2max);
      if (k < n3) a3[C(k, j, i)] = b_amb * x2f;
    }
    double *B1 = b1f + (size_t)m * f1, *B2 = b2f + (size_t)m * f2, *B3 = b3f + (size_t)m * f3;
    auto X1 = [&](int k, int j, int i) -> double & { return B1[((size_t)k * n2 + j) * (n1 + 1) + i]; };
    auto X2 = [&](int k, int j, int i) -> double & { return B2[((size_t)k * (n2 + 1) + j) * n1 + i]; };
    auto X3 = [&](int k, int j, int i) -> double & { return B3[((size_t)k * n2 + j) * n1 + i]; };
    for (int k = ks; k <= ke; ++k) for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) {
      X1(k, j, i) = (a3[C(k, j + 1, i)] - a3[C(k, j, i)]) / dx2;
      X2(k, j, i) = -(a3[C(k, j, i + 1)] - a3[C(k, j, i)]) / dx1;
      X3(k, j, i) = 0.0;
      if (i == ie) X1(k, j, i + 1) = (a3[C(k, j + 1, i + 1)] - a3[C(k, j, i + 1)]) / dx2;
      if (j == je) X2(k, j + 1, i) = -(a3[C(k, j + 1, i + 1)] - a3[C(k, j + 1, i)]) / dx1;
      if (k == ke) X3(k + 1, j, i) = 0.0;
    }
    for (int k = ks; k <= ke; ++k) for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) {
      const size_t o = C(k, j, i);
      double bx = 0.5 * (X1(k, j, i) + X1(k, j, i + 1));
      double by = 0.5 * (X2(k, j, i) + X2(k, j + 1, i));
      double bz = 0.5 * (X3(k, j, i) + X3(k + 1, j, i));
      // SingleP2C_IdealMHD, eos/ideal_c2p_mhd.hpp:75-84, with v = 0
      const double d = wd[o], vx = 0.0, vy = 0.0, vz = 0.0;
      U[0 * cc + o] = d;
      U[1 * cc + o] = d * vx;
      U[2 * cc + o] = d * vy;
      U[3 * cc + o] = d * vz;
      U[4 * cc + o] = we[o] + 0.5 * (d * (vx * vx + vy * vy + vz * vz) + (bx * bx + by * by + bz * bz));
    }
  }
  return 0;
}

// Orszag-Tang vortex, pgen/tests/orszag_tang.cpp:32-123 (x3-invariant; same libm sin/cos as the
// reference, so the ICs are bit-identical).  Active zones of the rank's blocks.
int ak_pgen_orszag_tang(void *mesh, int rank, int ng, double gamma, double *u0, double *b1f, double *b2f,
                        double *b3f) {
  auto *M = static_cast<HostMesh *>(mesh);
  if (!M || rank < 0 || rank >= M->nranks) return 1;
  const double B0 = 1.0 / std::sqrt(4.0 * M_PI);
  const double d0 = 25.0 / (36.0 * M_PI);
  const double v0 = 1.0;
  const double p0 = 5.0 / (12.0 * M_PI);
  auto A3 = [](double x1, double x2, double B0_) {
    return (B0_ / (4.0 * M_PI)) * (std::cos(4.0 * M_PI * x1) - 2.0 * std::cos(2.0 * M_PI * x2));
  };
  const int nx1 = M->nblock[0], nx2 = M->nblock[1], nx3 = M->nblock[2];
  const int n1 = nx1 + 2 * ng, n2 = nx2 + 2 * ng, n3 = nx3 + 2 * ng;
  const int is = ng, ie = ng + nx1 - 1, js = ng, je = ng + nx2 - 1, ks = ng, ke = ng + nx3 - 1;
  const size_t cc = (size_t)n3 * n2 * n1, f1 = (size_t)n3 * n2 * (n1 + 1), f2 = (size_t)n3 * (n2 + 1) * n1, f3 = (size_t)(n3 + 1) * n2 * n1;
  const double gm1 = gamma - 1.0;
  const int nmb = M->nmb_rank[rank], g0 = M->gids_start[rank];
  for (int m = 0; m < nmb; ++m) {
    const double *sz = &M->mb_size[(size_t)(g0 + m) * 9];
    const double x1min = sz[0], x2min = sz[1], x1max = sz[3], x2max = sz[4], dx1 = sz[6], dx2 = sz[7];
    double *U = u0 + (size_t)m * 5 * cc, *B1 = b1f + (size_t)m * f1, *B2 = b2f + (size_t)m * f2, *B3 = b3f + (size_t)m * f3;
    auto C = [&](int k, int j, int i) { return ((size_t)k * n2 + j) * n1 + i; };
    auto X1 = [&](int k, int j, int i) -> double & { return B1[((size_t)k * n2 + j) * (n1 + 1) + i]; };
    auto X2 = [&](int k, int j, int i) -> double & { return B2[((size_t)k * (n2 + 1) + j) * n1 + i]; };
    auto X3 = [&](int k, int j, int i) -> double & { return B3[((size_t)k * n2 + j) * n1 + i]; };
    for (int k = ks; k <= ke; ++k) for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) {
      double x1v = CellCenterX(i - is, nx1, x1min, x1max);
      double x2v = CellCenterX(j - js, nx2, x2min, x2max);
      const size_t o = C(k, j, i);
      U[0 * cc + o] = d0;
      U[1 * cc + o] = d0 * v0 * std::sin(2.0 * M_PI * x2v);
      U[2 * cc + o] = -d0 * v0 * std::sin(2.0 * M_PI * x1v);
      U[3 * cc + o] = 0.0;
      double x1f = LeftEdgeX(i - is, nx1, x1min, x1max), x1fp1 = LeftEdgeX(i + 1 - is, nx1, x1min, x1max);
      double x2f = LeftEdgeX(j - js, nx2, x2min, x2max), x2fp1 = LeftEdgeX(j + 1 - js, nx2, x2min, x2max);
      X1(k, j, i) = (A3(x1f, x2fp1, B0) - A3(x1f, x2f, B0)) / dx2;
      X2(k, j, i) = -(A3(x1fp1, x2f, B0) - A3(x1f, x2f, B0)) / dx1;
      X3(k, j, i) = 0.0;
      if (i == ie) X1(k, j, i + 1) = (A3(x1fp1, x2fp1, B0) - A3(x1fp1, x2f, B0)) / dx2;
      if (j == je) X2(k, j + 1, i) = -(A3(x1fp1, x2fp1, B0) - A3(x1f, x2fp1, B0)) / dx1;
      if (k == ke) X3(k + 1, j, i) = 0.0;
    }
    for (int k = ks; k <= ke; ++k) for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) {
      const size_t o = C(k, j, i);
      U[4 * cc + o] = p0 / gm1 + (0.5 / U[0 * cc + o]) * (U[1 * cc + o] * U[1 * cc + o] + U[2 * cc + o] * U[2 * cc + o] + U[3 * cc + o] * U[3 * cc + o]) +
                      0.5 * ((0.5 * (X1(k, j, i) + X1(k, j, i + 1))) * (0.5 * (X1(k, j, i) + X1(k, j, i + 1))) +
                             (0.5 * (X2(k, j, i) + X2(k, j + 1, i))) * (0.5 * (X2(k, j, i) + X2(k, j + 1, i))) +
                             (0.5 * (X3(k, j, i) + X3(k + 1, j, i))) * (0.5 * (X3(k, j, i) + X3(k + 1, j, i))));
    }
  }
  return 0;
}

}  // extern "C"
