"""Static mesh refinement, host maps (north_star: "bit-exact for MeshBlock index/neighbor
maps").  csrc/host_mesh.cpp mirrors MeshBlockTree (AddNode / Refine with the 2:1 level constraint, Z-ordered gids
across levels) and MeshBlock::SetNeighbors with coarser / finer neighbours.  Gates:
  * the Z-ordered LogicalLocation list == the unmodified reference's (header of its restart dump) for 1-D, 2-D and
    3-D decks with nested <refined_region> blocks, periodic and non-periodic boundaries -- committed fixtures
    (tests/golden/smr_trees.npz, tests/gen_golden.py smr_trees()) and a live run when oracle/_ref is present;
  * the 56-slot neighbour table (gid, level, rank, destination slot) == a brute-force restatement of
    src/mesh/meshblock.cpp:142-430 that looks leaves up in a dictionary instead of walking a tree;
  * block sizes / BC flags follow src/mesh/meshblock.cpp:25-131 at every level, adjacent leaves differ by at most
    one level, and the uniform-mesh entry point is unchanged."""
import os
import tempfile

import numpy as np
import pytest

import hydro_util as H
import oracle_lib as O
from athenak_b200.mesh import Mesh

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "smr_trees.npz"))
NAMES = sorted(k[:-5] for k in GOLD.files if k.endswith("_lloc"))


def mesh_of(deck_text, nranks=1):
    p = H.deck_params(deck_text)
    nmesh = tuple(int(p[f"mesh/nx{d}"]) for d in (1, 2, 3))
    nblock = tuple(int(p[f"meshblock/nx{d}"]) for d in (1, 2, 3))
    xmin = tuple(float(p[f"mesh/x{d}min"]) for d in (1, 2, 3))
    xmax = tuple(float(p[f"mesh/x{d}max"]) for d in (1, 2, 3))
    bcs = tuple(p[f"mesh/{s}x{d}_bc"] for d in (1, 2, 3) for s in ("i", "o"))
    regs = []
    for blk in sorted({k.split("/")[0] for k in p if k.startswith("refined_region")}):
        r = [float(p[blk + "/level"])]
        for d in (1, 2, 3):
            r += [float(p.get(f"{blk}/x{d}min", xmin[d - 1])), float(p.get(f"{blk}/x{d}max", xmax[d - 1]))]
        regs.append(r)
    return Mesh(nmesh, nblock, xmin=xmin, xmax=xmax, bcs=bcs, nranks=nranks, refined_regions=regs), bcs


@pytest.mark.parametrize("name", NAMES)
def test_smr_tree_matches_reference_fixture(name):
    mesh, _ = mesh_of(str(GOLD[name + "_deck"]))
    assert mesh.root_level == int(GOLD[name + "_root"])
    assert np.array_equal(mesh.lloc, GOLD[name + "_lloc"])


def test_smr_tree_matches_live_reference():
    if not O.have_reference("athena_builtin"):
        pytest.skip("oracle/_ref/athena_builtin not built")
    import struct
    deck = os.path.join(os.path.dirname(os.path.abspath(__file__)), "decks", "lwave3d_hydro_smr.athinput")
    ov = ["time/nlim=0", "mesh/nx1=64", "meshblock/nx1=8", "refined_region1/level=2", "refined_region1/x1min=0.0",
          "refined_region1/x1max=0.2", "refined_region2/x3min=1.45", "refined_region2/x3max=1.5"]
    with tempfile.TemporaryDirectory() as tmp:
        O.run_reference("athena_builtin", deck, tmp, ov)
        raw = open(os.path.join(tmp, "rst", "LinWave.00000.rst"), "rb").read()
    end = raw.index(b"<par_end>\n") + 10
    nmb, root = struct.unpack_from("<ii", raw, end)
    lloc = np.frombuffer(raw, "<i4", 4 * nmb, end + 8 + 72 + 76 + 76 + 16 + 4).reshape(nmb, 4)
    mesh, _ = mesh_of(raw[:end].decode())
    assert mesh.root_level == root and np.array_equal(mesh.lloc, lloc)


def brute_force_neighbors(mesh, bcs):
    """meshblock.cpp:142-430 restated with a {(lx1,lx2,lx3,level): gid} dictionary."""
    L = O.lib()
    idx = lambda ox, oy, oz, f1, f2: L.oracle_neighbor_index(ox, oy, oz, f1, f2)
    periodic = [O.BC[b] == O.BC["periodic"] for b in bcs]
    leaf = {tuple(int(v) for v in l): g for g, l in enumerate(mesh.lloc)}
    nroot = [mesh.nmesh[d] // mesh.nblock[d] for d in range(3)]
    act = [True, mesh.nmesh[1] > 1, mesh.nmesh[2] > 1]
    out = -np.ones((mesh.nmb_total, 56, 4), np.int32)
    for g, (lx, ly, lz, lev) in enumerate(mesh.lloc):
        me = (int(lx), int(ly), int(lz))
        par = [c & 1 for c in me]
        for oz in (-1, 0, 1):
            for oy in (-1, 0, 1):
                for ox in (-1, 0, 1):
                    o = (ox, oy, oz)
                    if o == (0, 0, 0) or any(o[d] and not act[d] for d in range(3)):
                        continue
                    q, ok = [me[d] + o[d] for d in range(3)], True
                    for d in range(3):
                        n = nroot[d] << (int(lev) - mesh.root_level) if act[d] else 1
                        if q[d] < 0 or q[d] >= n:
                            if periodic[2 * d]:
                                q[d] %= n
                            else:
                                ok = False
                    if not ok:
                        continue
                    free = [d for d in range(3) if o[d] == 0 and act[d]]      # transverse directions, x1 before x2 before x3
                    nz = [d for d in range(3) if o[d] != 0]
                    def put(slot, ng, dest):
                        out[g, slot] = (ng, mesh.lloc[ng, 3], mesh.rank_of[ng], dest)
                    if (q[0], q[1], q[2], int(lev)) in leaf:                         # same level
                        put(idx(*o, 0, 0), leaf[(q[0], q[1], q[2], int(lev))], idx(-ox, -oy, -oz, 0, 0))
                    elif (q[0] >> 1, q[1] >> 1, q[2] >> 1, int(lev) - 1) in leaf:    # coarser
                        ng = leaf[(q[0] >> 1, q[1] >> 1, q[2] >> 1, int(lev) - 1)]
                        if len(nz) == 1:
                            t = [d for d in range(3) if o[d] == 0]
                            f1, f2 = par[t[0]], par[t[1]]
                            put(idx(*o, f1, f2), ng, idx(-ox, -oy, -oz, f1, f2))
                        else:        # edges / corners: only those on the outside of the parent
                            if all(par[d] * 2 - 1 == o[d] for d in nz):
                                f1 = par[[d for d in range(3) if o[d] == 0][0]] if len(nz) == 2 else 0
                                put(idx(*o, f1, 0), ng, idx(-ox, -oy, -oz, f1, 0))
                    else:                                                            # finer
                        fs = [(0, 0)] if not free else [(a, 0) for a in (0, 1)] if len(free) == 1 else [(a, b) for b in (0, 1) for a in (0, 1)]
                        if len(nz) == 2 and len(free) == 0:
                            fs = [(0, 0)]
                        for f1, f2 in fs:
                            c = [2 * q[d] + (1 if o[d] < 0 else 0) for d in range(3)]
                            for fv, d in zip((f1, f2), free):
                                c[d] = 2 * q[d] + fv
                            for d in range(3):
                                if not act[d]:
                                    c[d] = 0
                            ng = leaf[(c[0], c[1], c[2], int(lev) + 1)]
                            put(idx(*o, f1, f2), ng, idx(-ox, -oy, -oz, f1, f2))
    return out


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("nranks", [1, 3])
def test_smr_neighbor_table_matches_definition(name, nranks):
    mesh, bcs = mesh_of(str(GOLD[name + "_deck"]), nranks=nranks)
    want = brute_force_neighbors(mesh, bcs)
    got = np.concatenate([mesh.rank_tables(r)[2] for r in range(nranks)])
    assert got.shape == want.shape
    bad = np.argwhere((got != want).any(axis=2))
    assert len(bad) == 0, f"first mismatch (gid, slot) = {bad[0]}: got {got[tuple(bad[0])]}, want {want[tuple(bad[0])]}"
    # every link is mutual: the destination slot of my entry, on the neighbour, points back at me
    for g in range(mesh.nmb_total):
        for s in range(56):
            ng, lev, rk, dest = got[g, s]
            if ng < 0:
                continue
            assert abs(int(lev) - int(mesh.lloc[g, 3])) <= 1            # 2:1 level constraint
            if lev == mesh.lloc[g, 3] and ng != g:
                assert got[ng, dest, 0] == g


@pytest.mark.parametrize("name", NAMES)
def test_smr_block_sizes(name):
    """meshblock.cpp:25-131: edges from LeftEdgeX at the block's own level, mesh flags on the mesh edge."""
    mesh, bcs = mesh_of(str(GOLD[name + "_deck"]))
    mbs, mbb, _ = mesh.rank_tables(0)
    for g, (lx, ly, lz, lev) in enumerate(mesh.lloc):
        for d, l in enumerate((lx, ly, lz)):
            if mesh.nmesh[d] == 1:
                assert mbs[g, d] == mesh.xmin[d] and mbs[g, 3 + d] == mesh.xmax[d]
                continue
            n = (mesh.nmesh[d] // mesh.nblock[d]) << (int(lev) - mesh.root_level)
            lo = mesh.xmin[d] if l == 0 else O._left_edge(int(l), n, np.float64(mesh.xmin[d]), np.float64(mesh.xmax[d]))
            hi = mesh.xmax[d] if l == n - 1 else O._left_edge(int(l) + 1, n, np.float64(mesh.xmin[d]), np.float64(mesh.xmax[d]))
            assert mbs[g, d] == lo and mbs[g, 3 + d] == hi and mbs[g, 6 + d] == (hi - lo) / np.float64(mesh.nblock[d])
            assert mbb[g, 2 * d] == (O.BC[bcs[2 * d]] if l == 0 else O.BC["block"])
            assert mbb[g, 2 * d + 1] == (O.BC[bcs[2 * d + 1]] if l == n - 1 else O.BC["block"])


def test_smr_rejects_odd_blocks():
    with pytest.raises(Exception):
        Mesh((30, 30, 30), (15, 15, 15), refined_regions=[(1, -0.1, 0.1, -0.1, 0.1, -0.1, 0.1)])


@pytest.mark.parametrize("name", NAMES)
def test_smr_plan_descriptor_counts(name):
    """ak_smr_create (host-side plan, no device needed): one exchange descriptor per neighbour entry, one fill
    descriptor per same-level entry (multi-D), one prolongation descriptor per coarser entry, one flux descriptor per
    coarser FACE entry -- and the totals are symmetric: every fine->coarse flux box has a finer entry on the other side."""
    import ctypes as C
    from athenak_b200 import lib as L
    lib = L.load()
    mesh, bcs = mesh_of(str(GOLD[name + "_deck"]))
    p = H.deck_params(str(GOLD[name + "_deck"]))
    _, _, ngh = mesh.rank_tables(0)
    lev = np.ascontiguousarray(mesh.lloc[:, 3], np.int32)
    ng = 4
    nb = mesh.nblock
    if any(n > 1 and n < 2 * ng for n in nb):
        ng = 2
    cfg = L.MhdCfg(mesh.nmb_total, ng, *nb, L.RECON["plm"], L.HYD_RSOLVER["hllc"], 1.4, 0, 0, 0, 0, 0)
    plan = C.c_void_p()
    ipp = C.POINTER(C.c_int)
    rc = lib.ak_smr_create(C.byref(cfg), ngh.ctypes.data_as(ipp), lev.ctypes.data_as(ipp), 0, 0, C.byref(plan))
    assert rc == 0, lib.ak_smr_last_error(None)
    cnt = (C.c_long * 4)()
    assert lib.ak_smr_counts(plan, cnt) == 0
    have = ngh[:, :, 0] >= 0
    same = have & (ngh[:, :, 1] == lev[:, None])
    coarser = have & (ngh[:, :, 1] < lev[:, None])
    finer = have & (ngh[:, :, 1] > lev[:, None])
    face = np.zeros(56, bool); face[:16] = True; face[24:32] = True
    assert cnt[0] == have.sum()
    assert cnt[1] == (same.sum() if mesh.nmesh[1] > 1 else 0)
    assert cnt[2] == coarser.sum()
    assert cnt[3] == (coarser & face[None, :]).sum() == (finer & face[None, :]).sum()
    lib.ak_smr_destroy(plan)
