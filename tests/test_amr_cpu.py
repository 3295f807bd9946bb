"""Adaptive mesh refinement, host maps (dynamic refinement, first part): the refinement criteria
(oracle/amr_oracle.c) and the MeshRefinement bookkeeping of the C++ host mesh (ak_mesh_set_amr / ak_mesh_adapt:
cycle counters, refinement interval, level limits, tree refinement with the 2:1 rule, derefinement of complete sibling
sets, Z-ordered block list, new->old / old->new maps) against an UNMODIFIED reference run that dumps a restart file
every cycle: at each cycle the criteria are evaluated on the reference's own data, the mesh is adapted, and the block
list must equal the one in the header of the reference's next dump.  Integer maps: bit-exact.  The run is made
here when oracle/_ref/athena_blast is present; the committed fixture tests/golden/amr_blast2d_llocs.npz holds the
same sequence (block lists + density per cycle) made by tests/gen_golden.py --amr-only."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

import golden_io as G
import hydro_util as H
import oracle_lib as O
from athenak_b200.mesh import Mesh
from athenak_b200.rst import read_rst

HERE = os.path.dirname(os.path.abspath(__file__))
DECK = os.path.join(HERE, "decks", "blast2d_amr.athinput")
METHOD = {"min_max": 0, "slope": 1, "second_deriv": 2, "location": 3}


def reference_sequence(overrides=(), threads=4):
    """[(lloc, u0 ghost-inclusive)] for cycle 0, 1, ... of a reference run of DECK."""
    with tempfile.TemporaryDirectory() as tmp:
        O.run_reference("athena_blast", DECK, tmp, list(overrides), threads=threads)
        files = sorted(f for f in os.listdir(os.path.join(tmp, "rst")) if f.endswith(".rst"))
        seq = []
        for f in files:
            r = read_rst(os.path.join(tmp, "rst", f))
            seq.append((r.ncycle, r.lloc.copy(), r.u0[:, 0].copy()))
    out, seen = [], set()
    for n, l, d in seq:          # the final dump repeats the last cycle
        if n not in seen:
            seen.add(n)
            out.append((l, d))
    return out


def follow(p, seq):
    """Evaluate the criteria on each cycle's data, adapt the mesh, compare with the next cycle's block list."""
    nmesh = tuple(int(p[f"mesh/nx{d}"]) for d in (1, 2, 3))
    nblock = tuple(int(p[f"meshblock/nx{d}"]) for d in (1, 2, 3))
    xmin = tuple(float(p[f"mesh/x{d}min"]) for d in (1, 2, 3))
    xmax = tuple(float(p[f"mesh/x{d}max"]) for d in (1, 2, 3))
    mesh = Mesh(nmesh, nblock, xmin=xmin, xmax=xmax)
    mesh.set_amr(int(p["mesh_refinement/num_levels"]), int(p.get("mesh_refinement/refinement_interval", 5)),
                 int(p.get("mesh_refinement/ncycle_check", 1)))
    ng = int(p["mesh/nghost"])
    L = O.lib()
    crit = sorted({k.split("/")[0] for k in p if k.startswith("amr_criterion")})
    assert np.array_equal(mesh.lloc, seq[0][0])
    nchanged = ncreated = ndeleted = 0
    for k in range(1, len(seq) - 1):
        lloc, dens = seq[k]
        assert np.array_equal(mesh.lloc, lloc), f"block list differs at cycle {k}"
        cfg = O.OCfg(mesh.nmb_total, ng, *nblock, O.RECON["plm"], O.RSOLVER["hlld"], 5.0 / 3.0, O.FLT_MIN, O.FLT_MIN, O.FLT_MIN,
                     O.FLT_MIN, O.FLT_MAX, 0, 0, 1.0, 0)
        flag = np.zeros(mesh.nmb_total, np.int32)
        mbs = np.ascontiguousarray(mesh.rank_tables(0)[0])
        q = np.ascontiguousarray(dens)             # mhd_w_d: the primitive density equals the conserved one (no floor fires)
        for blk in crit:
            loc = np.array([float(p.get(f"{blk}/location_x{d}", 0.0)) for d in (1, 2, 3)] + [float(p.get(f"{blk}/location_rad", 0.0))])
            L.oracle_amr_criterion(C.byref(cfg), METHOD[p[f"{blk}/method"]], O.p(q), 1, 0, C.c_double(float(p.get(f"{blk}/value_min", -O.FLT_MAX))),
                                   C.c_double(float(p.get(f"{blk}/value_max", O.FLT_MAX))), O.p(loc), O.p(mbs), O.ip(flag))
        nnew, ndel, what = mesh.adapt(k, flag)
        if nnew or ndel:
            nchanged += 1
            ncreated += nnew
            ndeleted += ndel
            assert mesh.nmb_total == len(what) + nnew - ndel
            # maps are consistent with the two block lists
            for n, o in enumerate(mesh.newtoold):
                lo, ln = lloc[o], mesh.lloc[n]
                sh = int(ln[3]) - int(lo[3])
                if sh == 0:
                    assert tuple(lo) == tuple(ln)
                elif sh == 1:       # refined: the old block is the parent
                    assert tuple(ln[:3] >> 1) == tuple(lo[:3])
                else:               # derefined: the old block is the first child
                    assert sh == -1 and tuple(lo[:3] >> 1) == tuple(ln[:3])
        assert np.array_equal(mesh.lloc, seq[k + 1][0]), f"block list after the AMR step of cycle {k} differs"
    return nchanged, ncreated, ndeleted


def test_amr_block_lists_match_fixture():
    d = G.load("amr_blast2d_llocs")
    p = H.deck_params(str(d["deck"]))
    n = int(d["ncycles"])
    seq = [(d[f"lloc_{k}"], d[f"dens_{k}"]) for k in range(n)]
    nchanged, ncreated, ndeleted = follow(p, seq)
    assert nchanged >= 3 and ncreated > 0 and ndeleted > 0


@pytest.mark.parametrize("name,overrides", [
    ("levels3", ("mesh_refinement/num_levels=3", "time/nlim=40")),
    ("second_deriv", ("amr_criterion0/method=second_deriv", "amr_criterion0/value_max=0.05", "time/nlim=25", "mesh_refinement/refinement_interval=2")),
    ("min_max", ("amr_criterion0/method=min_max", "amr_criterion0/value_max=1.02", "time/nlim=40")),
    ("location", ("amr_criterion0/method=location", "amr_criterion0/location_x1=0.21", "amr_criterion0/location_x2=-0.13",
                  "amr_criterion0/location_rad=0.06", "time/nlim=8", "mesh_refinement/num_levels=3", "mesh_refinement/refinement_interval=1")),
])
def test_amr_block_lists_match_live_reference(name, overrides):
    if not O.have_reference("athena_blast"):
        pytest.skip("oracle/_ref/athena_blast not built")
    p = H.deck_params(open(DECK).read())
    for o in overrides:
        k, v = o.split("=", 1)
        p[k] = v
    seq = reference_sequence(overrides)
    nchanged, ncreated, ndeleted = follow(p, seq)
    assert nchanged >= 1 and ncreated > 0


def test_amr_adapt_with_two_ranks_keeps_the_partition_consistent():
    """ak_mesh_adapt on a mesh distributed over two ranks (every rank holds the whole tree, as in the reference): after each
    change the Z-ordered list is re-partitioned by LoadBalance (load_balance.cpp:38-88, unit costs) -- contiguous gid ranges,
    sizes within one block of each other -- and every neighbour entry carries the rank that owns that gid."""
    d = G.load("amr_blast2d_llocs")
    p = H.deck_params(str(d["deck"]))
    nmesh = tuple(int(p[f"mesh/nx{q}"]) for q in (1, 2, 3))
    nblock = tuple(int(p[f"meshblock/nx{q}"]) for q in (1, 2, 3))
    mesh = Mesh(nmesh, nblock, nranks=2)
    mesh.set_amr(int(p["mesh_refinement/num_levels"]), int(p["mesh_refinement/refinement_interval"]), 1)
    L = O.lib()
    nchanged = 0
    for k in range(1, int(d["ncycles"]) - 1):
        cfg = O.OCfg(mesh.nmb_total, int(p["mesh/nghost"]), *nblock, 1, 2, 5.0 / 3.0, O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MAX, 0, 0, 1.0, 0)
        flag = np.zeros(mesh.nmb_total, np.int32)
        mbs = np.concatenate([mesh.rank_tables(r)[0] for r in range(2)])
        q = np.ascontiguousarray(d[f"dens_{k}"][:, None])
        L.oracle_amr_criterion(C.byref(cfg), 1, O.p(q), 1, 0, C.c_double(-O.FLT_MAX), C.c_double(float(p["amr_criterion0/value_max"])),
                               O.p(np.zeros(4)), O.p(np.ascontiguousarray(mbs)), O.ip(flag))
        nnew, ndel, _ = mesh.adapt(k, flag)
        assert np.array_equal(mesh.lloc, d[f"lloc_{k + 1}"])            # the same block lists as with one rank
        if nnew or ndel:
            nchanged += 1
        assert int(mesh.nmb_rank.sum()) == mesh.nmb_total and abs(int(mesh.nmb_rank[0]) - int(mesh.nmb_rank[1])) <= 1
        assert mesh.gids_start[0] == 0 and mesh.gids_start[1] == mesh.nmb_rank[0]
        assert np.array_equal(mesh.rank_of, np.repeat([0, 1], mesh.nmb_rank))
        for r in range(2):
            ngh = mesh.rank_tables(r)[2]
            valid = ngh[:, :, 0] >= 0
            assert np.array_equal(ngh[:, :, 2][valid], mesh.rank_of[ngh[:, :, 0][valid]])
    assert nchanged >= 3


def test_location_criterion_through_the_abi_matches_the_oracle():
    """ak_amr_criterion's location branch is host-only (refinement_criteria.cpp:341-386 never touches the device): the C ABI
    itself can be compared with the oracle here, on a three-level refined mesh, for points inside blocks, on block edges and
    outside the domain, in 2-D and 3-D."""
    from athenak_b200 import lib as _lib
    K = _lib.load()
    L = O.lib()
    rng = np.random.default_rng(5)
    for nmesh, nblock in (((64, 64, 1), (8, 8, 1)), ((32, 32, 32), (8, 8, 8))):
        mesh = Mesh(nmesh, nblock, refined_regions=[(1, -0.3, 0.1, -0.2, 0.2, -0.1, 0.3), (2, -0.1, 0.0, -0.1, 0.05, 0.0, 0.1)])
        mbs = np.ascontiguousarray(mesh.rank_tables(0)[0])
        ocfg = O.OCfg(mesh.nmb_total, 2, *nblock, 1, 2, 5.0 / 3.0, O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MAX, 0, 0, 1.0, 0)
        ecfg = _lib.MhdCfg(mesh.nmb_total, 2, *nblock, 1, 2, 5.0 / 3.0, O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MAX, 0, 0, 1.0, 0)
        pts = [rng.uniform(-0.6, 0.6, 3).tolist() + [float(rng.uniform(0.0, 0.3))] for _ in range(40)]
        pts += [[0.0, 0.0, 0.0, 0.125], [-0.25, 0.125, 0.0, 0.0], [0.5, 0.5, 0.5, 0.01]]
        nset = 0
        for pt in pts:
            loc = np.array(pt)
            fo = np.zeros(mesh.nmb_total, np.int32)
            fe = np.zeros(mesh.nmb_total, np.int32)
            fe[::3] = -1                                   # a location criterion overrides an earlier derefinement verdict
            fo[::3] = -1
            L.oracle_amr_criterion(C.byref(ocfg), 3, None, 1, 0, C.c_double(-O.FLT_MAX), C.c_double(O.FLT_MAX), O.p(loc), O.p(mbs), O.ip(fo))
            dp, ipp = C.POINTER(C.c_double), C.POINTER(C.c_int)
            assert K.ak_amr_criterion(C.byref(ecfg), 3, None, 1, 0, -O.FLT_MAX, O.FLT_MAX, loc.ctypes.data_as(dp), mbs.ctypes.data_as(dp),
                                      fe.ctypes.data_as(ipp), None) == 0
            assert np.array_equal(fe, fo)
            nset += int((fo == 1).sum())
        assert nset > 0

