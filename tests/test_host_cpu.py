"""CPU tests of the host layer and the C ABI surface (no GPU, no compute calls)."""
import ctypes as C
import os
import re
import tempfile

import numpy as np
import pytest

import halo_emul as H
import oracle_lib as O
from athenak_b200 import lib as _lib
from athenak_b200.mesh import Mesh
from athenak_b200.rst import read_rst

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_library_exports_every_declared_symbol(engine_lib):
    hdr = open(os.path.join(ROOT, "include", "athenak_b200.h")).read()
    declared = set(re.findall(r"\b(ak_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(engine_lib, name), f"{name} declared in include/athenak_b200.h but not exported"
    assert declared == set(_lib.EXPORTED)


def test_create_fails_loudly_without_gpu(engine_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = _lib.MhdCfg(1, 4, 16, 16, 16, 2, 2, 1.6666667, 1e-38, 1e-38, 1e-38, 1e-38, 1e38)
    h = C.c_void_p()
    assert engine_lib.ak_mhd_create(C.byref(cfg), C.byref(h)) != 0
    assert b"no CUDA device" in engine_lib.ak_last_error(None)
    from athenak_b200.mhd import MHD
    with pytest.raises(_lib.EngineError):
        MHD(1, 4, (16, 16, 16), np.zeros((1, 9)), -np.ones((1, 56, 4), np.int32), gamma=1.6666667)


def test_cfg_validation(engine_lib):
    h = C.c_void_p()
    bad = [_lib.MhdCfg(1, 2, 16, 16, 16, 2, 2, 1.6666667, 0, 0, 0, 0, 1e38),    # ppm4 needs ng>=3
           _lib.MhdCfg(1, 4, 16, 1, 16, 2, 2, 1.6666667, 0, 0, 0, 0, 1e38),     # nx2 = 1 with nx3 > 1
           _lib.MhdCfg(1, 4, 16, 16, 16, 9, 2, 1.6666667, 0, 0, 0, 0, 1e38),    # unknown recon
           _lib.MhdCfg(0, 4, 16, 16, 16, 2, 2, 1.6666667, 0, 0, 0, 0, 1e38)]    # no blocks
    for cfg in bad:
        assert engine_lib.ak_mhd_create(C.byref(cfg), C.byref(h)) != 0
        assert len(engine_lib.ak_last_error(None)) > 0


def test_neighbor_index_table(engine_lib):
    """Slot numbering 0-55 (src/mesh/nghbr_index.hpp:28-54): faces, edges, corners."""
    assert [engine_lib.ak_neighbor_index(*a, 0, 0) for a in [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0)]] == [0, 4, 8, 12]
    assert [engine_lib.ak_neighbor_index(*a, 0, 0) for a in [(-1, -1, 0), (1, -1, 0), (-1, 1, 0), (1, 1, 0)]] == [16, 18, 20, 22]
    assert [engine_lib.ak_neighbor_index(*a, 0, 0) for a in [(0, 0, -1), (0, 0, 1)]] == [24, 28]
    assert [engine_lib.ak_neighbor_index(*a, 0, 0) for a in [(-1, 0, -1), (1, 0, -1), (-1, 0, 1), (1, 0, 1)]] == [32, 34, 36, 38]
    assert [engine_lib.ak_neighbor_index(*a, 0, 0) for a in [(0, -1, -1), (0, 1, -1), (0, -1, 1), (0, 1, 1)]] == [40, 42, 44, 46]
    assert [engine_lib.ak_neighbor_index(ox, oy, oz, 0, 0) for oz in (-1, 1) for oy in (-1, 1) for ox in (-1, 1)] == list(range(48, 56))
    assert engine_lib.ak_neighbor_index(0, 0, 0, 0, 0) == -1
    for oz in (-1, 0, 1):
        for oy in (-1, 0, 1):
            for ox in (-1, 0, 1):
                assert engine_lib.ak_neighbor_index(ox, oy, oz, 0, 0) == O.lib().oracle_neighbor_index(ox, oy, oz, 0, 0)


@pytest.mark.parametrize("nmesh,nblock,periodic", [
    ((32, 32, 32), (32, 32, 32), True), ((32, 32, 32), (16, 16, 16), True), ((32, 16, 16), (16, 4, 4), True),
    ((48, 32, 16), (16, 16, 16), True), ((64, 64, 64), (16, 16, 16), False), ((40, 24, 8), (8, 8, 8), True)])
def test_host_mesh_matches_definition(nmesh, nblock, periodic):
    """Z-order gids, block extents and the neighbour table against the independent Python
    restatement in tests/oracle_lib.py (incomplete root octs, non-periodic edges included)."""
    bcs = ("periodic",) * 6 if periodic else ("outflow",) * 6
    M = Mesh(nmesh, nblock, bcs=bcs, nranks=1)
    lloc, ngh, mbs, rl = O.uniform_mesh(nmesh, nblock, periodic=periodic)
    assert M.root_level == rl and M.nmb_total == len(lloc)
    assert (M.lloc[:, :3] == lloc).all() and (M.lloc[:, 3] == rl).all()
    mb_size, mb_bcs, nghbr = M.rank_tables(0)
    assert np.abs(mb_size - mbs).max() == 0.0
    assert np.array_equal(mb_bcs, O.block_bcs(lloc, nmesh, nblock, [O.BC[b] for b in bcs]))   # meshblock.cpp:25-131
    assert (nghbr[:, :, 0] == ngh).all()
    has = ngh >= 0
    assert (nghbr[:, :, 1][has] == rl).all() and (nghbr[:, :, 2][has] == 0).all()
    # dest slot = slot of the opposite offset
    L = _lib.load()
    for oz in (-1, 0, 1):
        for oy in (-1, 0, 1):
            for ox in (-1, 0, 1):
                if ox == oy == oz == 0:
                    continue
                s, dslot = L.ak_neighbor_index(ox, oy, oz, 0, 0), L.ak_neighbor_index(-ox, -oy, -oz, 0, 0)
                col = nghbr[:, s, :]
                assert (col[col[:, 0] >= 0, 3] == dslot).all()


@pytest.mark.parametrize("nblock", [(16, 16, 16), (32, 16, 8)])
def test_zorder_and_sizes_match_live_reference(nblock):
    """lloc_eachmb / mesh header written by the unmodified reference into its restart dump."""
    if not O.have_reference("athena_blast"):
        pytest.skip("oracle/_ref not built")
    with tempfile.TemporaryDirectory() as tmp:
        O.run_reference("athena_blast", os.path.join(HERE, "decks", "blast3d.athinput"), tmp,
                        ["time/nlim=0"] + [f"meshblock/nx{d+1}={nblock[d]}" for d in range(3)])
        r = read_rst(os.path.join(tmp, "rst", "Blast.00000.rst"))
    M = Mesh((32, 32, 32), nblock)
    assert M.nmb_total == r.nmb and M.root_level == r.root_level
    assert (M.lloc == r.lloc).all()


@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 8])
def test_load_balance_contiguous_ranges(nranks):
    """LoadBalance (load_balance.cpp:38-88): contiguous gid ranges filled from the end, so
    the LOWER ranks get the smaller share; 64 blocks on 8 ranks = one octant each."""
    M = Mesh((64, 64, 64), (16, 16, 16), nranks=nranks)
    assert M.nmb_rank.sum() == 64 and (np.diff(M.gids_start) == M.nmb_rank[:-1]).all()
    assert (np.diff(M.rank_of) >= 0).all()
    if 64 % nranks == 0:
        assert (M.nmb_rank == 64 // nranks).all()
    else:
        assert M.nmb_rank[0] <= M.nmb_rank[-1]
    if nranks == 8:
        for r in range(8):
            l = M.lloc[M.gids_start[r]:M.gids_start[r] + 8, :3]
            assert (l.max(0) - l.min(0) == 1).all()     # a 2x2x2 octant


def _exchange_multi_rank(nmesh, nblock, nranks, rng):
    """Runs the per-rank plans against each other in-process (messages = numpy buffers) and
    returns (engine arrays, oracle arrays) after one CC and one FC exchange."""
    M = Mesh(nmesh, nblock, nranks=nranks)
    lloc, ngh, mbs, _ = O.uniform_mesh(nmesh, nblock)
    ng = 4
    cfg_all = O.make_cfg(M.nmb_total, ng, nblock)
    A = O.Arrays(cfg_all)
    A.u0[:] = rng.standard_normal(A.u0.shape)
    for a in A.b0:
        a[:] = rng.standard_normal(a.shape)
    u = A.u0.copy()
    b = [a.copy() for a in A.b0]
    L = O.lib()
    L.oracle_halo_cc(C.byref(cfg_all), O.p(A.u0), 5, O.ip(ngh))
    L.oracle_halo_fc(C.byref(cfg_all), *[O.p(a) for a in A.b0], O.ip(ngh))
    plans, views = [], []
    for r in range(nranks):
        g0, n = int(M.gids_start[r]), int(M.nmb_rank[r])
        cfg = _lib.MhdCfg(n, ng, *nblock, 2, 2, 1.6666667, 0, 0, 0, 0, 1e38)
        _, _, nghbr = M.rank_tables(r)
        plans.append(H.make_plan(cfg, nghbr, r, g0))
        views.append(({-1: u[g0:g0 + n]}, {q: b[q][g0:g0 + n] for q in range(3)}))
    # pack everything first (sends complete before any unpack), then local gather + unpack
    msgs = {}
    for r in range(nranks):
        for ip, (peer, ncc, nfc) in enumerate(H.peers(plans[r])):
            bc, bf = np.zeros(5 * ncc), np.zeros(nfc)
            H.pack(H.descs(plans[r], 2, ip), views[r][0], 5, bc)
            H.pack(H.descs(plans[r], 4, ip), views[r][1], 1, bf)
            msgs[(r, peer)] = (bc, bf)
    for r in range(nranks):
        H.local(H.descs(plans[r], 0), views[r][0])
        H.local(H.descs(plans[r], 1), views[r][1])
        for ip, (peer, ncc, nfc) in enumerate(H.peers(plans[r])):
            bc, bf = msgs[(peer, r)]
            assert bc.size == 5 * ncc and bf.size == nfc
            H.unpack(H.descs(plans[r], 3, ip), views[r][0], 5, bc)
            H.unpack(H.descs(plans[r], 5, ip), views[r][1], 1, bf)
    for p in plans:
        _lib.load().ak_halo_destroy(p)
    return (u, b), (A.u0, A.b0)


@pytest.mark.parametrize("nmesh,nblock,nranks", [((16, 16, 16), (16, 16, 16), 1), ((16, 16, 16), (8, 8, 8), 1),
                                                 ((16, 16, 16), (8, 8, 8), 2), ((32, 16, 16), (8, 8, 8), 3),
                                                 ((16, 16, 16), (8, 8, 8), 8), ((32, 16, 16), (16, 16, 16), 2),
                                                 ((32, 32, 1), (16, 16, 1), 1), ((32, 32, 1), (8, 8, 1), 3),      # 2-D
                                                 ((64, 1, 1), (64, 1, 1), 1), ((64, 1, 1), (16, 1, 1), 2)])       # 1-D
def test_halo_plan_equals_reference_exchange(nmesh, nblock, nranks):
    """The resolved copy plan (same-rank gather + per-peer messages) reproduces the reference's
    pack/unpack exchange bit for bit for every rank count, incl. a block that is its own
    neighbour and two ranks that are each other's neighbours on both sides."""
    (u, b), (ou, ob) = _exchange_multi_rank(nmesh, nblock, nranks, np.random.default_rng(11))
    assert np.array_equal(u, ou)
    for x, y in zip(b, ob):
        assert np.array_equal(x, y)


def test_pgens_match_live_reference():
    """ak_pgen_blast / ak_pgen_orszag_tang produce the reference's initial conditions bit for bit
    (active zones of every block; 8 blocks)."""
    if not (O.have_reference("athena_blast") and O.have_reference("athena_builtin")):
        pytest.skip("oracle/_ref not built")
    for binary, deck, gamma, gen in (("athena_blast", "blast3d.athinput", 1.6666667, "pgen_blast"),
                                     ("athena_builtin", "ot3d.athinput", 1.666666667, "pgen_orszag_tang")):
        with tempfile.TemporaryDirectory() as tmp:
            O.run_reference(binary, os.path.join(HERE, "decks", deck), tmp,
                            ["time/nlim=0"] + [f"meshblock/nx{d}=16" for d in (1, 2, 3)])
            base = "Blast" if "blast" in deck else "OrszagTang"
            r = read_rst(os.path.join(tmp, "rst", base + ".00000.rst"))
        M = Mesh((32, 32, 32), (16, 16, 16))
        u0, b1, b2, b3 = getattr(M, gen)(0, 4, gamma)
        s = slice(4, -4)
        assert np.array_equal(u0.numpy()[:, :, s, s, s], r.u0[:, :, s, s, s]), gen
        assert np.array_equal(b1.numpy()[:, s, s, 4:-4], r.b1f[:, s, s, 4:-4])
        assert np.array_equal(b2.numpy()[:, s, 4:-4, s], r.b2f[:, s, 4:-4, s])
        assert np.array_equal(b3.numpy()[:, 4:-4, s, s], r.b3f[:, 4:-4, s, s])


def test_rst_roundtrip_and_reference_restart(tmp_path):
    """write_rst(read_rst(x)) is byte-identical to the reference's dump, and the reference
    restarts from a dump written by this repo and continues to the same bits."""
    if not O.have_reference("athena_blast"):
        pytest.skip("oracle/_ref not built")
    from athenak_b200.rst import write_rst
    deck = os.path.join(HERE, "decks", "blast3d.athinput")
    ov = [f"meshblock/nx{d}=16" for d in (1, 2, 3)] + [f"mesh/nx{d}=16" for d in (1, 2, 3)]
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    O.run_reference("athena_blast", deck, a, ov + ["time/nlim=4"])
    src = os.path.join(a, "rst", "Blast.00002.rst")
    r = read_rst(src)
    mine = str(tmp_path / "mine.rst")
    write_rst(mine, r)
    assert open(mine, "rb").read() == open(src, "rb").read()
    # restart the reference from OUR file and run to cycle 4
    import subprocess
    exe = os.path.join(O.REF_DIR, "athena_blast")
    os.makedirs(b, exist_ok=True)
    p = subprocess.run([exe, "-r", mine, "time/nlim=4"], cwd=b, capture_output=True, text=True,
                       env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
    cont = read_rst(os.path.join(b, "rst", "Blast.00004.rst"))
    full = read_rst(os.path.join(a, "rst", "Blast.00004.rst"))
    assert cont.ncycle == full.ncycle == 4 and cont.time == full.time
    assert np.array_equal(cont.u0, full.u0) and np.array_equal(cont.b1f, full.b1f)
