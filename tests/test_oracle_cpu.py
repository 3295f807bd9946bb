"""CPU tests (no GPU): the oracle against the reference's golden vectors and, when the
reference binaries are present (oracle/_ref), against live runs of the unmodified reference."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

import golden_io as G
import oracle_lib as O
from athenak_b200.rst import read_rst

HERE = os.path.dirname(os.path.abspath(__file__))
INTEG = {"rk1": 1, "rk2": 2, "rk3": 3}


def run_oracle(A, nghbr, mbs, integ, ncyc, dt0, cfl=0.3, tlim=0.3):
    st = A.state()
    t, dt = C.c_double(0.0), C.c_double(dt0)
    hist = np.zeros(ncyc + 1)
    n = O.lib().oracle_run(C.byref(A.cfg), C.byref(st), O.p(mbs), O.ip(nghbr), INTEG[integ], C.c_double(cfl),
                           C.c_double(tlim), ncyc, C.byref(t), C.byref(dt), O.p(hist))
    assert n == ncyc
    return t.value, dt.value, hist[ncyc - 1]


@pytest.mark.parametrize("name,recon,rs,integ,cycles", [
    ("blast16_ppm4_hlld_rk3", "ppm4", "hlld", "rk3", (1, 4)),
    ("blast16_plm_hlle_rk2", "plm", "hlle", "rk2", (2,)),
    ("blast16_wenoz_llf_rk3", "wenoz", "llf", "rk3", (2,)),
    ("blast16_ppmx_hlld_rk1", "ppmx", "hlld", "rk1", (2,)),
    ("blast16_teno_hlle_rk2", "teno", "hlle", "rk2", (3,)),
])
def test_oracle_matches_golden(name, recon, rs, integ, cycles):
    """Bitwise (max|delta| = 0) agreement of the C restatement with reference restart dumps."""
    d = G.load(name)
    cfg = O.make_cfg(1, 4, (16, 16, 16), recon, rs)
    A = O.Arrays(cfg)
    G.fill_active(A, d, 0)
    _, ngh, mbs, _ = O.uniform_mesh((16, 16, 16), (16, 16, 16))
    assert G.oracle_init(A, ngh) == [0, 0, 0]
    dt0 = O.lib().oracle_initial_dt(C.byref(cfg), C.byref(A.state()), O.p(mbs), C.c_double(0.3), C.c_double(0.3), C.c_double(0.0))
    assert dt0 == float(d["dt_c0"])
    done, dt = 0, dt0
    for c in cycles:
        t, dt, dt_used = run_oracle(A, ngh, mbs, integ, c - done, dt)
        done = c
        u, b1, b2, b3 = G.active_of(A)
        assert np.abs(u - d[f"u0_c{c}"][0]).max() == 0.0
        assert np.abs(b1 - d[f"b1f_c{c}"][0]).max() == 0.0
        assert np.abs(b2 - d[f"b2f_c{c}"][0]).max() == 0.0
        assert np.abs(b3 - d[f"b3f_c{c}"][0]).max() == 0.0
        # a dump written after cycle c holds the dt that cycle used (driver.cpp:433-450)
        assert dt_used == float(d[f"dt_c{c}"])


def test_known_answer_dt_sequence():
    """SURVEY.md section 8c known-answer: 32^3 blast dt sequence of the reference."""
    if not O.have_reference("athena_blast"):
        pytest.skip("oracle/_ref not built")
    with tempfile.TemporaryDirectory() as tmp:
        out = O.run_reference("athena_blast", os.path.join(HERE, "decks", "blast3d.athinput"), tmp, ["time/nlim=4"])
    dts = [l.split("dt=")[1] for l in out.splitlines() if l.startswith("elapsed=")]
    assert dts[:4] == ["3.895199e-03", "3.910879e-03", "3.987093e-03", "4.116940e-03"]


@pytest.mark.parametrize("nblock", [32, 16])
def test_oracle_matches_live_reference(nblock):
    """32^3 blast, 1 block and 8 blocks: oracle (incl. its multi-block halo exchange and the
    independent uniform-mesh neighbour table) == reference restart dump, bitwise, ghosts included."""
    if not O.have_reference("athena_blast"):
        pytest.skip("oracle/_ref not built")
    ncyc = 2
    with tempfile.TemporaryDirectory() as tmp:
        O.run_reference("athena_blast", os.path.join(HERE, "decks", "blast3d.athinput"), tmp,
                        [f"time/nlim={ncyc}"] + [f"meshblock/nx{d}={nblock}" for d in (1, 2, 3)])
        r0 = read_rst(os.path.join(tmp, "rst", "Blast.00000.rst"))
        r1 = read_rst(os.path.join(tmp, "rst", "Blast.%05d.rst" % ncyc))
    nmb = (32 // nblock) ** 3
    assert r0.nmb == nmb
    lloc, ngh, mbs, root_level = O.uniform_mesh((32, 32, 32), (nblock,) * 3)
    assert root_level == r0.root_level
    assert (lloc == r0.lloc[:, :3]).all() and (r0.lloc[:, 3] == root_level).all()   # Z-order gids
    cfg = O.make_cfg(nmb, 4, (nblock,) * 3)
    A = O.Arrays(cfg)
    A.u0[:], A.b0[0][:], A.b0[1][:], A.b0[2][:] = r0.u0, r0.b1f, r0.b2f, r0.b3f
    A.c2p_all()
    run_oracle(A, ngh, mbs, "rk3", ncyc, r0.dt)
    assert np.abs(A.u0 - r1.u0).max() == 0.0
    for a, b in zip(A.b0, (r1.b1f, r1.b2f, r1.b3f)):
        assert np.abs(a - b).max() == 0.0


@pytest.mark.parametrize("nblock,bcs", [
    (32, ("reflect",) * 6),
    (16, ("outflow", "outflow", "reflect", "reflect", "periodic", "periodic")),
    (16, ("diode", "reflect", "outflow", "diode", "reflect", "outflow")),
    (16, ("reflect", "outflow", "periodic", "periodic", "diode", "diode")),
])
def test_oracle_physical_bcs_match_live_reference(nblock, bcs):
    """Task 26 ApplyPhysicalBCs: a blast wide enough to cross the mesh edges with non-periodic
    flags, reference restart dump == oracle bitwise, ghosts included."""
    if not O.have_reference("athena_blast"):
        pytest.skip("oracle/_ref not built")
    ncyc = 3
    names = ("ix1_bc", "ox1_bc", "ix2_bc", "ox2_bc", "ix3_bc", "ox3_bc")
    with tempfile.TemporaryDirectory() as tmp:
        O.run_reference("athena_blast", os.path.join(HERE, "decks", "blast3d.athinput"), tmp,
                        [f"time/nlim={ncyc}", "problem/inner_radius=0.42", "problem/outer_radius=0.55"]
                        + [f"mesh/{k}={v}" for k, v in zip(names, bcs)]
                        + [f"meshblock/nx{d}={nblock}" for d in (1, 2, 3)])
        r0 = read_rst(os.path.join(tmp, "rst", "Blast.00000.rst"))
        r1 = read_rst(os.path.join(tmp, "rst", "Blast.%05d.rst" % ncyc))
    nmb = (32 // nblock) ** 3
    per = tuple(bcs[2 * d] == "periodic" for d in range(3))
    lloc, ngh, mbs, _ = O.uniform_mesh((32, 32, 32), (nblock,) * 3, periodic=per)
    mesh_bcs = np.array([O.BC[b] for b in bcs], np.int32)
    mb_bcs = O.block_bcs(lloc, (32,) * 3, (nblock,) * 3, mesh_bcs)
    cfg = O.make_cfg(nmb, 4, (nblock,) * 3)
    A = O.Arrays(cfg)
    A.u0[:], A.b0[0][:], A.b0[1][:], A.b0[2][:] = r0.u0, r0.b1f, r0.b2f, r0.b3f
    A.c2p_all()
    O.lib().oracle_set_bcs(O.ip(mesh_bcs), O.ip(mb_bcs))
    try:
        run_oracle(A, ngh, mbs, "rk3", ncyc, r0.dt)
    finally:
        O.lib().oracle_set_bcs(None, None)
    assert np.abs(A.u0 - r1.u0).max() == 0.0
    for a, b in zip(A.b0, (r1.b1f, r1.b2f, r1.b3f)):
        assert np.abs(a - b).max() == 0.0


GOLD_BCS = ("reflect", "outflow", "diode", "reflect", "outflow", "diode")


def test_oracle_physical_bcs_match_golden():
    """Committed fixture of the reference run with mixed physical BCs (ghost-inclusive final state)."""
    d = G.load("blast16_bcs_mixed")
    cfg = O.make_cfg(1, 4, (16, 16, 16))
    A = O.Arrays(cfg)
    G.fill_active(A, d, 0)
    lloc, ngh, mbs, _ = O.uniform_mesh((16, 16, 16), (16, 16, 16), periodic=False)
    mesh_bcs = np.array([O.BC[b] for b in GOLD_BCS], np.int32)
    mb_bcs = O.block_bcs(lloc, (16,) * 3, (16,) * 3, mesh_bcs)
    # Driver::Initialize: no neighbours at all -> ghosts come from the physical BCs alone
    assert O.lib().oracle_apply_bcs(C.byref(cfg), O.ip(mesh_bcs), O.ip(mb_bcs), O.p(A.u0), *[O.p(a) for a in A.b0]) == 0
    A.c2p_all()
    dt0 = O.lib().oracle_initial_dt(C.byref(cfg), C.byref(A.state()), O.p(mbs), C.c_double(0.3), C.c_double(0.3), C.c_double(0.0))
    assert dt0 == float(d["dt_c0"])
    O.lib().oracle_set_bcs(O.ip(mesh_bcs), O.ip(mb_bcs))
    try:
        run_oracle(A, ngh, mbs, "rk3", 3, dt0)
    finally:
        O.lib().oracle_set_bcs(None, None)
    assert np.array_equal(A.u0, d["u0g_c3"])
    for a, k in zip(A.b0, ("b1fg_c3", "b2fg_c3", "b3fg_c3")):
        assert np.array_equal(a, d[k])
