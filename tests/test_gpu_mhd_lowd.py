"""1-D and 2-D MHD on the GPU (SURVEY.md 8a rows a11/a13 "1-D/2-D branches", 8c): Brio & Wu shock
tube in 1-D (single and multi-block), the same tube along x2 on a 2-D multi-block mesh with
reflect / outflow / diode boundaries, and the 2-D Orszag-Tang vortex.  Synthetic inputs = the
reference's own initial condition (cycle-0 restart dump of oracle/_ref/athena_builtin, run live);
gate: final u0 and face fields (ghost zones included) == the reference's final dump, bitwise,
through the per-task ABI, the fused ABI and the C++ host driver; and == oracle/mhd_oracle.c."""
import os
import tempfile

import numpy as np
import pytest

import hydro_util as H
import oracle_lib as O
from athenak_b200.rst import read_rst

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
DECKS = os.path.join(HERE, "decks")
_CACHE = {}

CASES = [
    ("bw1d.athinput", "BrioWu", ()),
    ("bw1d.athinput", "BrioWu", ("mhd/rsolver=hlld", "mhd/reconstruct=ppm4", "mesh/nghost=3", "time/integrator=rk3", "mesh/nx1=128", "meshblock/nx1=32")),
    ("bw1d.athinput", "BrioWu", ("mhd/rsolver=hlle", "mhd/reconstruct=wenoz", "mesh/nghost=3", "time/integrator=rk3", "time/cfl_number=0.3")),
    ("bw2d.athinput", "BrioWu", ()),
    ("bw2d.athinput", "BrioWu", ("mhd/rsolver=llf", "mhd/reconstruct=plm", "mesh/nghost=2", "time/integrator=rk2")),
    ("ot2d.athinput", "OrszagTang", ()),
    ("bw1d.athinput", "BrioWu", ("mhd/nscalars=1",)),                                 # passive scalars
    ("bw2d.athinput", "BrioWu", ("mhd/nscalars=2",)),
]
IDS = [f"{c[0].split('.')[0]}-{'-'.join(o.split('=')[1] for o in c[2]) or 'deck'}" for c in CASES]


def reference(deck, base, overrides=()):
    if not O.have_reference("athena_builtin"):
        pytest.skip("oracle/_ref/athena_builtin not built")
    key = (deck, base, tuple(overrides))
    if key not in _CACHE:
        with tempfile.TemporaryDirectory() as tmp:
            O.run_reference("athena_builtin", os.path.join(DECKS, deck), tmp, list(overrides), threads=4)
            ns = max([int(o.split("=")[1]) for o in overrides if o.startswith("mhd/nscalars=")] + [0])
            _CACHE[key] = (read_rst(os.path.join(tmp, "rst", base + ".00000.rst"), nmhd=5 + ns),
                           read_rst(os.path.join(tmp, "rst", base + ".00001.rst"), nmhd=5 + ns))
    return _CACHE[key]


def check(u, b, r1):
    assert np.array_equal(u, r1.u0), f"max|dU| = {np.abs(u - r1.u0).max()}"
    for x, y, nm in zip(b, (r1.b1f, r1.b2f, r1.b3f), ("b1f", "b2f", "b3f")):
        assert np.array_equal(x, y), f"max|d{nm}| = {np.abs(x - y).max()}"


@pytest.mark.parametrize("deck,base,ov", CASES, ids=IDS)
@pytest.mark.parametrize("fused", [True, False], ids=["fused", "pertask"])
def test_mhd_1d_2d_matches_reference_bitwise(deck, base, ov, fused):
    from athenak_b200.driver import Driver, assemble_mhd_tasks
    from athenak_b200.mesh import Mesh
    from athenak_b200.mhd import MHD
    r0, r1 = reference(deck, base, ov)
    case = H.Case(r0.deck)
    mesh = Mesh(case.nmesh, case.nblock, xmin=case.xmin, xmax=case.xmax, bcs=case.bc_names)
    mb_size, mb_bcs, nghbr = mesh.rank_tables(0)
    assert np.array_equal(nghbr[:, :, 0], case.nghbr)
    e = MHD(case.nmb, case.ng, case.nblock, mb_size, nghbr, recon=case.recon, rsolver=case.rsolver, gamma=case.gamma,
            mesh_bcs=mesh.bcs, mb_bcs=mb_bcs, nscalars=case.nscalars)
    e.load_state(r0.u0, r0.b1f, r0.b2f, r0.b3f)
    drv = Driver(case.integrator, case.cfl, tlim=case.tlim)
    tl = assemble_mhd_tasks(e, fused=fused)
    drv.Initialize(e)
    assert drv.dt == r0.dt
    drv.Execute(e, tl)
    assert drv.ncycle == r1.ncycle and drv.time == r1.time
    check(e.u0.cpu().numpy(), [x.cpu().numpy() for x in e.b0], r1)


@pytest.mark.parametrize("deck,base,ov", [CASES[1], CASES[3], CASES[5], CASES[7]], ids=[IDS[1], IDS[3], IDS[5], IDS[7]])
def test_mhd_1d_2d_cxx_host_driver(deck, base, ov):
    from athenak_b200.mesh import Mesh
    from athenak_b200.sim import Simulation
    r0, r1 = reference(deck, base, ov)
    case = H.Case(r0.deck)
    mesh = Mesh(case.nmesh, case.nblock, xmin=case.xmin, xmax=case.xmax, bcs=case.bc_names)
    sim = Simulation(mesh, 0, ng=case.ng, recon=case.recon, rsolver=case.rsolver, gamma=case.gamma,
                     integrator=case.integrator, cfl_number=case.cfl, tlim=case.tlim, fused=True, nscalars=case.nscalars)
    sim.set_state(*[np.ascontiguousarray(a) for a in (r0.u0, r0.b1f, r0.b2f, r0.b3f)])
    sim.initialize()
    assert sim.status()["dt"] == r0.dt
    sim.execute(-1)
    st = sim.status()
    assert st["ncycle"] == r1.ncycle and st["time"] == r1.time
    u, b1, b2, b3 = sim.get_state()
    check(u, (b1, b2, b3), r1)


def test_mhd_2d_matches_oracle_after_one_cycle():
    """Device state == oracle/mhd_oracle.c after one cycle (u0, w0, bcc0, faces; ghosts included)."""
    from athenak_b200.driver import Driver, assemble_mhd_tasks
    from athenak_b200.mesh import Mesh
    from athenak_b200.mhd import MHD
    r0, _ = reference("bw2d.athinput", "BrioWu")
    case = H.Case(r0.deck)
    mesh = Mesh(case.nmesh, case.nblock, xmin=case.xmin, xmax=case.xmax, bcs=case.bc_names)
    mb_size, mb_bcs, nghbr = mesh.rank_tables(0)
    e = MHD(case.nmb, case.ng, case.nblock, mb_size, nghbr, recon=case.recon, rsolver=case.rsolver, gamma=case.gamma,
            mesh_bcs=mesh.bcs, mb_bcs=mb_bcs)
    e.load_state(r0.u0, r0.b1f, r0.b2f, r0.b3f)
    drv = Driver(case.integrator, case.cfl, tlim=case.tlim)
    drv.Initialize(e)
    drv.Execute(e, assemble_mhd_tasks(e, fused=True), ncycles=1)
    A, t, dt, n, dt0 = H.oracle_run_mhd(case, r0.u0, (r0.b1f, r0.b2f, r0.b3f), ncycles=1)
    assert n == 1 and drv.time == t and drv.dt == dt
    assert np.array_equal(e.u0.cpu().numpy(), A.u0) and np.array_equal(e.w0.cpu().numpy(), A.w0)
    assert np.array_equal(e.bcc0.cpu().numpy(), A.bcc0)
    for x, y in zip(e.b0, A.b0):
        assert np.array_equal(x.cpu().numpy(), y)
