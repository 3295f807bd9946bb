"""Isothermal EOS and first-order flux correction (FOFC) on the GPU: isothermal LLF / HLLE / HLLD (MHD) and LLF / HLLE / Roe
(hydro) through the same sweep kernels with four fluid variables, density-floor-only ConsToPrim and the
isothermal fast / sound speed in NewTimeStep; 1-D / 2-D / 3-D, multi-block, physical boundaries, passive
scalars.  Gate: final u0 and face fields (ghost zones included) == the unmodified reference's final
dump, bitwise -- against the committed fixtures always, and against live reference runs of more
combinations when oracle/_ref is present -- through the per-task ABI, the fused ABI and the C++ host
driver; history sums have no tot-E column.  FOFC cases (mhd_fofc.cpp) are ones in which the correction fires:
the number of flagged cells must equal the count in the reference's event log."""
import numpy as np
import pytest

import golden_io as G
import hydro_util as H
import iso_util as I
import oracle_lib as O

pytestmark = pytest.mark.gpu
_CACHE = {}


class _Dump:
    pass


def initial_and_final(name, live):
    """(case, u0, (b1f,b2f,b3f) or None, final dump) from the committed fixture or a live reference run."""
    if live:
        if not O.have_reference("athena_builtin"):
            pytest.skip("oracle/_ref/athena_builtin not built")
        if name not in _CACHE:
            _CACHE[name] = I.reference(name)
        r0, r1 = _CACHE[name]
        case = H.Case(r0.deck)
        return case, r0.u0, (None if r0.b1f is None else (r0.b1f, r0.b2f, r0.b3f)), r1
    d = G.load(name)
    case = H.Case(str(d["deck"]))
    r1 = _Dump()
    r1.u0, r1.ncycle, r1.time, r1.nfofc = d["u0_final"], int(d["ncycle_final"]), float(d["time_final"]), int(d["nfofc"])
    b = None
    if case.physics == "mhd":
        b = (d["b1f_c0"], d["b2f_c0"], d["b3f_c0"])
        r1.b1f, r1.b2f, r1.b3f = d["b1f_final"], d["b2f_final"], d["b3f_final"]
    return case, d["u0_c0"], b, r1


def run_python_driver(case, u_init, b_init, fused, ncycles=None):
    from athenak_b200.driver import Driver, assemble_mhd_tasks
    from athenak_b200.hydro import Hydro, assemble_hydro_tasks
    from athenak_b200.mesh import Mesh
    from athenak_b200.mhd import MHD
    mesh = Mesh(case.nmesh, case.nblock, xmin=case.xmin, xmax=case.xmax, bcs=case.bc_names)
    mb_size, mb_bcs, nghbr = mesh.rank_tables(0)
    kw = dict(recon=case.recon, rsolver=case.rsolver, gamma=case.gamma, mesh_bcs=mesh.bcs, mb_bcs=mb_bcs,
              nscalars=case.nscalars, eos="isothermal" if case.iso else "ideal", iso_sound_speed=case.iso_cs, dfloor=case.dfloor)
    if case.physics == "hydro":
        e = Hydro(case.nmb, case.ng, case.nblock, mb_size, nghbr, **kw)
        e.load_state(u_init)
        tl = assemble_hydro_tasks(e, fused=fused)
    else:
        e = MHD(case.nmb, case.ng, case.nblock, mb_size, nghbr, fofc=case.fofc, **kw)
        e.load_state(u_init, *b_init)
        tl = assemble_mhd_tasks(e, fused=fused)
    drv = Driver(case.integrator, case.cfl, tlim=case.tlim)
    drv.Initialize(e)
    drv.Execute(e, tl, ncycles=ncycles)
    return e, drv


def check(u, b, r1):
    assert np.array_equal(u, r1.u0), f"max|dU| = {np.abs(u - r1.u0).max()}"
    if b is not None:
        for x, y, nm in zip(b, (r1.b1f, r1.b2f, r1.b3f), ("b1f", "b2f", "b3f")):
            assert np.array_equal(x, y), f"max|d{nm}| = {np.abs(x - y).max()}"


@pytest.mark.parametrize("name", sorted(I.CASES))
@pytest.mark.parametrize("fused", [True, False], ids=["fused", "pertask"])
def test_iso_matches_golden_bitwise(name, fused):
    case, u0, b0, r1 = initial_and_final(name, live=False)
    e, drv = run_python_driver(case, u0, b0, fused, ncycles=r1.ncycle)
    assert drv.ncycle == r1.ncycle and drv.time == r1.time
    check(e.u0.cpu().numpy(), None if b0 is None else [x.cpu().numpy() for x in e.b0], r1)
    assert e.launches > 0 and e.u0.shape[1] == case.nfl + case.nscalars
    if case.fofc and not fused:       # the per-task list reads the device counter back every stage
        assert e.nfofc == r1.nfofc > 0


@pytest.mark.parametrize("name", sorted(set(I.LIVE_CASES) - set(I.CASES)))
def test_iso_matches_live_reference_bitwise(name):
    case, u0, b0, r1 = initial_and_final(name, live=True)
    fused = not case.fofc
    e, drv = run_python_driver(case, u0, b0, fused, ncycles=r1.ncycle)
    assert drv.ncycle == r1.ncycle and drv.time == r1.time
    check(e.u0.cpu().numpy(), None if b0 is None else [x.cpu().numpy() for x in e.b0], r1)
    if case.fofc:
        assert e.nfofc == r1.nfofc


@pytest.mark.parametrize("name", sorted(I.CASES))
def test_iso_cxx_host_driver(name):
    """The same fixtures through the C++ host driver (csrc/host_driver.cpp, ak_sim_*)."""
    from athenak_b200.mesh import Mesh
    from athenak_b200.sim import Simulation
    case, u0, b0, r1 = initial_and_final(name, live=False)
    mesh = Mesh(case.nmesh, case.nblock, xmin=case.xmin, xmax=case.xmax, bcs=case.bc_names)
    sim = Simulation(mesh, 0, ng=case.ng, recon=case.recon, rsolver=case.rsolver, gamma=case.gamma,
                     integrator=case.integrator, cfl_number=case.cfl, tlim=case.tlim, fused=True, physics=case.physics,
                     nscalars=case.nscalars, eos="isothermal" if case.iso else "ideal", iso_sound_speed=case.iso_cs,
                     dfloor=case.dfloor, fofc=case.fofc)
    sim.set_state(*[np.ascontiguousarray(a) for a in ((u0,) if b0 is None else (u0, *b0))])
    sim.initialize()
    sim.execute(r1.ncycle)
    st = sim.status()
    assert st["ncycle"] == r1.ncycle and st["time"] == r1.time
    out = sim.get_state()
    check(out[0], None if b0 is None else out[1:], r1)


def test_iso_matches_oracle_after_one_cycle():
    """Device u0, w0, bcc0 and faces == oracle/mhd_oracle.c after one cycle (ghosts included)."""
    from athenak_b200.driver import Driver, assemble_mhd_tasks
    case, u0, b0, _ = initial_and_final("mhd_iso_bw2d_wenoz_hlle_rk2", live=False)
    from athenak_b200.mesh import Mesh
    from athenak_b200.mhd import MHD
    mesh = Mesh(case.nmesh, case.nblock, xmin=case.xmin, xmax=case.xmax, bcs=case.bc_names)
    mb_size, mb_bcs, nghbr = mesh.rank_tables(0)
    e = MHD(case.nmb, case.ng, case.nblock, mb_size, nghbr, recon=case.recon, rsolver=case.rsolver, gamma=case.gamma,
            mesh_bcs=mesh.bcs, mb_bcs=mb_bcs, eos="isothermal", iso_sound_speed=case.iso_cs)
    e.load_state(u0, *b0)
    drv = Driver(case.integrator, case.cfl, tlim=case.tlim)
    drv.Initialize(e)
    drv.Execute(e, assemble_mhd_tasks(e, fused=True), ncycles=1)
    A, t, dt, n, dt0 = H.oracle_run_mhd(case, u0, b0, ncycles=1)
    assert n == 1 and drv.time == t and drv.dt == dt
    assert np.array_equal(e.u0.cpu().numpy(), A.u0) and np.array_equal(e.w0.cpu().numpy(), A.w0)
    assert np.array_equal(e.bcc0.cpu().numpy(), A.bcc0)
    for x, y in zip(e.b0, A.b0):
        assert np.array_equal(x.cpu().numpy(), y)


def test_iso_history_sums():
    """history.cpp:84-88,278-282: 7 (hydro) / 10 (MHD) columns without tot-E."""
    case, u0, b0, _ = initial_and_final("mhd_iso_bw2d_wenoz_hlle_rk2", live=False)
    from athenak_b200.mesh import Mesh
    from athenak_b200.mhd import MHD
    mesh = Mesh(case.nmesh, case.nblock, xmin=case.xmin, xmax=case.xmax, bcs=case.bc_names)
    mb_size, mb_bcs, nghbr = mesh.rank_tables(0)
    e = MHD(case.nmb, case.ng, case.nblock, mb_size, nghbr, recon=case.recon, rsolver=case.rsolver, gamma=case.gamma,
            mesh_bcs=mesh.bcs, mb_bcs=mb_bcs, eos="isothermal", iso_sound_speed=case.iso_cs)
    e.load_state(u0, *b0)
    s = e.HistorySums()
    assert s.shape == (10,)
    ng = case.ng
    vol = case.mb_size[:, 6] * case.mb_size[:, 7] * case.mb_size[:, 8]
    a = u0[:, :, :, ng:-ng, ng:-ng]
    want = [(vol[:, None, None, None] * a[:, n]).sum() for n in range(4)]
    want += [(vol[:, None, None, None] * 0.5 * a[:, n] ** 2 / a[:, 0]).sum() for n in (1, 2, 3)]
    assert np.allclose(s[:7], want, rtol=1e-12, atol=1e-14)
    b1, b2, b3 = b0
    me1 = (vol[:, None, None, None] * 0.25 * (b1[:, :, ng:-ng, ng + 1:-ng] ** 2 + b1[:, :, ng:-ng, ng:-ng - 1] ** 2)).sum()
    assert np.allclose(s[7], me1, rtol=1e-12, atol=1e-14)
