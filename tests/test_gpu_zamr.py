"""Adaptive mesh refinement on the GPU (dynamic refinement): refinement criteria (ak_amr_criterion),
mesh bookkeeping (ak_mesh_adapt), data movement as gathers from the old arrays into the new ones (ak_amr_plan / ak_amr_move),
RepairAMRFC (ak_amr_repair_fc) and the AdaptiveMeshRefinement sequence (athenak_b200/amr.py) around the multilevel MHD task
list.  Gate: from the unmodified reference's state of cycle k the engine reproduces the reference's state of cycle k+1
BITWISE -- U and B with ghost zones, block list, time, dt -- across one refinement and one derefinement event (committed
fixture; passed on a B200 for the fused and the per-task list), and across whole live runs when oracle/_ref/athena_blast is
present -- U, B (every face, far-field values at the denormal floor included), block list, time and dt, all bit for bit.
History: in round 1 the 2-D three-level live run differed after 23 cycles in four faces of x2f at the 5e-175 level (the PLM
limiter's guard-free division with a numerator at the denormal floor) and B was compared with |dB| <= 1e-100; the
range guard now lives inside every division / reciprocal / square root (csrc/mhd_device.cuh) and the tolerance is 0.
The AMR loop exists twice: athenak_b200/amr.py around the per-task ABI, and inside the C++ host driver (ak_sim_execute_amr).
(File name: sorted last on purpose -- the newest, least-measured tests should not stop an `-x` run before the others.)"""
import numpy as np
import pytest

import amr_util as AU
import golden_io as G
import hydro_util as H
import oracle_lib as O

pytestmark = pytest.mark.gpu


class EngineStepper:
    def __init__(self, p):
        self.p, self.mesh = p, AU.make_mesh(p)
        self.nadapt = self.ncreated = self.ndeleted = self.nfofc = 0

    def make_mhd(self, mesh):
        from athenak_b200.mhd import MHD
        p = self.p
        mb_size, mb_bcs, nghbr = mesh.rank_tables(0)
        return MHD(mesh.nmb_total, int(p["mesh/nghost"]), mesh.nblock, mb_size, nghbr, recon=p["mhd/reconstruct"], rsolver=p["mhd/rsolver"],
                   gamma=float(p["mhd/gamma"]), mb_lev=mesh.lloc[:, 3], mesh_bcs=mesh.bcs, mb_bcs=mb_bcs,
                   fofc=p.get("mhd/fofc", "false") == "true")

    def advance(self, r, ncycle, fused=True):
        import torch
        from athenak_b200 import amr
        from athenak_b200.driver import Driver, assemble_mhd_tasks
        p, mesh = self.p, self.mesh
        assert np.array_equal(mesh.lloc, r.lloc)
        mhd = self.make_mhd(mesh)
        mhd.load_state(r.u0, r.b1f, r.b2f, r.b3f)
        drv = Driver(p["time/integrator"], float(p["time/cfl_number"]), tlim=float(p["time/tlim"]))
        drv.time, drv.dt, drv.ncycle = r.time, r.dt, ncycle
        mhd.ConToPrim(drv, 0)
        mhd.NewTimeStep(drv, drv.nexp_stages)        # dtnew as the last stage of the cycle left it
        nmb0 = mhd.nmb
        mhd, changed = amr.adaptive_mesh_refinement(mhd, mesh, drv, amr.criteria_from_params(p), self.make_mhd)
        if changed:
            self.nadapt += 1
            self.ncreated += max(0, mhd.nmb - nmb0)
            self.ndeleted += max(0, nmb0 - mhd.nmb)
        drv.NewTimeStep(mhd)                          # Mesh::NewTimeStep after AMR (driver.cpp:449-450)
        dt = drv.dt
        drv.Execute(mhd, assemble_mhd_tasks(mhd, fused=fused), ncycles=1)
        torch.cuda.synchronize()
        self.nfofc += mhd.nfofc
        return mhd, drv.time, dt


def check_step(S, r, rn, k, fused=True, b_atol=0.0):
    mhd, t, dt = S.advance(r, k, fused)
    assert np.array_equal(S.mesh.lloc, rn.lloc), f"block list after cycle {k}"
    assert t == rn.time and dt == rn.dt, f"time / dt after cycle {k}: {t} {rn.time} {dt} {rn.dt}"
    u = mhd.u0.cpu().numpy()
    assert np.array_equal(u, rn.u0), f"cycle {k}: max|dU| = {np.abs(u - rn.u0).max()}"
    for v, (x, y) in enumerate(zip(mhd.b0, (rn.b1f, rn.b2f, rn.b3f))):
        x = x.cpu().numpy()
        if not np.array_equal(x, y):
            bad = np.argwhere(np.abs(x - y) > b_atol)
            w = np.argwhere(x != y)[0]
            assert len(bad) == 0, (f"cycle {k}: x{v + 1}f differs in {np.count_nonzero(x != y)} faces, {len(bad)} beyond {b_atol}; first "
                                   f"{tuple(w)}: {x[tuple(w)]!r} vs {y[tuple(w)]!r}")
    assert mhd.launches > 0


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "pertask"])
def test_amr_events_match_fixture_bitwise(fused):
    d = G.load("amr_blast2d_events")
    p = H.deck_params(str(d["deck"]))
    S = EngineStepper(p)

    class R:
        pass

    def state(k):
        r = R()
        r.lloc, r.u0, r.b1f, r.b2f, r.b3f = d[f"lloc_{k}"], d[f"u0_{k}"], d[f"b1f_{k}"], d[f"b2f_{k}"], d[f"b3f_{k}"]
        r.time, r.dt = float(d[f"time_{k}"]), float(d[f"dt_{k}"])
        return r

    events = [int(k) for k in d["events"]]
    L = O.lib()
    for k in range(1, int(d["ncycles"]) - 1):
        if k in events:
            check_step(S, state(k), state(k + 1), k, fused)
        else:       # replay the mesh bookkeeping only (criteria verdict from the oracle on the stored density)
            cfg = AU.ocfg(S.mesh, p)
            mbs = np.ascontiguousarray(S.mesh.rank_tables(0)[0])
            S.mesh.adapt(k, AU.criteria_flags(L, cfg, p, np.ascontiguousarray(d[f"dens_{k}"][:, None]), mbs))
        assert np.array_equal(S.mesh.lloc, d[f"lloc_{k + 1}"])
    assert S.ncreated > 0 and S.ndeleted > 0


def test_amr_execute_loop_across_both_events():
    """athenak_b200.amr.execute (cycle -> AdaptiveMeshRefinement -> new dt, task lists rebuilt on a mesh change) from the
    fixture's state of the refinement cycle to the state after the derefinement cycle, without reloading in between."""
    import torch
    from athenak_b200 import amr
    from athenak_b200.driver import Driver, assemble_mhd_tasks
    d = G.load("amr_blast2d_events")
    p = H.deck_params(str(d["deck"]))
    events = [int(k) for k in d["events"]]
    k0, k1 = events[0], events[-1] + 1
    assert all(f"u0_{k}" in d for k in (k0, k1))
    S = EngineStepper(p)
    L = O.lib()
    for k in range(1, k0):          # mesh bookkeeping up to the first event
        cfg = AU.ocfg(S.mesh, p)
        mbs = np.ascontiguousarray(S.mesh.rank_tables(0)[0])
        S.mesh.adapt(k, AU.criteria_flags(L, cfg, p, np.ascontiguousarray(d[f"dens_{k}"][:, None]), mbs))
    mhd = S.make_mhd(S.mesh)
    mhd.load_state(d[f"u0_{k0}"], d[f"b1f_{k0}"], d[f"b2f_{k0}"], d[f"b3f_{k0}"])
    drv = Driver(p["time/integrator"], float(p["time/cfl_number"]), tlim=float(p["time/tlim"]))
    drv.time, drv.dt, drv.ncycle = float(d[f"time_{k0}"]), float(d[f"dt_{k0}"]), k0
    mhd.ConToPrim(drv, 0)
    mhd.NewTimeStep(drv, drv.nexp_stages)
    crit = amr.criteria_from_params(p)
    mhd, changed = amr.adaptive_mesh_refinement(mhd, S.mesh, drv, crit, S.make_mhd)     # the AMR step of cycle k0
    assert changed
    drv.NewTimeStep(mhd)
    mhd, nchanged = amr.execute(mhd, S.mesh, drv, crit, S.make_mhd, ncycles=k1 - 1 - k0)    # cycles k0+1 .. k1-1, each followed by AMR
    assert nchanged >= 1 and drv.ncycle == k1 - 1
    drv.Execute(mhd, assemble_mhd_tasks(mhd), ncycles=1)                                  # cycle k1 (its AMR step comes after the dump)
    torch.cuda.synchronize()
    assert drv.ncycle == k1 and drv.time == float(d[f"time_{k1}"])
    assert np.array_equal(S.mesh.lloc, d[f"lloc_{k1}"])
    assert np.array_equal(mhd.u0.cpu().numpy(), d[f"u0_{k1}"])
    for x, key in zip(mhd.b0, ("b1f", "b2f", "b3f")):
        assert np.array_equal(x.cpu().numpy(), d[f"{key}_{k1}"])


def test_amr_loop_of_the_cxx_driver_across_both_events():
    """The same stretch through the C++ host driver's own AMR loop (ak_sim_adapt / ak_sim_execute_amr, csrc/host_driver.cpp):
    state of the refinement cycle loaded restart-style, AMR step, then cycle -> AdaptiveMeshRefinement -> new dt up to the
    cycle after the derefinement; block list, time, U and B must equal the reference's dump bit for bit."""
    from athenak_b200 import amr
    from athenak_b200.sim import Simulation
    d = G.load("amr_blast2d_events")
    p = H.deck_params(str(d["deck"]))
    events = [int(k) for k in d["events"]]
    k0, k1 = events[0], events[-1] + 1
    S = EngineStepper(p)
    L = O.lib()
    for k in range(1, k0):          # mesh bookkeeping up to the first event
        cfg = AU.ocfg(S.mesh, p)
        mbs = np.ascontiguousarray(S.mesh.rank_tables(0)[0])
        S.mesh.adapt(k, AU.criteria_flags(L, cfg, p, np.ascontiguousarray(d[f"dens_{k}"][:, None]), mbs))
    mesh = S.mesh
    sim = Simulation(mesh, 0, ng=int(p["mesh/nghost"]), recon=p["mhd/reconstruct"], rsolver=p["mhd/rsolver"], gamma=float(p["mhd/gamma"]),
                     integrator=p["time/integrator"], cfl_number=float(p["time/cfl_number"]), tlim=float(p["time/tlim"]), fused=True,
                     multilevel=True)
    sim.set_state(*[np.ascontiguousarray(d[f"{q}_{k0}"]) for q in ("u0", "b1f", "b2f", "b3f")])
    sim.restart_from(float(d[f"time_{k0}"]), float(d[f"dt_{k0}"]), k0)
    crit = amr.criteria_from_params(p)
    changed, st = sim.adapt(crit)                       # the AMR step of cycle k0 (+ Mesh::NewTimeStep)
    assert changed and st["nadapt"] == 1
    n, st = sim.execute_amr(crit, ncycles=k1 - 1 - k0)  # cycles k0+1 .. k1-1, each followed by AMR
    assert n == k1 - 1 - k0 and st["nadapt"] >= 2 and st["ndeleted"] > 0
    assert sim.execute(1) == 1                          # cycle k1 (its AMR step comes after the dump)
    stt = sim.status()
    assert stt["ncycle"] == k1 and stt["time"] == float(d[f"time_{k1}"])
    assert np.array_equal(mesh.lloc, d[f"lloc_{k1}"])
    u, x1, x2, x3 = sim.get_state()
    assert np.array_equal(u, d[f"u0_{k1}"])
    for x, key in zip((x1, x2, x3), ("b1f", "b2f", "b3f")):
        assert np.array_equal(x, d[f"{key}_{k1}"])


@pytest.mark.parametrize("name", ["2d_3levels", "3d", "3d_deref"])
def test_amr_cycle_by_cycle_matches_live_reference_bitwise(name):
    if not O.have_reference("athena_blast"):
        pytest.skip("oracle/_ref/athena_blast not built")
    from test_amr_data_cpu import CASES
    ov = CASES[name]
    dumps = AU.reference_dumps(ov)
    S = EngineStepper(AU.params(ov))
    for k in range(1, len(dumps) - 1):
        check_step(S, dumps[k], dumps[k + 1], k, b_atol=0.0)
    assert S.nadapt >= 1 and S.ncreated > 0


def test_fofc_on_a_refined_mesh_matches_live_reference():
    """FOFC inside the multilevel stage (per-task order: CopyCons, Fluxes + FOFC, flux correction, ...) on a strong magnetised
    blast where the correction fires, with an AMR event on the way; same tolerance rule as the other live runs."""
    if not O.have_reference("athena_blast"):
        pytest.skip("oracle/_ref/athena_blast not built")
    from test_amr_data_cpu import _FOFC
    ov = _FOFC + ("mhd/reconstruct=ppm4", "time/integrator=rk3")
    dumps = AU.reference_dumps(ov)
    S = EngineStepper(AU.params(ov))
    for k in range(1, len(dumps) - 1):
        check_step(S, dumps[k], dumps[k + 1], k, b_atol=0.0)
    assert S.nfofc > 100 and S.nadapt >= 1


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "pertask"])
def test_mhd_smr_with_passive_scalar_and_bcs_matches_live_reference(fused):
    """Static refinement + physical boundaries + one passive scalar (nvar = 6) through the engine's multilevel MHD list:
    Brio-Wu along x2, 49 cycles, bitwise with ghost zones."""
    import smr_util as SU
    import torch
    from athenak_b200.driver import Driver, assemble_mhd_tasks
    from athenak_b200.mhd import MHD
    if not O.have_reference("athena_builtin"):
        pytest.skip("oracle/_ref/athena_builtin not built")
    r0, r1 = SU.reference_mhd_deck(SU.SCALAR_DECK, 6)
    mesh, p = SU.mesh_of(r0.deck)
    mb_size, mb_bcs, nghbr = mesh.rank_tables(0)
    mhd = MHD(mesh.nmb_total, int(p["mesh/nghost"]), mesh.nblock, mb_size, nghbr, recon=p["mhd/reconstruct"], rsolver=p["mhd/rsolver"],
              gamma=float(p["mhd/gamma"]), mb_lev=mesh.lloc[:, 3], mesh_bcs=mesh.bcs, mb_bcs=mb_bcs, nscalars=1)
    mhd.load_state(r0.u0, r0.b1f, r0.b2f, r0.b3f)
    drv = Driver(p["time/integrator"], float(p["time/cfl_number"]), tlim=float(p["time/tlim"]))
    drv.Initialize(mhd)
    assert drv.dt == r0.dt
    drv.Execute(mhd, assemble_mhd_tasks(mhd, fused=fused))
    torch.cuda.synchronize()
    assert drv.ncycle == r1.ncycle and drv.time == r1.time and drv.dt == r1.dt
    u = mhd.u0.cpu().numpy()
    assert np.array_equal(u, r1.u0), f"max|dU| = {np.abs(u - r1.u0).max()}"
    for x, y in zip(mhd.b0, (r1.b1f, r1.b2f, r1.b3f)):
        assert np.array_equal(x.cpu().numpy(), y)


def test_hydro_amr_cycle_by_cycle_matches_live_reference_bitwise():
    """The hydro twin through the same AMR sequence (cell-centred gather only, one ghost-zone rebuild, no RepairAMRFC):
    PLM+HLLC+RK2 blast, 39 cycles with refinement and derefinement events."""
    import torch
    from athenak_b200 import amr
    from athenak_b200.driver import Driver
    from athenak_b200.hydro import Hydro, assemble_hydro_tasks
    if not O.have_reference("athena_blast"):
        pytest.skip("oracle/_ref/athena_blast not built")
    p = AU.params(deck=AU.HYDRO_DECK)
    dumps = AU.reference_dumps(deck=AU.HYDRO_DECK, physics="hydro")
    mesh = AU.make_mesh(p)

    def make(mesh):
        mb_size, mb_bcs, nghbr = mesh.rank_tables(0)
        return Hydro(mesh.nmb_total, int(p["mesh/nghost"]), mesh.nblock, mb_size, nghbr, recon=p["hydro/reconstruct"], rsolver=p["hydro/rsolver"],
                     gamma=float(p["hydro/gamma"]), mb_lev=mesh.lloc[:, 3], mesh_bcs=mesh.bcs, mb_bcs=mb_bcs)

    crit = amr.criteria_from_params(p)
    nchanged = 0
    for k in range(1, len(dumps) - 1):
        r, rn = dumps[k], dumps[k + 1]
        assert np.array_equal(mesh.lloc, r.lloc)
        hyd = make(mesh)
        hyd.load_state(r.u0)
        drv = Driver(p["time/integrator"], float(p["time/cfl_number"]), tlim=float(p["time/tlim"]))
        drv.time, drv.dt, drv.ncycle = r.time, r.dt, k
        hyd.ConToPrim(drv, 0)
        hyd.NewTimeStep(drv, drv.nexp_stages)
        hyd, changed = amr.adaptive_mesh_refinement(hyd, mesh, drv, crit, make)
        nchanged += int(changed)
        drv.NewTimeStep(hyd)
        dt = drv.dt
        drv.Execute(hyd, assemble_hydro_tasks(hyd), ncycles=1)
        torch.cuda.synchronize()
        assert np.array_equal(mesh.lloc, rn.lloc) and drv.time == rn.time and dt == rn.dt, f"cycle {k}"
        u = hyd.u0.cpu().numpy()
        assert np.array_equal(u, rn.u0), f"cycle {k}: max|dU| = {np.abs(u - rn.u0).max()}"
    assert nchanged >= 2

