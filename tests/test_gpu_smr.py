"""Hydro on statically refined meshes on the GPU (BASELINE config 5's fine/coarse data path
for cell-centred variables): restriction, ghost exchange across refinement levels, coarse-buffer fill +
slope-limited prolongation and flux correction as descriptor-driven kernels (csrc/smr.cu, ak_smr_*).
Gate: u0 INCLUDING the prolongated ghost zones == the unmodified reference's dump, bitwise -- committed fixtures
always, live reference runs (3 levels in 3-D, WENOZ+Roe, level-3 2-D, 1-D) when oracle/_ref is present -- and
== oracle/smr_oracle.c after every cycle."""
import numpy as np
import pytest

import golden_io as G
import hydro_util as H
import oracle_lib as O
import smr_util as S

pytestmark = pytest.mark.gpu
_CACHE = {}


def run_engine(deck_text, u_init, fused, ncycles=None):
    from athenak_b200.driver import Driver
    from athenak_b200.hydro import Hydro, assemble_hydro_tasks
    mesh, p = S.mesh_of(deck_text)
    mb_size, mb_bcs, nghbr = mesh.rank_tables(0)
    hyd = Hydro(mesh.nmb_total, int(p["mesh/nghost"]), mesh.nblock, mb_size, nghbr, recon=p["hydro/reconstruct"],
                rsolver=p["hydro/rsolver"], gamma=float(p["hydro/gamma"]), mb_lev=mesh.lloc[:, 3], mesh_bcs=mesh.bcs, mb_bcs=mb_bcs)
    hyd.load_state(u_init)
    drv = Driver(p["time/integrator"], float(p["time/cfl_number"]), tlim=float(p["time/tlim"]))
    tl = assemble_hydro_tasks(hyd, fused=fused)
    drv.Initialize(hyd)
    dt0 = drv.dt
    drv.Execute(hyd, tl, ncycles=ncycles)
    return hyd, drv, dt0


@pytest.mark.parametrize("name", sorted(S.CASES))
@pytest.mark.parametrize("fused", [True, False], ids=["fused_c2p", "pertask"])
def test_smr_matches_golden_bitwise(name, fused):
    d = G.load(name)
    u_init, u_final, act = S.with_ghosts(d)
    hyd, drv, dt0 = run_engine(str(d["deck"]), u_init, fused, ncycles=0)
    u = hyd.u0.cpu().numpy()
    assert dt0 == float(d["dt_c0"]) and np.array_equal(u[:, :, act[0], act[1], act[2]], d["u0_c0"])    # Driver::Initialize
    hyd, drv, dt0 = run_engine(str(d["deck"]), u_init, fused)
    assert drv.ncycle == int(d["ncycle_final"]) and drv.time == float(d["time_final"]) and drv.dt == float(d["dt_final"])
    u = hyd.u0.cpu().numpy()[:, :, act[0], act[1], act[2]]
    assert np.array_equal(u, u_final), f"max|dU| = {np.abs(u - u_final).max()}"
    assert hyd.launches > 0


@pytest.mark.parametrize("name", sorted(set(S.LIVE_CASES) - set(S.CASES)))
def test_smr_matches_live_reference_bitwise(name):
    if not O.have_reference("athena_builtin"):
        pytest.skip("oracle/_ref/athena_builtin not built")
    if name not in _CACHE:
        _CACHE[name] = S.reference(name)
    r0, r1 = _CACHE[name]
    hyd, drv, dt0 = run_engine(r0.deck, r0.u0, True)
    assert dt0 == r0.dt and drv.ncycle == r1.ncycle and drv.time == r1.time
    u = hyd.u0.cpu().numpy()
    assert np.array_equal(u, r1.u0), f"max|dU| = {np.abs(u - r1.u0).max()}"        # ghost zones included


def test_smr_matches_oracle_every_cycle():
    d = G.load("hydro_smr2d_plm_llf_rk2")
    for nc in (1, 2, 5):
        hyd, drv, dt0 = run_engine(str(d["deck"]), d["u0_c0"], True, ncycles=nc)
        uo, wo, t, dtn, n, dt0o = S.oracle_run(str(d["deck"]), d["u0_c0"], ncycles=nc)
        assert n == nc and dt0 == dt0o and drv.time == t and drv.dt == dtn
        assert np.array_equal(hyd.u0.cpu().numpy(), uo) and np.array_equal(hyd.w0.cpu().numpy(), wo)


@pytest.mark.parametrize("name", sorted(S.CASES))
@pytest.mark.parametrize("fused", [True, False], ids=["fused_c2p", "pertask"])
def test_smr_cxx_host_driver(name, fused):
    """The same fixtures through the C++ host driver (csrc/host_driver.cpp: ak_sim_set_levels, multilevel task list)."""
    from athenak_b200.sim import Simulation
    d = G.load(name)
    u_init, u_final, act = S.with_ghosts(d)
    mesh, p = S.mesh_of(str(d["deck"]))
    sim = Simulation(mesh, 0, ng=int(p["mesh/nghost"]), recon=p["hydro/reconstruct"], rsolver=p["hydro/rsolver"],
                     gamma=float(p["hydro/gamma"]), integrator=p["time/integrator"], cfl_number=float(p["time/cfl_number"]),
                     tlim=float(p["time/tlim"]), fused=fused, physics="hydro", multilevel=True)
    sim.set_state(np.ascontiguousarray(u_init))
    sim.initialize()
    assert sim.status()["dt"] == float(d["dt_c0"])
    sim.execute(-1)
    st = sim.status()
    assert st["ncycle"] == int(d["ncycle_final"]) and st["time"] == float(d["time_final"]) and st["dt"] == float(d["dt_final"])
    (u,) = sim.get_state()
    u = u[:, :, act[0], act[1], act[2]]
    assert np.array_equal(u, u_final), f"max|dU| = {np.abs(u - u_final).max()}"
