"""Hydro twin of the stage update on the GPU (BASELINE config 1 = Sod shock tube,
config 4's HLLC leg = 3-D hydro linear wave).  Synthetic inputs = the reference's own initial condition
(cycle-0 restart dump of oracle/_ref/athena_builtin, run live); gates:
  * final u0 (ghost zones included) == the reference's final dump, bitwise, for 1-D / 2-D / 3-D meshes,
    single and multi-block, LLF / HLLE / HLLC / Roe, PLM / PPM4 / PPMX / WENOZ / TENO, RK1-3,
    outflow / reflect / diode / periodic boundaries -- through the per-task ABI and the fused one;
  * the same state == oracle/hydro_oracle.c;
  * Sod: 175 cycles and L1(rho) against the analytic solution == the reference's known answer
    (3.549279e-03) and below the reference test's threshold;
  * linear wave: RMS-L1 == the value the reference writes to LinWave-errs.dat, and converging.
"""
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


def reference(deck, base, overrides=(), want_errs=False):
    if not O.have_reference("athena_builtin"):
        pytest.skip("oracle/_ref/athena_builtin not built")
    key = (deck, base, tuple(overrides), want_errs)
    if key not in _CACHE:
        _CACHE[key] = _reference(deck, base, overrides, want_errs)
    return _CACHE[key]


def _reference(deck, base, overrides, want_errs):
    with tempfile.TemporaryDirectory() as tmp:
        O.run_reference("athena_builtin", os.path.join(DECKS, deck), tmp, list(overrides), threads=4)
        ns = max([int(o.split("=")[1]) for o in overrides if o.startswith("hydro/nscalars=")] + [0])
        r0 = read_rst(os.path.join(tmp, "rst", base + ".00000.rst"), physics="hydro", nmhd=5 + ns)
        r1 = read_rst(os.path.join(tmp, "rst", base + ".00001.rst"), physics="hydro", nmhd=5 + ns)
        err = None
        if want_errs:
            err = float(open(os.path.join(tmp, base + "-errs.dat")).read().split("\n")[1].split()[4])
    return r0, r1, err


def run_engine(case: H.Case, u_init, fused, tlim=None, ncycles=None):
    from athenak_b200.driver import Driver
    from athenak_b200.hydro import Hydro, assemble_hydro_tasks
    from athenak_b200.mesh import Mesh
    mesh = Mesh(case.nmesh, case.nblock, xmin=case.xmin, xmax=case.xmax, bcs=case.bc_names)
    assert (mesh.lloc[:, :3] == case.lloc).all()
    mb_size, mb_bcs, nghbr = mesh.rank_tables(0)
    assert np.array_equal(mb_size, case.mb_size) and np.array_equal(mb_bcs, case.mb_bcs)
    assert np.array_equal(nghbr[:, :, 0], case.nghbr)
    hyd = Hydro(case.nmb, case.ng, case.nblock, mb_size, nghbr, recon=case.recon, rsolver=case.rsolver,
                gamma=case.gamma, mesh_bcs=mesh.bcs, mb_bcs=mb_bcs, nscalars=case.nscalars)
    hyd.load_state(u_init)
    drv = Driver(case.integrator, case.cfl, tlim=case.tlim if tlim is None else tlim)
    tl = assemble_hydro_tasks(hyd, fused=fused)
    drv.Initialize(hyd)
    dt0 = drv.dt
    drv.Execute(hyd, tl, ncycles=ncycles)
    return hyd.u0.cpu().numpy(), hyd.w0.cpu().numpy(), drv, dt0, hyd


CASES = [
    ("sod1d.athinput", "Sod", ()),                                                   # BASELINE config 1 as is
    ("sod1d.athinput", "Sod", ("hydro/rsolver=hllc", "hydro/reconstruct=ppm4", "mesh/nghost=3", "time/integrator=rk3", "time/cfl_number=0.4")),
    ("sod1d.athinput", "Sod", ("hydro/rsolver=roe", "hydro/reconstruct=wenoz", "mesh/nghost=3", "time/integrator=rk3", "time/cfl_number=0.4")),
    ("sod1d.athinput", "Sod", ("hydro/rsolver=hlle", "hydro/reconstruct=ppmx", "mesh/nghost=4", "time/integrator=rk1", "time/cfl_number=0.3")),
    ("sod1d.athinput", "Sod", ("hydro/rsolver=hllc", "hydro/reconstruct=dc", "mesh/nx1=64", "meshblock/nx1=16")),   # 4 blocks in 1-D
    ("sod2d.athinput", "Sod", ()),                                                   # 2-D, 8 blocks, reflect/outflow/diode
    ("sod2d.athinput", "Sod", ("hydro/rsolver=roe", "hydro/reconstruct=teno")),
    ("sod2d.athinput", "Sod", ("hydro/rsolver=llf", "hydro/reconstruct=plm", "mesh/nghost=2", "time/integrator=rk2")),
    ("lwave3d_hydro.athinput", "LinWave", ("hydro/rsolver=hlle", "hydro/reconstruct=wenoz")),
    ("sod1d.athinput", "Sod", ("hydro/nscalars=2",)),                                # passive scalars
    ("sod2d.athinput", "Sod", ("hydro/nscalars=1",)),
    ("sod2d.athinput", "Sod", ("hydro/nscalars=3", "hydro/rsolver=llf", "hydro/reconstruct=wenoz")),
]


@pytest.mark.parametrize("deck,base,ov", CASES, ids=[f"{c[0].split('.')[0]}-{'-'.join(o.split('=')[1] for o in c[2]) or 'deck'}" for c in CASES])
@pytest.mark.parametrize("fused", [True, False], ids=["fused", "pertask"])
def test_hydro_matches_reference_bitwise(deck, base, ov, fused):
    r0, r1, _ = reference(deck, base, ov)
    case = H.Case(r0.deck)
    u, w, drv, dt0, hyd = run_engine(case, r0.u0, fused)
    assert dt0 == r0.dt
    assert drv.ncycle == r1.ncycle and drv.time == r1.time
    assert np.array_equal(u, r1.u0), f"max|dU| = {np.abs(u - r1.u0).max()}"
    assert hyd.launches > 0


def test_hydro_matches_oracle_every_cycle():
    """Device state == oracle/hydro_oracle.c after 1, 2, ... cycles (u0 and w0, ghosts included)."""
    r0, _, _ = reference("sod2d.athinput", "Sod")
    case = H.Case(r0.deck)
    for nc in (1, 3):
        u, w, drv, dt0, _ = run_engine(case, r0.u0, True, ncycles=nc)
        uo, wo, t, dtn, n, dt0o = H.oracle_run(case, r0.u0, ncycles=nc)
        assert n == nc and dt0 == dt0o and drv.time == t and drv.dt == dtn
        assert np.array_equal(u, uo) and np.array_equal(w, wo)


def test_sod_l1_error_gate():
    """BASELINE config 1: Sod 256, PLM+LLF+RK2, cfl 0.8 -- 175 cycles, L1(rho) = 3.549279e-03."""
    r0, r1, _ = reference("sod1d.athinput", "Sod")
    case = H.Case(r0.deck)
    u, w, drv, _, _ = run_engine(case, r0.u0, True)
    assert drv.ncycle == 175
    ng = case.ng
    x = case.xmin[0] + (np.arange(256) + 0.5) / 256.0 * (case.xmax[0] - case.xmin[0])
    l1 = float(np.abs(w[0, 0, 0, 0, ng:-ng] - H.sod_density(x, 0.25)).mean())
    assert abs(l1 - 3.549279e-03) < 1e-7, l1     # the known answer was computed from a %12.5e tab dump
    assert l1 < 0.01       # tst/test_suite/nr/test_nr_sod_cpu.py threshold for 256 zones


def rms_l1(u, u_ic, case):
    s = [slice(case.ng, -case.ng) if n > 1 else slice(None) for n in case.nblock]
    errs = [np.abs(u[:, n, s[2], s[1], s[0]] - u_ic[:, n, s[2], s[1], s[0]]).mean() for n in range(5)]
    return float(np.sqrt(sum(e * e for e in errs)))


@pytest.mark.parametrize("recon,rs", [("ppm4", "hllc"), ("plm", "hllc")])
def test_hydro_linear_wave_convergence(recon, rs):
    """BASELINE config 4, HLLC leg: 3-D hydro sound wave, one period, 32x16x16 -> 64x32x32."""
    errs = []
    for n in (32, 64):
        ov = [f"mesh/nx1={n}", f"mesh/nx2={n // 2}", f"mesh/nx3={n // 2}", f"meshblock/nx1={n // 2}",
              f"meshblock/nx2={n // 8}", f"meshblock/nx3={n // 8}", f"hydro/reconstruct={recon}", f"hydro/rsolver={rs}"]
        r0, r1, ref_err = reference("lwave3d_hydro.athinput", "LinWave", ov, want_errs=True)
        case = H.Case(r0.deck)
        u, w, drv, dt0, _ = run_engine(case, r0.u0, True)
        assert drv.ncycle == r1.ncycle and np.array_equal(u, r1.u0)
        e = rms_l1(u, r0.u0, case)
        assert abs(e - ref_err) <= 2e-6 * ref_err, (e, ref_err)
        errs.append(e)
    assert errs[1] / errs[0] < (0.4 if recon == "plm" else 0.25)


@pytest.mark.parametrize("deck,base,ov", [CASES[0], CASES[5], CASES[8], CASES[11]], ids=["sod1d", "sod2d", "lwave3d", "sod2d-3scalars"])
@pytest.mark.parametrize("fused", [True, False], ids=["fused", "pertask"])
def test_hydro_cxx_host_driver(deck, base, ov, fused):
    """The same cases through the C++ host driver (TaskList/Driver/Hydro tasks in csrc/host_driver.cpp, ak_sim_*)."""
    from athenak_b200.mesh import Mesh
    from athenak_b200.sim import Simulation
    r0, r1, _ = reference(deck, base, ov)
    case = H.Case(r0.deck)
    mesh = Mesh(case.nmesh, case.nblock, xmin=case.xmin, xmax=case.xmax, bcs=case.bc_names)
    sim = Simulation(mesh, 0, ng=case.ng, recon=case.recon, rsolver=case.rsolver, gamma=case.gamma,
                     integrator=case.integrator, cfl_number=case.cfl, tlim=case.tlim, fused=fused, physics="hydro",
                     nscalars=case.nscalars)
    sim.set_state(np.ascontiguousarray(r0.u0))
    sim.initialize()
    assert sim.status()["dt"] == r0.dt
    sim.execute(-1)
    st = sim.status()
    assert st["ncycle"] == r1.ncycle and st["time"] == r1.time
    (u,) = sim.get_state()
    assert np.array_equal(u, r1.u0), f"max|dU| = {np.abs(u - r1.u0).max()}"


def _read_hst(path):
    rows = [ln.split() for ln in open(path) if not ln.startswith("#")]
    return np.array(rows, dtype=np.float64)


def test_history_sums_hydro_and_mhd():
    """the hst sums as a device reduction == the reference's .hst rows (written
    with %24.16e) to reassociation accuracy, for the hydro twin (2-D Sod, 8 blocks) and for MHD
    (3-D linear wave, 32 blocks)."""
    if not O.have_reference("athena_builtin"):
        pytest.skip("oracle/_ref/athena_builtin not built")
    # hydro
    with tempfile.TemporaryDirectory() as tmp:
        O.run_reference("athena_builtin", os.path.join(DECKS, "sod2d.athinput"), tmp, threads=4)
        r0 = read_rst(os.path.join(tmp, "rst", "Sod.00000.rst"), physics="hydro")
        hst = _read_hst(os.path.join(tmp, "Sod.hydro.hst"))
    case = H.Case(r0.deck)
    u, w, drv, dt0, hyd = run_engine(case, r0.u0, True, ncycles=0)
    s0 = hyd.HistorySums()
    assert s0.shape == (8,)
    assert np.allclose(s0, hst[0, 2:10], rtol=1e-13, atol=1e-15), (s0, hst[0, 2:10])
    u, w, drv, dt0, hyd = run_engine(case, r0.u0, True)
    s1 = hyd.HistorySums()
    assert drv.time == hst[-1, 0]
    assert np.allclose(s1, hst[-1, 2:10], rtol=1e-12, atol=1e-14), (s1, hst[-1, 2:10])
    # MHD
    from athenak_b200.driver import Driver
    from athenak_b200.mesh import Mesh
    from athenak_b200.mhd import MHD
    with tempfile.TemporaryDirectory() as tmp:
        O.run_reference("athena_builtin", os.path.join(DECKS, "lwave3d.athinput"), tmp, ["time/tlim=0.0"], threads=4)
        r0 = read_rst(os.path.join(tmp, "rst", "LinWave.00000.rst"))
        hst = _read_hst(os.path.join(tmp, "LinWave.mhd.hst"))
    mesh = Mesh((32, 16, 16), (16, 4, 4), xmin=(0.0, 0.0, 0.0), xmax=(3.0, 1.5, 1.5))
    mb_size, _, nghbr = mesh.rank_tables(0)
    e = MHD(mesh.nmb_total, 4, (16, 4, 4), mb_size, nghbr, recon="ppm4", rsolver="hlld", gamma=1.66666666667)
    e.load_state(r0.u0, r0.b1f, r0.b2f, r0.b3f)
    sm = e.HistorySums()
    assert sm.shape == (11,)
    assert np.allclose(sm, hst[0, 2:13], rtol=1e-13, atol=1e-15), (sm, hst[0, 2:13])
