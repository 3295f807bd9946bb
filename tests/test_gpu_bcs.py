"""GPU parity tests for task 26 ApplyPhysicalBCs (src/mhd/mhd_tasks.cpp:504-520,
src/bvals/physics/hydro_bcs.cpp, bfield_bcs.cpp) through the C ABI (ak_bcs_*), against the oracle
(pinned bitwise to live reference runs with the same flags, tests/test_oracle_cpu.py) and against
the committed reference fixture.  Gate: max|delta| == 0 (pure copies / sign flips)."""
import ctypes as C

import numpy as np
import pytest

import golden_io as G
import oracle_lib as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

FLAGSETS = [
    ("reflect",) * 6,
    ("outflow",) * 6,
    ("diode", "reflect", "outflow", "diode", "reflect", "outflow"),
    ("vacuum", "diode", "periodic", "periodic", "reflect", "vacuum"),
    ("periodic", "periodic", "periodic", "periodic", "outflow", "reflect"),
]
GOLD_BCS = ("reflect", "outflow", "diode", "reflect", "outflow", "diode")


def _tables(nmesh, nblock, bcs):
    per = tuple(bcs[2 * d] == "periodic" for d in range(3))
    lloc, ngh, mbs, _ = O.uniform_mesh(nmesh, nblock, periodic=per)
    mesh_bcs = np.array([O.BC[b] for b in bcs], np.int32)
    return lloc, ngh, mbs, mesh_bcs, O.block_bcs(lloc, nmesh, nblock, mesh_bcs)


@pytest.mark.parametrize("bcs", FLAGSETS)
@pytest.mark.parametrize("nblock", [(16, 16, 16), (8, 8, 8), (16, 8, 4)])
@pytest.mark.parametrize("split", [False, True])
def test_bc_fill_bitwise(bcs, nblock, split):
    """Random data everywhere (ghosts included): the three ordered passes fill faces, edges and
    corners exactly as the oracle does; HydroBCs/BFieldBCs separately and as one fused call."""
    from athenak_b200 import lib as _lib
    L = _lib.load()
    lloc, ngh, mbs, mesh_bcs, mb_bcs = _tables((16, 16, 16), nblock, bcs)
    cfg = O.make_cfg(len(lloc), 3, nblock)
    A = O.Arrays(cfg)
    rng = np.random.default_rng(7)
    for a in (A.u0, *A.b0):
        a[:] = rng.standard_normal(a.shape)
    dev = [torch.from_numpy(a.copy()).cuda() for a in (A.u0, *A.b0)]
    assert O.lib().oracle_apply_bcs(C.byref(cfg), O.ip(mesh_bcs), O.ip(mb_bcs), O.p(A.u0), *[O.p(a) for a in A.b0]) == 0
    mcfg = _lib.MhdCfg(cfg.nmb, cfg.ng, cfg.nx1, cfg.nx2, cfg.nx3, 2, 2, cfg.gamma, cfg.dfloor, cfg.pfloor,
                       cfg.tfloor, cfg.sfloor, cfg.sigma_max)
    h = C.c_void_p()
    ipp = C.POINTER(C.c_int)
    assert L.ak_bcs_create(C.byref(mcfg), mesh_bcs.ctypes.data_as(ipp), mb_bcs.ctypes.data_as(ipp), C.byref(h)) == 0
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    if split:
        assert L.ak_bcs_hydro(h, dev[0].data_ptr(), 5, st) == 0
        assert L.ak_bcs_bfield(h, *[t.data_ptr() for t in dev[1:]], st) == 0
    else:
        assert L.ak_bcs_apply(h, dev[0].data_ptr(), 5, *[t.data_ptr() for t in dev[1:]], st) == 0
    torch.cuda.synchronize()
    L.ak_bcs_destroy(h)
    for t, a in zip(dev, (A.u0, *A.b0)):
        assert np.array_equal(t.cpu().numpy(), a)


def test_bc_unsupported_flags_fail_loudly():
    from athenak_b200 import lib as _lib
    L = _lib.load()
    mcfg = _lib.MhdCfg(1, 4, 16, 16, 16, 2, 2, 1.4, 1e-30, 1e-30, 1e-30, 1e-30, 1e30)
    ipp = C.POINTER(C.c_int)
    for bad in ("inflow", "user", "shear_periodic"):
        flags = np.array([O.BC[bad]] * 6, np.int32)
        h = C.c_void_p()
        assert L.ak_bcs_create(C.byref(mcfg), flags.ctypes.data_as(ipp), flags.ctypes.data_as(ipp), C.byref(h)) != 0
        assert b"not supported" in L.ak_bcs_last_error(None)


def _py_run(mesh, bcs, ncyc, fused, prm):
    from athenak_b200.driver import Driver, assemble_mhd_tasks
    from athenak_b200.mhd import MHD
    mb_size, mb_bcs, nghbr = mesh.rank_tables(0)
    e = MHD(mesh.nmb_total, 4, mesh.nblock, mb_size, nghbr, recon="ppm4", rsolver="hlld", gamma=1.6666667,
            fused=fused, mesh_bcs=mesh.bcs, mb_bcs=mb_bcs)
    u0, b1, b2, b3 = mesh.pgen_blast(0, 4, 1.6666667, **prm)
    e.load_state(u0.numpy(), b1.numpy(), b2.numpy(), b3.numpy())
    drv = Driver("rk3", 0.3, 0.3)
    drv.Initialize(e)
    dt0 = drv.dt
    drv.Execute(e, assemble_mhd_tasks(e, fused=fused), ncycles=ncyc)
    torch.cuda.synchronize()
    return dt0, drv.time, e.u0.cpu().numpy(), [b.cpu().numpy() for b in e.b0]


def _cxx_run(mesh, ncyc, fused, prm):
    from athenak_b200.sim import Simulation
    sim = Simulation(mesh, 0, ng=4, recon="ppm4", rsolver="hlld", gamma=1.6666667, integrator="rk3",
                     cfl_number=0.3, tlim=0.3, fused=fused)
    sim.set_state(*mesh.pgen_blast(0, 4, 1.6666667, **prm))
    sim.initialize()
    dt0 = sim.status()["dt"]
    assert sim.execute(ncyc) == ncyc
    u, x1, x2, x3 = sim.get_state()
    return dt0, sim.status()["time"], u, [x1, x2, x3]


WIDE = dict(inner_radius=0.42, outer_radius=0.55)


@pytest.mark.parametrize("driver", ["python", "cxx"])
@pytest.mark.parametrize("fused", [False, True])
def test_run_with_physical_bcs_matches_reference_fixture(driver, fused):
    """16^3 blast crossing every mesh edge, reflect/outflow/diode mix: 3 cycles through both host
    drivers == the reference's restart dump, ghost zones included."""
    from athenak_b200.mesh import Mesh
    d = G.load("blast16_bcs_mixed")
    mesh = Mesh((16, 16, 16), (16, 16, 16), bcs=GOLD_BCS)
    dt0, time, u, b = (_py_run(mesh, GOLD_BCS, 3, fused, WIDE) if driver == "python" else _cxx_run(mesh, 3, fused, WIDE))
    assert dt0 == float(d["dt_c0"]) and time == float(d["time_c3"])
    assert np.array_equal(u, d["u0g_c3"])
    for a, k in zip(b, ("b1fg_c3", "b2fg_c3", "b3fg_c3")):
        assert np.array_equal(a, d[k])


@pytest.mark.parametrize("bcs", FLAGSETS[:3] + FLAGSETS[4:])
def test_multiblock_run_with_physical_bcs_matches_oracle(bcs):
    """8 blocks of 8^3 (block faces + physical faces + EMF correction with missing neighbours),
    C++ driver, fused list, 3 cycles == oracle_run with the same flags, ghosts included."""
    from athenak_b200.mesh import Mesh
    mesh = Mesh((16, 16, 16), (8, 8, 8), bcs=bcs)
    lloc, ngh, mbs, mesh_bcs, mb_bcs = _tables((16, 16, 16), (8, 8, 8), bcs)
    assert np.array_equal(mesh.rank_tables(0)[1], mb_bcs)
    cfg = O.make_cfg(8, 4, (8, 8, 8))
    A = O.Arrays(cfg)
    u0, b1, b2, b3 = mesh.pgen_blast(0, 4, 1.6666667, **WIDE)
    A.u0[:], A.b0[0][:], A.b0[1][:], A.b0[2][:] = u0.numpy(), b1.numpy(), b2.numpy(), b3.numpy()
    G.oracle_init_bcs(A, ngh, mesh_bcs, mb_bcs)
    dt0 = O.lib().oracle_initial_dt(C.byref(cfg), C.byref(A.state()), O.p(mbs), C.c_double(0.3), C.c_double(0.3), C.c_double(0.0))
    st = A.state()
    t, dt = C.c_double(0.0), C.c_double(dt0)
    O.lib().oracle_set_bcs(O.ip(mesh_bcs), O.ip(mb_bcs))
    try:
        O.lib().oracle_run(C.byref(cfg), C.byref(st), O.p(mbs), O.ip(ngh), 3, C.c_double(0.3), C.c_double(0.3), 3,
                           C.byref(t), C.byref(dt), None)
    finally:
        O.lib().oracle_set_bcs(None, None)
    gdt0, gtime, u, b = _cxx_run(mesh, 3, True, WIDE)
    assert gdt0 == dt0 and gtime == t.value
    assert np.array_equal(u, A.u0)
    for x, y in zip(b, A.b0):
        assert np.array_equal(x, y)
