"""CPU tests of the hydro twin: oracle/hydro_oracle.c against the committed golden fixtures
(tests/golden/hydro_*.npz, made from the unmodified reference by tests/gen_golden.py) and, when
oracle/_ref is built, against a live reference run; host mesh tables for 1-D / 2-D meshes."""
import os
import tempfile

import numpy as np
import pytest

import golden_io as G
import hydro_util as H
import oracle_lib as O

FIXTURES = ["hydro_sod1d_plm_llf_rk2", "hydro_sod1d_wenoz_roe_rk3", "hydro_sod2d_ppm4_hllc_rk3",
            "hydro_lwave3d_ppmx_hlle_rk2"]


@pytest.mark.parametrize("name", FIXTURES)
def test_hydro_oracle_matches_golden(name):
    d = G.load(name)
    case = H.Case(str(d["deck"]))
    assert np.array_equal(case.lloc, d["lloc"][:, :3])
    u, w, t, dt, n, dt0 = H.oracle_run(case, d["u0_c0"])
    assert dt0 == float(d["dt_c0"])
    assert n == int(d["ncycle_final"]) and t == float(d["time_final"]) and dt == float(d["dt_final"])
    assert np.array_equal(u, d["u0_final"]), f"max|dU| = {np.abs(u - d['u0_final']).max()}"


def test_sod_known_answer():
    """Sod as is -> 175 cycles, L1(rho) vs the reference test's analytic profile = 3.549279e-03."""
    d = G.load("hydro_sod1d_plm_llf_rk2")
    assert int(d["ncycle_final"]) == 175
    u = d["u0_final"]
    x = -0.5 + (np.arange(256) + 0.5) / 256.0
    l1 = float(np.abs(u[0, 0, 0, 0, 2:-2] - H.sod_density(x, 0.25)).mean())
    assert abs(l1 - 3.549279e-03) < 1e-7, l1     # the known answer was computed from a %12.5e tab dump


@pytest.mark.parametrize("ov", [("hydro/rsolver=hllc", "hydro/reconstruct=teno", "mesh/nghost=3", "time/integrator=rk3", "time/cfl_number=0.4"),
                                ("hydro/rsolver=hlle", "hydro/reconstruct=ppm4", "mesh/nghost=3")])
def test_hydro_oracle_matches_live_reference(ov):
    if not O.have_reference("athena_builtin"):
        pytest.skip("oracle/_ref/athena_builtin not built")
    from athenak_b200.rst import read_rst
    deck = os.path.join(os.path.dirname(os.path.abspath(__file__)), "decks", "sod1d.athinput")
    with tempfile.TemporaryDirectory() as tmp:
        O.run_reference("athena_builtin", deck, tmp, list(ov))
        r0 = read_rst(os.path.join(tmp, "rst", "Sod.00000.rst"), physics="hydro")
        r1 = read_rst(os.path.join(tmp, "rst", "Sod.00001.rst"), physics="hydro")
    case = H.Case(r0.deck)
    u, w, t, dt, n, dt0 = H.oracle_run(case, r0.u0)
    assert dt0 == r0.dt and n == r1.ncycle and t == r1.time
    assert np.array_equal(u, r1.u0)


@pytest.mark.parametrize("nmesh,nblock,bcs", [((64, 1, 1), (16, 1, 1), ("outflow", "outflow") + ("periodic",) * 4),
                                              ((32, 64, 1), (16, 16, 1), ("reflect", "reflect", "outflow", "diode", "periodic", "periodic")),
                                              ((32, 16, 1), (8, 8, 1), ("periodic",) * 6)])
def test_host_mesh_1d_2d(nmesh, nblock, bcs):
    """Product host mesh (csrc/host_mesh.cpp) == the definition (oracle_lib.uniform_mesh) on 1-D/2-D meshes:
    neighbours exist only in the active dimensions (meshblock.cpp:170-425)."""
    from athenak_b200.mesh import Mesh
    mesh = Mesh(nmesh, nblock, bcs=bcs)
    flags = np.array([O.BC[b] for b in bcs], np.int32)
    periodic = [flags[2 * d] == O.BC["periodic"] for d in range(3)]
    lloc, nghbr, mb_size, _ = O.uniform_mesh(nmesh, nblock, periodic=periodic)
    assert np.array_equal(mesh.lloc[:, :3], lloc)
    mbs, mbb, ngh = mesh.rank_tables(0)
    assert np.array_equal(ngh[:, :, 0], nghbr)
    assert np.array_equal(mbs, mb_size)
    assert np.array_equal(mbb, O.block_bcs(lloc, nmesh, nblock, flags))
    # no neighbour slot involving an inactive dimension is populated
    for oz in (-1, 0, 1):
        for oy in (-1, 0, 1):
            for ox in (-1, 0, 1):
                if (oy and nmesh[1] == 1) or (oz and nmesh[2] == 1):
                    assert (ngh[:, O.lib().oracle_neighbor_index(ox, oy, oz, 0, 0), 0] == -1).all()


def test_hydro_halo_plan_1d_2d():
    """ak_halo_create for the hydro twin on 1-D / 2-D blocks: cell-centred descriptors only, boxes
    confined to the active dimensions (host-side plan, needs no device)."""
    import ctypes as C
    from athenak_b200 import lib as L
    from athenak_b200.mesh import Mesh
    lib = L.load()
    for nmesh, nblock, ng in (((64, 1, 1), (16, 1, 1), 2), ((32, 32, 1), (16, 16, 1), 3)):
        mesh = Mesh(nmesh, nblock)
        _, _, ngh = mesh.rank_tables(0)
        cfg = L.MhdCfg(mesh.nmb_total, ng, *nblock, L.RECON["plm"], L.HYD_RSOLVER["hllc"], 1.4, 0, 0, 0, 0, 0)
        plan = C.c_void_p()
        assert lib.ak_halo_create(C.byref(cfg), ngh.ctypes.data_as(C.POINTER(C.c_int)), 0, 0, C.byref(plan)) == 0
        out = np.zeros((4096, 13), np.int64)
        n = lib.ak_halo_descs(plan, 0, 0, out.ctypes.data_as(C.POINTER(C.c_long)), 4096)
        nd = 2 if nmesh[1] == 1 else 8
        assert n == mesh.nmb_total * nd
        d = out[:n]
        if nmesh[1] == 1:
            assert (d[:, 4:6] == 0).all() and (d[:, 7:9] == 0).all()      # j,k boxes are [0,0]
        assert (d[:, 5] == 0).all() and (d[:, 8] == 0).all()               # k box is [0,0]
        assert lib.ak_halo_descs(plan, 1, 0, out.ctypes.data_as(C.POINTER(C.c_long)), 4096) == 0   # no face-field descriptors
        lib.ak_halo_destroy(plan)


@pytest.mark.parametrize("name", ["mhd_bw1d_ppm4_hlld_rk3", "mhd_bw2d_ppm4_hlld_rk3", "mhd_ot2d_plm_hlle_rk2"])
def test_mhd_oracle_1d_2d_matches_golden(name):
    """oracle/mhd_oracle.c on 1-D / 2-D meshes (CornerE 1-D/2-D branches, CT, EMF correction, BCs)
    against fixtures made from the unmodified reference."""
    d = G.load(name)
    case = H.Case(str(d["deck"]))
    A, t, dt, n, dt0 = H.oracle_run_mhd(case, d["u0_c0"], (d["b1f_c0"], d["b2f_c0"], d["b3f_c0"]))
    assert dt0 == float(d["dt_c0"]) and n == int(d["ncycle_final"]) and t == float(d["time_final"])
    assert np.array_equal(A.u0, d["u0_final"])
    for x, k in zip(A.b0, ("b1f_final", "b2f_final", "b3f_final")):
        assert np.array_equal(x, d[k])
