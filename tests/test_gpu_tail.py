"""The fused forms of the stage -- the flux + update fusion (ak_mhd_fluxes_update: sweeps that finish the conserved update, no flux
arrays) and the two-kernel stage (ak_mhd_stage_tail + shell kernels, csrc/stage_tail.cu) -- against (i) the task-by-task kernels of the same
library (AKB_TAIL=0: flux sweeps, corner_emf, update_ct, cons2prim as separate launches -- themselves bitwise against the oracle
and the reference in test_gpu_parity.py) and (ii) the reference's golden dumps.  Everything is compared bit for bit, ghost zones
and primitives included, after several cycles, so a wrong frame cell, a missed shell face, a wrong buffer rotation or a lost
new-dt contribution all show up."""
import os

import numpy as np
import pytest
import torch

import golden_io as G

pytestmark = pytest.mark.gpu


MODES = {"tail": {"AKB_TAIL": "1", "AKB_FLUXUPD": "0"},        # two-kernel stage (stage_tail.cu)
         "fluxupd": {"AKB_TAIL": "0", "AKB_FLUXUPD": "1"},     # flux + update fusion (AccArgs, flux_kernels.cuh)
         "tasks": {"AKB_TAIL": "0", "AKB_FLUXUPD": "0"}}       # one kernel per task group (round-1 fused list)


def _run(mesh, state, ncycles, mode, **kw):
    from athenak_b200.sim import Simulation
    if mode is True:
        mode = "tail"
    old = {k: os.environ.get(k) for k in MODES[mode]}
    os.environ.update(MODES[mode])
    try:
        sim = Simulation(mesh, 0, ng=kw.pop("ng", 4), cfl_number=0.3, tlim=1.0e9, fused=True, **kw)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    sim.set_state(*state)
    sim.initialize()
    dts = [sim.status()["dt"]]
    for _ in range(ncycles):
        assert sim.execute(1) == 1
        dts.append(sim.status()["dt"])
    out = sim.get_state() + list(sim.get_prim())
    st = sim.status()
    return out, dts, st


def _compare(mesh, state, ncycles, **kw):
    b, dtb, stb = _run(mesh, state, ncycles, "tasks", **dict(kw))
    scal = kw.get("nscalars", 0) > 0
    for mode in ("tail", "fluxupd"):
        a, dta, sta = _run(mesh, state, ncycles, mode, **dict(kw))
        assert dta == dtb, (mode, dta, dtb)
        assert sta["time"] == stb["time"] and sta["nfloor"] == stb["nfloor"], mode
        if not (mode == "fluxupd" and scal):        # passive scalars: the flux + update fusion falls back to the task kernels
            assert sta["launches"] < stb["launches"], mode          # the fused path really ran (fewer launches per stage)
        for name, x, y in zip(("u0", "b1f", "b2f", "b3f", "w0", "bcc0"), a, b):
            bad = np.argwhere(x != y)
            assert len(bad) == 0, f"{mode} {name}: {len(bad)} entries differ, first {bad[0]}, {x[tuple(bad[0])]!r} vs {y[tuple(bad[0])]!r}"
    return a


@pytest.mark.parametrize("nb", [1, 2])
@pytest.mark.parametrize("integrator", ["rk3", "rk2", "rk1"])
def test_tail_equals_task_kernels_blast(nb, integrator):
    from athenak_b200.mesh import Mesh
    n = 32
    m = n // nb
    mesh = Mesh((n, n, n), (m, m, m))
    _compare(mesh, mesh.pgen_blast(0, 4, 1.6666667), 5, recon="ppm4", rsolver="hlld", gamma=1.6666667, integrator=integrator)


@pytest.mark.parametrize("shape", [((48, 24, 40), (24, 12, 20)), ((36, 40, 20), (36, 10, 20)), ((70, 12, 12), (35, 6, 6))])
def test_tail_equals_task_kernels_orszag_tang_ragged_tiles(shape):
    """Orszag-Tang (the EMF correction is NOT an identity there) on blocks whose extents are not multiples of
    the tile size, with several blocks per direction: partial tiles, k-runs that start inside a tile, shell faces on every side."""
    from athenak_b200.mesh import Mesh
    nmesh, nblock = shape
    mesh = Mesh(nmesh, nblock, xmin=(-0.5, -0.5, -0.5), xmax=(0.5, 0.5, 0.5))
    u0, b1, b2, b3 = mesh.pgen_orszag_tang(0, 4, 1.6666667)
    _compare(mesh, (u0, b1, b2, b3), 4, recon="ppm4", rsolver="hlld", gamma=1.6666667, integrator="rk3")


@pytest.mark.parametrize("recon,rsolver,ng", [("plm", "hlle", 2), ("wenoz", "llf", 3), ("ppmx", "hlld", 4)])
def test_tail_other_methods_and_ghost_widths(recon, rsolver, ng):
    from athenak_b200.mesh import Mesh
    mesh = Mesh((32, 32, 32), (16, 16, 16))
    _compare(mesh, mesh.pgen_blast(0, ng, 1.6666667), 4, recon=recon, rsolver=rsolver, gamma=1.6666667, integrator="rk2", ng=ng)


def test_tail_isothermal_and_scalars():
    from athenak_b200.mesh import Mesh
    mesh = Mesh((32, 32, 32), (16, 16, 16))
    u0, b1, b2, b3 = mesh.pgen_blast(0, 4, 1.6666667)
    _compare(mesh, (u0[:, :4].contiguous(), b1, b2, b3), 4, recon="ppm4", rsolver="hlld", integrator="rk3", eos="isothermal",
             iso_sound_speed=1.0)
    us = torch.cat([u0, 0.5 * u0[:, :1], 0.25 * u0[:, :1] * (1.0 + u0[:, 4:5])], 1).contiguous()
    _compare(mesh, (us, b1, b2, b3), 4, recon="ppm4", rsolver="hlld", gamma=1.6666667, integrator="rk3", nscalars=2)


def test_tail_floors_fire():
    """A state that trips the density and energy floors inside the tail kernel's c2p and in the shell kernel: counters and
    floored values must equal the task kernels'."""
    from athenak_b200.mesh import Mesh
    mesh = Mesh((32, 32, 32), (16, 16, 16))
    u0, b1, b2, b3 = mesh.pgen_blast(0, 4, 1.6666667)
    u0 = u0.clone()
    u0[:, 4] *= 0.02          # total energy far below kinetic + magnetic in the blast region -> energy floor
    u0[:, 0, 10:14, 3:9, 5:30] = 1.0e-9
    a = _compare(mesh, (u0, b1, b2, b3), 3, recon="plm", rsolver="llf", gamma=1.6666667, integrator="rk2", dfloor=1.0e-3, pfloor=1.0e-4)
    assert np.isfinite(a[0]).all()


@pytest.mark.parametrize("mode", ["tail", "fluxupd"])
@pytest.mark.parametrize("nb", [1, 2])
def test_tail_matches_reference_golden(nb, mode):
    """End to end against the reference's restart dump (16^3 blast, 4 cycles), through the two-kernel stage."""
    from athenak_b200.mesh import Mesh
    d = G.load("blast16_ppm4_hlld_rk3")
    m = 16 // nb
    mesh = Mesh((16, 16, 16), (m, m, m))
    (u, x1, x2, x3, _, _), dts, st = _run(mesh, mesh.pgen_blast(0, 4, 1.6666667), 4, mode, recon="ppm4", rsolver="hlld",
                                           gamma=1.6666667, integrator="rk3")
    assert dts[0] == float(d["dt_c0"]) and st["time"] == float(d["time_c4"])
    for g in range(mesh.nmb_total):
        lx, ly, lz = (int(v) * m for v in mesh.lloc[g, :3])
        assert np.array_equal(u[g, :, 4:-4, 4:-4, 4:-4], d["u0_c4"][0][:, lz:lz + m, ly:ly + m, lx:lx + m])
        assert np.array_equal(x1[g, 4:-4, 4:-4, 4:-4], d["b1f_c4"][0][lz:lz + m, ly:ly + m, lx:lx + m + 1])
        assert np.array_equal(x2[g, 4:-4, 4:-4, 4:-4], d["b2f_c4"][0][lz:lz + m, ly:ly + m + 1, lx:lx + m])
        assert np.array_equal(x3[g, 4:-4, 4:-4, 4:-4], d["b3f_c4"][0][lz:lz + m + 1, ly:ly + m, lx:lx + m])
