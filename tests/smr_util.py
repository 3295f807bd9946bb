"""Static-mesh-refinement parity cases for the hydro path.  TEST INFRASTRUCTURE ONLY.
Synthetic inputs = the unmodified reference's own initial condition on the refined mesh (cycle-0 restart dump of
oracle/_ref/athena_builtin on tests/decks/lwave3d_hydro_smr.athinput: nonlinear sound wave, nested
<refined_region> blocks, periodic); the oracle is oracle/smr_oracle.c + oracle/hydro_oracle.c."""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np

import hydro_util as H
import oracle_lib as O
from athenak_b200.mesh import Mesh
from athenak_b200.rst import read_rst

HERE = os.path.dirname(os.path.abspath(__file__))
DECK = os.path.join(HERE, "decks", "lwave3d_hydro_smr.athinput")

_B8 = ("meshblock/nx1=8", "meshblock/nx2=8", "meshblock/nx3=8")
# name -> overrides.  Blocks need nx >= 2*nghost in every active direction (the coarse send boxes take nghost
# coarse cells out of a block's nx/2, src/bvals/buffs_cc.cpp:66-76).
CASES = {
    "hydro_smr2d_plm_llf_rk2": ("problem/amp=0.1", "time/tlim=0.1", "mesh/nx3=1", "meshblock/nx3=1", "mesh/nx2=32", "meshblock/nx2=8",
                                "hydro/reconstruct=plm", "hydro/rsolver=llf", "time/integrator=rk2"),
    # 3-D: 16 root blocks, one of them refined (23 blocks); the fixture keeps the ACTIVE zones only
    "hydro_smr3d_ppm4_hllc_rk3": ("problem/amp=0.1", "time/tlim=0.04", *_B8,
                                  "refined_region1/x1min=0.1", "refined_region1/x1max=0.2", "refined_region1/x2min=0.1", "refined_region1/x2max=0.2",
                                  "refined_region1/x3min=0.1", "refined_region1/x3max=0.2",
                                  "refined_region2/level=1", "refined_region2/x1min=0.1", "refined_region2/x1max=0.2",
                                  "refined_region2/x2min=0.1", "refined_region2/x2max=0.2", "refined_region2/x3min=0.1", "refined_region2/x3max=0.2"),
}
# physical boundaries on a refined mesh: HydroBCsCoarse before the prolongation, HydroBCs after it
CASES["hydro_smr2d_bcs_ppm4_hllc_rk3"] = (
    "problem/amp=0.1", "time/tlim=0.05", "mesh/nx3=1", "meshblock/nx3=1", "mesh/nx2=32", "meshblock/nx2=8",
    "mesh/ix1_bc=outflow", "mesh/ox1_bc=reflect", "mesh/ix2_bc=reflect", "mesh/ox2_bc=diode",
    "refined_region1/x1min=0.0", "refined_region1/x1max=0.3", "refined_region1/x2min=0.0", "refined_region1/x2max=0.2",
    "refined_region2/level=2", "refined_region2/x1min=2.8", "refined_region2/x1max=3.0", "refined_region2/x2min=1.4", "refined_region2/x2max=1.5")
LIVE_CASES = dict(CASES)
LIVE_CASES.update({
    "hydro_smr3d_bcs": ("problem/amp=0.1", "time/tlim=0.03", *_B8, "mesh/ix1_bc=reflect", "mesh/ox1_bc=outflow", "mesh/ix3_bc=outflow",
                        "mesh/ox3_bc=reflect", "refined_region1/x1min=0.0", "refined_region1/x1max=0.2", "refined_region1/x3min=1.3",
                        "refined_region1/x3max=1.5", "refined_region2/level=1"),
    "hydro_smr3d_3levels": ("problem/amp=0.1", "time/tlim=0.05", *_B8),                        # 135 blocks, levels 2-4
    "hydro_smr3d_wenoz_roe": ("problem/amp=0.1", "time/tlim=0.03", *_B8, "hydro/reconstruct=wenoz", "hydro/rsolver=roe", "refined_region2/level=1"),
    "hydro_smr2d_ppm4_hlle_l3": ("problem/amp=0.1", "time/tlim=0.05", "mesh/nx3=1", "meshblock/nx3=1", "mesh/nx2=32", "meshblock/nx2=8",
                                 "hydro/rsolver=hlle", "refined_region2/level=3", "refined_region2/x1min=2.6", "refined_region2/x1max=3.0",
                                 "refined_region2/x2min=1.3", "refined_region2/x2max=1.5"),
    "hydro_smr1d_plm_hllc": ("problem/amp=0.1", "time/tlim=0.2", "mesh/nx2=1", "meshblock/nx2=1", "mesh/nx3=1", "meshblock/nx3=1",
                             "mesh/nx1=128", "meshblock/nx1=8", "problem/along_x1=true", "hydro/reconstruct=plm", "mesh/nghost=2", "refined_region2/level=3"),
})


def reference(name, threads=4):
    with tempfile.TemporaryDirectory() as tmp:
        O.run_reference("athena_builtin", DECK, tmp, list(LIVE_CASES[name]), threads=threads)
        r0 = read_rst(os.path.join(tmp, "rst", "LinWave.00000.rst"), physics="hydro")
        r1 = read_rst(os.path.join(tmp, "rst", "LinWave.00001.rst"), physics="hydro")
    return r0, r1


def with_ghosts(d):
    """(u0_c0, u0_final, active-zone slices) of a fixture; 3-D fixtures hold the active zones only and are padded
    with zero ghost zones (Driver::Initialize fills them)."""
    u0, uf, ng = d["u0_c0"], d["u0_final"], int(d["ng"])
    p = H.deck_params(str(d["deck"]))
    nb = [int(p[f"meshblock/nx{k}"]) for k in (1, 2, 3)]
    if nb[2] > 1 and u0.shape[2] == nb[2]:
        full = np.zeros(u0.shape[:2] + tuple(n + 2 * ng for n in nb[::-1]))
        full[:, :, ng:-ng, ng:-ng, ng:-ng] = u0
        return full, uf, (slice(ng, -ng),) * 3
    return u0, uf, (slice(None),) * 3


def mesh_of(deck_text, nranks=1, overrides=()):
    """Mesh of a deck; overrides = command-line style "block/key=value" strings applied on top (as the reference does)."""
    p = H.deck_params(deck_text)
    for o in overrides:
        k, v = o.split("=", 1)
        p[k] = v
    nmesh = tuple(int(p[f"mesh/nx{d}"]) for d in (1, 2, 3))
    nblock = tuple(int(p[f"meshblock/nx{d}"]) for d in (1, 2, 3))
    xmin = tuple(float(p[f"mesh/x{d}min"]) for d in (1, 2, 3))
    xmax = tuple(float(p[f"mesh/x{d}max"]) for d in (1, 2, 3))
    bcs = tuple(p[f"mesh/{s}x{d}_bc"] for d in (1, 2, 3) for s in ("i", "o"))
    regs = []
    for blk in sorted({k.split("/")[0] for k in p if k.startswith("refined_region")}):
        r = [float(p[blk + "/level"])]
        for d in (1, 2, 3):
            r += [float(p.get(f"{blk}/x{d}min", xmin[d - 1])), float(p.get(f"{blk}/x{d}max", xmax[d - 1]))]
        regs.append(r)
    return Mesh(nmesh, nblock, xmin=xmin, xmax=xmax, bcs=bcs, nranks=nranks, refined_regions=regs), p


def block_bcs(mesh, p):
    """mb_bcs from the definition (meshblock.cpp:25-131) at each block's own level; (mesh_bcs, mb_bcs, strictly periodic)."""
    names = tuple(p[f"mesh/{s}x{d}_bc"] for d in (1, 2, 3) for s in ("i", "o"))
    flags = np.array([O.BC[b] for b in names], np.int32)
    out = np.zeros((mesh.nmb_total, 6), np.int32)
    for g, l in enumerate(mesh.lloc):
        for d in range(3):
            if mesh.nmesh[d] == 1:
                out[g, 2 * d], out[g, 2 * d + 1] = flags[2 * d], flags[2 * d + 1]
                continue
            n = (mesh.nmesh[d] // mesh.nblock[d]) << (int(l[3]) - mesh.root_level)
            out[g, 2 * d] = flags[2 * d] if l[d] == 0 else O.BC["block"]
            out[g, 2 * d + 1] = flags[2 * d + 1] if l[d] == n - 1 else O.BC["block"]
    active = [d for d in range(3) if mesh.nmesh[d] > 1]
    return flags, out, all(flags[2 * d] == O.BC["periodic"] for d in active)


def oracle_run(deck_text, u_init, ncycles=-1):
    """Driver::Initialize + Execute on the SMR oracle; returns (u0, w0, time, dt_next, ncycles, dt0)."""
    mesh, p = mesh_of(deck_text)
    mbs, mbb, ngh = mesh.rank_tables(0)
    flags, mb_bcs, periodic = block_bcs(mesh, p)
    assert np.array_equal(mbb, mb_bcs)
    cfg = O.OCfg(mesh.nmb_total, int(p["mesh/nghost"]), *mesh.nblock, O.RECON[p["hydro/reconstruct"]], O.HYD_RSOLVER[p["hydro/rsolver"]],
                 float(p["hydro/gamma"]), O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MAX, 0, 0, 1.0, 0)
    L = O.lib()
    L.oracle_smr_coarse_cells.restype = C.c_size_t
    L.oracle_hydro_smr_init.restype = C.c_double
    L.oracle_hydro_smr_run.restype = C.c_int
    ncc = L.oracle_smr_coarse_cells(C.byref(cfg))
    u0 = np.ascontiguousarray(u_init, np.float64).copy()
    u1, w0 = np.zeros_like(u0), np.zeros_like(u0)
    cu0 = np.zeros((mesh.nmb_total, u0.shape[1], ncc))
    ngh = np.ascontiguousarray(ngh, np.int32)
    lev = np.ascontiguousarray(mesh.lloc[:, 3], np.int32)
    cfl, tlim = float(p["time/cfl_number"]), float(p["time/tlim"])
    if not periodic:
        L.oracle_set_bcs(O.ip(flags), O.ip(mb_bcs))
    try:
        return _oracle_run(L, cfg, u0, u1, cu0, w0, mbs, ngh, lev, p, cfl, tlim, ncycles)
    finally:
        L.oracle_set_bcs(None, None)


def _oracle_run(L, cfg, u0, u1, cu0, w0, mbs, ngh, lev, p, cfl, tlim, ncycles):
    dt0 = L.oracle_hydro_smr_init(C.byref(cfg), O.p(u0), O.p(cu0), O.p(w0), O.p(mbs), O.ip(ngh), O.ip(lev), C.c_double(cfl),
                                  C.c_double(tlim), C.c_double(0.0))
    t, dt = C.c_double(0.0), C.c_double(dt0)
    n = L.oracle_hydro_smr_run(C.byref(cfg), O.p(u0), O.p(u1), O.p(cu0), O.p(w0), O.p(mbs), O.ip(ngh), O.ip(lev),
                               H.INTEGRATOR[p["time/integrator"]], C.c_double(cfl), C.c_double(tlim), ncycles, C.byref(t), C.byref(dt))
    return u0, w0, t.value, dt.value, n, dt0


# ---- MHD on refined meshes (tests/decks/lwave3d_smr.athinput: nonlinear fast wave, nested regions, periodic) -------
MHD_DECK = os.path.join(HERE, "decks", "lwave3d_smr.athinput")
MHD_CASES = {
    "mhd_smr2d_ppm4_hlld_rk3": ("problem/amp=0.1", "time/tlim=0.05", "mesh/nx3=1", "meshblock/nx3=1", "mesh/nx2=32", "meshblock/nx2=8"),
    # 3-D: 16 root blocks, one refined (23 blocks); fixture keeps the ACTIVE zones / faces only
    "mhd_smr3d_ppm4_hlld_rk3": ("problem/amp=0.1", "time/tlim=0.02", *_B8,
                                "refined_region1/x1min=0.1", "refined_region1/x1max=0.2", "refined_region1/x2min=0.1", "refined_region1/x2max=0.2",
                                "refined_region1/x3min=0.1", "refined_region1/x3max=0.2",
                                "refined_region2/level=1", "refined_region2/x1min=0.1", "refined_region2/x1max=0.2",
                                "refined_region2/x2min=0.1", "refined_region2/x2max=0.2", "refined_region2/x3min=0.1", "refined_region2/x3max=0.2"),
    # physical boundaries: coarse-array BCs (HydroBCsCoarse / BFieldBCsCoarse) before the prolongation, fine BCs after
    "mhd_smr2d_bcs": ("problem/amp=0.1", "time/tlim=0.05", "mesh/nx3=1", "meshblock/nx3=1", "mesh/nx2=32", "meshblock/nx2=8",
                      "mesh/ix1_bc=outflow", "mesh/ox1_bc=reflect", "mesh/ix2_bc=reflect", "mesh/ox2_bc=diode",
                      "refined_region1/x1min=0.0", "refined_region1/x1max=0.3", "refined_region1/x2min=0.0", "refined_region1/x2max=0.2",
                      "refined_region2/level=2", "refined_region2/x1min=2.8", "refined_region2/x1max=3.0", "refined_region2/x2min=1.4",
                      "refined_region2/x2max=1.5"),
}
MHD_LIVE_CASES = dict(MHD_CASES)
MHD_LIVE_CASES.update({
    "mhd_smr3d_3levels": ("problem/amp=0.1", "time/tlim=0.02", *_B8),                       # 135 blocks, levels 2-4
    "mhd_smr3d_plm_hlle_rk2": ("problem/amp=0.1", "time/tlim=0.03", *_B8, "mhd/reconstruct=plm", "mhd/rsolver=hlle", "time/integrator=rk2",
                               "refined_region2/level=1"),
    "mhd_smr2d_wenoz_llf_l3": ("problem/amp=0.1", "time/tlim=0.04", "mesh/nx3=1", "meshblock/nx3=1", "mesh/nx2=32", "meshblock/nx2=8",
                               "mhd/reconstruct=wenoz", "mhd/rsolver=llf", "refined_region2/level=3", "refined_region2/x1min=2.6",
                               "refined_region2/x1max=3.0", "refined_region2/x2min=1.3", "refined_region2/x2max=1.5"),
    "mhd_smr3d_bcs": ("problem/amp=0.1", "time/tlim=0.02", *_B8, "mesh/ix1_bc=reflect", "mesh/ox1_bc=outflow", "mesh/ix3_bc=outflow",
                      "mesh/ox3_bc=reflect", "refined_region1/x1min=0.0", "refined_region1/x1max=0.2", "refined_region1/x3min=1.3",
                      "refined_region1/x3max=1.5", "refined_region2/level=1"),
    "mhd_smr1d_plm_hlld": ("problem/amp=0.1", "time/tlim=0.2", "mesh/nx2=1", "meshblock/nx2=1", "mesh/nx3=1", "meshblock/nx3=1",
                           "mesh/nx1=128", "meshblock/nx1=8", "problem/along_x1=true", "mhd/reconstruct=plm", "mesh/nghost=2",
                           "refined_region2/level=3"),
})


def reference_mhd(name, threads=4):
    with tempfile.TemporaryDirectory() as tmp:
        O.run_reference("athena_builtin", MHD_DECK, tmp, list(MHD_LIVE_CASES[name]), threads=threads)
        r0 = read_rst(os.path.join(tmp, "rst", "LinWave.00000.rst"))
        r1 = read_rst(os.path.join(tmp, "rst", "LinWave.00001.rst"))
    return r0, r1


def active_faces(b, ng, nblock):
    """Active faces (block-boundary faces included) of ghost-inclusive face arrays (b1f, b2f, b3f)."""
    out = []
    for v, a in enumerate(b):
        sl = [slice(None)]
        for d in (2, 1, 0):           # array axes k, j, i
            if nblock[d] > 1 or d == 0:
                sl.append(slice(ng, -ng))       # n (+1 along the staggered axis) entries: s .. e (+1)
            else:
                sl.append(slice(None))
        out.append(a[tuple(sl)])
    return out


def oracle_run_mhd(deck_text, u_init, b_init, ncycles=-1, active_only=False):
    """Driver::Initialize + Execute of the MHD SMR oracle.  active_only: u_init / b_init hold active zones / faces
    only (ghost zones start at zero; Initialize rebuilds them).  Returns (Arrays, time, dt_next, ncycles, dt0)."""
    mesh, p = mesh_of(deck_text)
    mbs, _, ngh = mesh.rank_tables(0)
    cfg = O.OCfg(mesh.nmb_total, int(p["mesh/nghost"]), *mesh.nblock, O.RECON[p["mhd/reconstruct"]], O.RSOLVER[p["mhd/rsolver"]],
                 float(p["mhd/gamma"]), O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MAX, int(p.get("mhd/nscalars", 0)), 0, 1.0,
                 1 if p.get("mhd/fofc", "false") == "true" else 0)
    L = O.lib()
    L.oracle_smr_coarse_cells.restype = C.c_size_t
    L.oracle_smr_coarse_faces.restype = C.c_size_t
    L.oracle_mhd_smr_run.restype = C.c_int
    L.oracle_initial_dt.restype = C.c_double
    A = O.Arrays(cfg)
    ng = cfg.ng
    if active_only:
        act = tuple(slice(ng, -ng) if n > 1 else slice(None) for n in mesh.nblock[::-1])
        A.u0[(slice(None), slice(None)) + act] = u_init
        for dst, src in zip(active_faces(A.b0, ng, mesh.nblock), b_init):
            dst[:] = src
    else:
        A.u0[:] = u_init
        for dst, src in zip(A.b0, b_init):
            dst[:] = src
    cu0 = np.zeros((mesh.nmb_total, A.u0.shape[1], L.oracle_smr_coarse_cells(C.byref(cfg))))
    cb = [np.zeros((mesh.nmb_total, L.oracle_smr_coarse_faces(C.byref(cfg), v))) for v in range(3)]
    ngh = np.ascontiguousarray(ngh, np.int32)
    lev = np.ascontiguousarray(mesh.lloc[:, 3], np.int32)
    st = A.state()
    flags, mb_bcs, periodic = block_bcs(mesh, p)
    if not periodic:
        L.oracle_set_bcs(O.ip(flags), O.ip(mb_bcs))
    try:
        return _oracle_run_mhd(L, cfg, st, A, cu0, cb, mbs, ngh, lev, p, ncycles)
    finally:
        L.oracle_set_bcs(None, None)


def _oracle_run_mhd(L, cfg, st, A, cu0, cb, mbs, ngh, lev, p, ncycles):
    L.oracle_mhd_smr_ghosts_and_prims(C.byref(cfg), C.byref(st), O.p(cu0), *[O.p(x) for x in cb], O.ip(ngh), O.ip(lev))
    cfl, tlim = float(p["time/cfl_number"]), float(p["time/tlim"])
    dt0 = L.oracle_initial_dt(C.byref(cfg), C.byref(st), O.p(mbs), C.c_double(cfl), C.c_double(tlim), C.c_double(0.0))
    t, dt = C.c_double(0.0), C.c_double(dt0)
    n = L.oracle_mhd_smr_run(C.byref(cfg), C.byref(st), O.p(cu0), *[O.p(x) for x in cb], O.p(mbs), O.ip(ngh), O.ip(lev),
                             H.INTEGRATOR[p["time/integrator"]], C.c_double(cfl), C.c_double(tlim), ncycles, C.byref(t), C.byref(dt))
    return A, t.value, dt.value, n, dt0


SCALAR_DECK = os.path.join(HERE, "decks", "bw2d_smr_scalars.athinput")


def reference_mhd_deck(deck, nmhd, threads=4):
    """First and last restart dump of athena_builtin on a deck (nmhd = number of cell-centred variables incl. scalars)."""
    with tempfile.TemporaryDirectory() as tmp:
        O.run_reference("athena_builtin", deck, tmp, [], threads=threads)
        files = sorted(f for f in os.listdir(os.path.join(tmp, "rst")) if f.endswith(".rst"))
        return read_rst(os.path.join(tmp, "rst", files[0]), nmhd=nmhd), read_rst(os.path.join(tmp, "rst", files[-1]), nmhd=nmhd)

