"""MHD on statically refined meshes on the GPU (BASELINE config 5's fine/coarse data path for
face-centred fields): RestrictFC, SendB/RecvB across levels as a gather with host-resolved winners, FillCoarseInBndryFC,
divergence-preserving ProlongateFC and the fine/coarse EMF correction (csrc/smr.cu + smr_fc_body.h, ak_smr_*_fc,
ak_smr_emf_correct) under the multilevel MHD task list.
Gates: (1) each step on RANDOM data == oracle/smr_fc_oracle.c bitwise (every element differs, so owners / order /
boxes are all observable); (2) whole runs == the unmodified reference's dumps bitwise -- committed fixtures always,
live reference runs (3 levels in 3-D, PLM+HLLE+RK2, WENOZ+LLF level-3 2-D, 1-D) when oracle/_ref is present, ghost
zones included -- with identical dt sequences."""
import ctypes as C

import numpy as np
import pytest

import golden_io as G
import oracle_lib as O
import smr_util as S

pytestmark = pytest.mark.gpu
_CACHE = {}


def make_engine(deck_text, overrides=()):
    from athenak_b200.mhd import MHD
    mesh, p = S.mesh_of(deck_text, overrides=overrides)
    mb_size, mb_bcs, nghbr = mesh.rank_tables(0)
    mhd = MHD(mesh.nmb_total, int(p["mesh/nghost"]), mesh.nblock, mb_size, nghbr, recon=p["mhd/reconstruct"], rsolver=p["mhd/rsolver"],
              gamma=float(p["mhd/gamma"]), mb_lev=mesh.lloc[:, 3], mesh_bcs=mesh.bcs, mb_bcs=mb_bcs)
    return mhd, mesh, p


def run_engine(deck_text, u_init, b_init, fused, ncycles=None, active_only=False):
    import torch
    from athenak_b200.driver import Driver, assemble_mhd_tasks
    mhd, mesh, p = make_engine(deck_text)
    ng = int(p["mesh/nghost"])
    if active_only:
        act = tuple(slice(ng, -ng) if n > 1 else slice(None) for n in mesh.nblock[::-1])
        u = np.zeros(tuple(mhd.u0.shape))
        u[(slice(None), slice(None)) + act] = u_init
        b = [np.zeros(tuple(a.shape)) for a in mhd.b0]
        for dst, src in zip(S.active_faces(b, ng, mesh.nblock), b_init):
            dst[:] = src
        mhd.load_state(u, *b)
    else:
        mhd.load_state(u_init, *b_init)
    drv = Driver(p["time/integrator"], float(p["time/cfl_number"]), tlim=float(p["time/tlim"]))
    tl = assemble_mhd_tasks(mhd, fused=fused)
    drv.Initialize(mhd)
    dt0 = drv.dt
    drv.Execute(mhd, tl, ncycles=ncycles)
    torch.cuda.synchronize()
    return mhd, mesh, p, drv, dt0


class RandomSetup:
    """Engine + oracle configuration of one refined mesh, random fields."""

    def __init__(self, name):
        self.mhd, self.mesh, p = make_engine(open(S.MHD_DECK).read(), S.MHD_LIVE_CASES[name])
        _, _, ngh = self.mesh.rank_tables(0)
        self.ngh = np.ascontiguousarray(ngh, np.int32)
        self.lev = np.ascontiguousarray(self.mesh.lloc[:, 3], np.int32)
        self.ocfg = O.OCfg(self.mesh.nmb_total, int(p["mesh/nghost"]), *self.mesh.nblock, O.RECON["plm"], O.RSOLVER["hlld"], 5.0 / 3.0,
                           O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MAX, 0, 0, 1.0, 0)


def _setup(name):
    if name not in _CACHE:
        _CACHE[name] = RandomSetup(name)
    return _CACHE[name]


def _dev(arrs, like):
    import torch
    for t, a in zip(like, arrs):
        t.copy_(torch.from_numpy(a).reshape(t.shape))


@pytest.mark.parametrize("name", sorted(S.MHD_LIVE_CASES))
def test_fc_steps_random_data_match_oracle(name):
    import torch
    s = _setup(name)
    mhd, OL = s.mhd, O.lib()
    rng = np.random.default_rng(11)
    b = [rng.standard_normal(tuple(a.shape)) for a in mhd.b0]
    cb = [rng.standard_normal(tuple(a.shape)) for a in mhd.coarse_b0]
    e = [rng.standard_normal(tuple(a.shape)) for a in mhd.efld]
    st = mhd._st()
    bp, cbp = [a.data_ptr() for a in mhd.b0], [a.data_ptr() for a in mhd.coarse_b0]

    def check(tag, ob, ocb):
        torch.cuda.synchronize()
        for v in range(3):
            assert np.array_equal(mhd.b0[v].cpu().numpy(), ob[v]), f"{tag}: fine x{v + 1}f"
            assert np.array_equal(mhd.coarse_b0[v].cpu().numpy(), ocb[v]), f"{tag}: coarse x{v + 1}f"

    # RestrictB + SendB/RecvB
    ob, ocb = [x.copy() for x in b], [x.copy() for x in cb]
    OL.oracle_smr_restrict_fc(C.byref(s.ocfg), *[O.p(x) for x in ob], *[O.p(x) for x in ocb])
    OL.oracle_smr_halo_fc(C.byref(s.ocfg), *[O.p(x) for x in ob], *[O.p(x) for x in ocb], O.ip(s.ngh), O.ip(s.lev))
    _dev(b, mhd.b0); _dev(cb, mhd.coarse_b0)
    mhd._smr_ck(mhd.L.ak_smr_restrict_fc(mhd.smr, *bp, *cbp, st))
    mhd._smr_ck(mhd.L.ak_smr_exchange_fc(mhd.smr, *bp, *cbp, st))
    check("exchange", ob, ocb)
    # FillCoarseInBndryFC, ProlongateFC
    for steps in (1, 2, 3):
        ob, ocb = [x.copy() for x in b], [x.copy() for x in cb]
        OL.oracle_smr_prolongate_fc_steps(C.byref(s.ocfg), *[O.p(x) for x in ob], *[O.p(x) for x in ocb], O.ip(s.ngh), O.ip(s.lev), steps)
        _dev(b, mhd.b0); _dev(cb, mhd.coarse_b0)
        mhd._smr_ck(mhd.L.ak_smr_prolongate_fc_steps(mhd.smr, *bp, *cbp, steps, st))
        check(f"prolongate steps={steps}", ob, ocb)
    # SendE / RecvE
    oe = [x.copy() for x in e]
    OL.oracle_smr_emf_correct(C.byref(s.ocfg), *[O.p(x) for x in oe], O.ip(s.ngh), O.ip(s.lev))
    _dev(e, mhd.efld)
    mhd._smr_ck(mhd.L.ak_smr_emf_correct(mhd.smr, *[a.data_ptr() for a in mhd.efld], st))
    torch.cuda.synchronize()
    for v in range(3):
        assert np.array_equal(mhd.efld[v].cpu().numpy(), oe[v]), f"emf e{v + 1}"


@pytest.mark.parametrize("name", sorted(S.MHD_CASES))
@pytest.mark.parametrize("fused", [True, False], ids=["fused_c2p", "pertask"])
def test_mhd_smr_matches_golden_bitwise(name, fused):
    d = G.load(name)
    active_only = bool(d["active_only"])
    b0 = (d["b1f_c0"], d["b2f_c0"], d["b3f_c0"])
    mhd, mesh, p, drv, dt0 = run_engine(str(d["deck"]), d["u0_c0"], b0, fused, active_only=active_only)
    assert dt0 == float(d["dt_c0"])
    assert drv.ncycle == int(d["ncycle_final"]) and drv.time == float(d["time_final"]) and drv.dt == float(d["dt_final"])
    ng = int(p["mesh/nghost"])
    u, b = mhd.u0.cpu().numpy(), [a.cpu().numpy() for a in mhd.b0]
    if active_only:
        act = tuple(slice(ng, -ng) if k > 1 else slice(None) for k in mesh.nblock[::-1])
        u, b = u[(slice(None), slice(None)) + act], S.active_faces(b, ng, mesh.nblock)
    assert np.array_equal(u, d["u0_final"]), f"max|dU| = {np.abs(u - d['u0_final']).max()}"
    for x, k in zip(b, ("b1f_final", "b2f_final", "b3f_final")):
        assert np.array_equal(x, d[k]), f"{k}: max|dB| = {np.abs(x - d[k]).max()}"
    assert mhd.launches > 0


@pytest.mark.parametrize("name", sorted(S.MHD_LIVE_CASES))
def test_mhd_smr_matches_oracle_bitwise(name):
    """Whole runs against oracle/smr_fc_oracle.c from the same start (fixture ICs where committed, otherwise a smooth
    synthetic state): ghost zones, dt sequence and cycle count included.  Runs without the reference binary."""
    deck = open(S.MHD_DECK).read()
    mesh, p = S.mesh_of(deck, overrides=S.MHD_LIVE_CASES[name])
    txt = deck_with(deck, S.MHD_LIVE_CASES[name])
    from athenak_b200.mhd import MHD  # noqa: F401  (fail early if the extension is missing)
    u, b = synthetic_state(mesh, p)
    A, t, dt, n, dt0 = S.oracle_run_mhd(txt, u, b, ncycles=3)
    mhd, mesh, p, drv, edt0 = run_engine(txt, u, b, True, ncycles=3)
    assert edt0 == dt0 and drv.ncycle == n and drv.time == t and drv.dt == dt
    assert np.array_equal(mhd.u0.cpu().numpy(), A.u0), f"max|dU| = {np.abs(mhd.u0.cpu().numpy() - A.u0).max()}"
    for x, y in zip(mhd.b0, A.b0):
        assert np.array_equal(x.cpu().numpy(), y)


@pytest.mark.parametrize("name", sorted(set(S.MHD_LIVE_CASES) - set(S.MHD_CASES)))
def test_mhd_smr_matches_live_reference_bitwise(name):
    if not O.have_reference("athena_builtin"):
        pytest.skip("oracle/_ref/athena_builtin not built")
    r0, r1 = S.reference_mhd(name)
    mhd, mesh, p, drv, dt0 = run_engine(r0.deck, r0.u0, (r0.b1f, r0.b2f, r0.b3f), True)
    assert dt0 == r0.dt and drv.ncycle == r1.ncycle and drv.time == r1.time and drv.dt == r1.dt
    u = mhd.u0.cpu().numpy()
    assert np.array_equal(u, r1.u0), f"max|dU| = {np.abs(u - r1.u0).max()}"        # ghost zones included
    for x, y in zip(mhd.b0, (r1.b1f, r1.b2f, r1.b3f)):
        assert np.array_equal(x.cpu().numpy(), y)


def _sim_of(mesh, p, fused):
    from athenak_b200.sim import Simulation
    return Simulation(mesh, 0, ng=int(p["mesh/nghost"]), recon=p["mhd/reconstruct"], rsolver=p["mhd/rsolver"], gamma=float(p["mhd/gamma"]),
                      integrator=p["time/integrator"], cfl_number=float(p["time/cfl_number"]), tlim=float(p["time/tlim"]), fused=fused,
                      physics="mhd", multilevel=True)


@pytest.mark.parametrize("name", sorted(S.MHD_LIVE_CASES))
@pytest.mark.parametrize("fused", [True, False], ids=["fused_c2p", "pertask"])
def test_mhd_smr_cxx_host_driver(name, fused):
    """The C++ host driver (csrc/host_driver.cpp: ak_sim_set_levels for MHD, multilevel MHD task list) against the
    oracle from the same synthetic start: ghost zones, dt sequence and cycle count included."""
    deck = open(S.MHD_DECK).read()
    txt = deck_with(deck, S.MHD_LIVE_CASES[name])
    mesh, p = S.mesh_of(txt)
    u, b = synthetic_state(mesh, p)
    A, t, dt, n, dt0 = S.oracle_run_mhd(txt, u, b, ncycles=2)
    sim = _sim_of(mesh, p, fused)
    sim.set_state(np.ascontiguousarray(u), *[np.ascontiguousarray(x) for x in b])
    sim.initialize()
    assert sim.status()["dt"] == dt0
    sim.execute(2)
    st = sim.status()
    assert st["ncycle"] == n and st["time"] == t and st["dt"] == dt and st["launches"] > 0
    out = sim.get_state()
    assert np.array_equal(out[0], A.u0), f"max|dU| = {np.abs(out[0] - A.u0).max()}"
    for x, y in zip(out[1:], A.b0):
        assert np.array_equal(x, y)


def _blast_smr_mesh(nroot=128, nb=32):
    from athenak_b200.mesh import Mesh
    return Mesh((nroot,) * 3, (nb,) * 3, refined_regions=[(1, -0.25, 0.25, -0.25, 0.25, -0.25, 0.25), (2, -0.125, 0.125, -0.125, 0.125, -0.125, 0.125)])


def test_mhd_smr_blast_three_levels_divb_conservation_and_oracle():
    """BASELINE config 5's static variant at working size: 3-D MHD blast, three levels, 176 blocks of 32^3 (5.8 M cells),
    PPM4+HLLD+RK3.  Size-independent properties: div B = 0 to round-off in every active cell of every level after
    several cycles (divergence-preserving prolongation + EMF correction), mass / momentum / energy conserved to
    round-off on the refined mesh (flux correction); and the first cycles equal the oracle bitwise, ghost zones included."""
    from athenak_b200.sim import Simulation
    ng, gamma = 4, 1.6666667
    mesh = _blast_smr_mesh()
    assert mesh.nmb_total == 176 and len(np.unique(mesh.lloc[:, 3])) == 3
    u0, b1, b2, b3 = [t.numpy() for t in mesh.pgen_blast(0, ng, gamma)]
    mbs = np.asarray(mesh.rank_tables(0)[0]).reshape(-1, 9)
    vol = (mbs[:, 6] * mbs[:, 7] * mbs[:, 8])[:, None]
    a = slice(ng, -ng)

    def totals(u):
        return (u[:, :, a, a, a].sum(axis=(2, 3, 4)) * vol).sum(axis=0)

    def max_divb(b):
        d = ((b[0][:, a, a, ng + 1:-ng] - b[0][:, a, a, ng:-ng - 1]) / mbs[:, 6, None, None, None]
             + (b[1][:, a, ng + 1:-ng, a] - b[1][:, a, ng:-ng - 1, a]) / mbs[:, 7, None, None, None]
             + (b[2][:, ng + 1:-ng, a, a] - b[2][:, ng:-ng - 1, a, a]) / mbs[:, 8, None, None, None])
        return np.abs(d * mbs[:, 6, None, None, None]).max()

    def sim_of():
        sim = Simulation(mesh, 0, ng=ng, recon="ppm4", rsolver="hlld", gamma=gamma, integrator="rk3", cfl_number=0.3, tlim=1.0,
                         fused=True, physics="mhd", multilevel=True)
        sim.set_state(u0, b1, b2, b3)
        sim.initialize()
        return sim

    sim = sim_of()
    t0 = totals(u0)
    sim.execute(6)
    out = sim.get_state()
    assert sim.status()["ncycle"] == 6 and np.isfinite(out[0]).all()
    t1 = totals(out[0])
    assert abs(t1[0] - t0[0]) <= 1e-12 * abs(t0[0]), (t0[0], t1[0])             # mass
    assert abs(t1[4] - t0[4]) <= 1e-12 * abs(t0[4]), (t0[4], t1[4])             # total energy
    assert np.abs(t1[1:4] - t0[1:4]).max() <= 1e-12 * abs(t0[4])                 # momentum (zero initially)
    assert max_divb(out[1:]) <= 1e-12, max_divb(out[1:])                         # |div B| dx / B0, B0 = 1
    assert np.abs(out[0][:, 1:4]).max() > 1e-3                                   # the blast is moving
    # first cycles against the oracle (same ICs), bitwise with ghost zones
    deck = ("<mesh>\nnghost = 4\nnx1 = 128\nx1min = -0.5\nx1max = 0.5\nix1_bc = periodic\nox1_bc = periodic\n"
            "nx2 = 128\nx2min = -0.5\nx2max = 0.5\nix2_bc = periodic\nox2_bc = periodic\n"
            "nx3 = 128\nx3min = -0.5\nx3max = 0.5\nix3_bc = periodic\nox3_bc = periodic\n\n"
            "<meshblock>\nnx1 = 32\nnx2 = 32\nnx3 = 32\n\n<mesh_refinement>\nrefinement = static\n\n"
            "<refined_region1>\nlevel = 1\nx1min = -0.25\nx1max = 0.25\nx2min = -0.25\nx2max = 0.25\nx3min = -0.25\nx3max = 0.25\n\n"
            "<refined_region2>\nlevel = 2\nx1min = -0.125\nx1max = 0.125\nx2min = -0.125\nx2max = 0.125\nx3min = -0.125\nx3max = 0.125\n\n"
            "<time>\nintegrator = rk3\ncfl_number = 0.3\ntlim = 1.0\n\n<mhd>\nreconstruct = ppm4\nrsolver = hlld\ngamma = 1.6666667\n")
    m2, _ = S.mesh_of(deck)
    assert np.array_equal(m2.lloc, mesh.lloc)
    A, t, dt, n, dt0 = S.oracle_run_mhd(deck, u0, (b1, b2, b3), ncycles=1)
    sim = sim_of()
    assert sim.status()["dt"] == dt0
    sim.execute(1)
    st = sim.status()
    assert st["time"] == t and st["dt"] == dt
    out = sim.get_state()
    assert np.array_equal(out[0], A.u0)
    for x, y in zip(out[1:], A.b0):
        assert np.array_equal(x, y)


def deck_with(deck_text, overrides):
    """Deck text with command-line style overrides folded in (the engine and the oracle both parse text)."""
    import hydro_util as H
    p = H.deck_params(deck_text)
    for o in overrides:
        k, v = o.split("=", 1)
        p[k] = v
    blocks = {}
    for k, v in p.items():
        blk, key = k.split("/", 1)
        blocks.setdefault(blk, []).append((key, v))
    return "".join(f"<{blk}>\n" + "".join(f"{k} = {v}\n" for k, v in kv) + "\n" for blk, kv in blocks.items())


def synthetic_state(mesh, p):
    """Smooth periodic state with div B = 0 to round-off at every level: B from a vector potential evaluated at
    the block's own edges, ghost zones zero (Driver::Initialize fills them)."""
    ng = int(p["mesh/nghost"])
    nb = mesh.nblock
    act3 = [True, nb[1] > 1, nb[2] > 1]
    n = [nb[d] + 2 * ng if act3[d] else 1 for d in range(3)]
    mbs, _, _ = mesh.rank_tables(0)
    mbs = np.asarray(mbs).reshape(-1, 9)
    nmb = mesh.nmb_total
    xmin = [float(p[f"mesh/x{d}min"]) for d in (1, 2, 3)]
    xmax = [float(p[f"mesh/x{d}max"]) for d in (1, 2, 3)]
    Lx = [xmax[d] - xmin[d] for d in range(3)]
    u = np.zeros((nmb, 5, n[2], n[1], n[0]))
    b = [np.zeros((nmb, n[2], n[1], n[0] + 1)), np.zeros((nmb, n[2], n[1] + 1, n[0])), np.zeros((nmb, n[2] + 1, n[1], n[0]))]
    s = [ng if act3[d] else 0 for d in range(3)]
    two_pi = 2.0 * np.pi
    for m in range(nmb):
        x1min, x2min, x3min, x1max, x2max, x3max, dx1, dx2, dx3 = mbs[m]       # RegionSize (mesh.hpp:25-29)
        lo, dx = [x1min, x2min, x3min], [dx1, dx2, dx3]
        # node and centre coordinates of the active region
        nd = [lo[d] + dx[d] * np.arange(nb[d] + 1) for d in range(3)]
        ce = [lo[d] + dx[d] * (np.arange(nb[d]) + 0.5) for d in range(3)]
        ph = lambda x, d: two_pi * (x - xmin[d]) / Lx[d]
        # A3(x,y) and A1(y,z): B1 = dA3/dy, B2 = -dA3/dx + dA1/dz, B3 = -dA1/dy  (differences of edge values)
        A3 = lambda x, y: 0.3 * np.cos(ph(x, 0)) * (np.cos(ph(y, 1)) if act3[1] else 1.0)
        A1 = lambda y, z: (0.2 * np.sin(ph(y, 1)) * np.sin(ph(z, 2))) if act3[2] else 0.0 * y
        X1n, X2n, X3n = nd
        if act3[1]:
            a3 = A3(X1n[None, :], X2n[:, None])                           # (j node, i node)
            b1 = (a3[1:, :] - a3[:-1, :]) / dx2                            # (j, i node)
            b2 = -(a3[:, 1:] - a3[:, :-1]) / dx1                           # (j node, i)
        else:
            b1 = np.full((1, nb[0] + 1), 0.5)
            b2 = 0.4 + 0.1 * np.sin(ph(ce[0], 0))[None, :] * np.ones((2, 1))
        K = nb[2] if act3[2] else 1
        b1f = np.broadcast_to(b1[None], (K,) + b1.shape).copy() + 0.5
        b2f = np.broadcast_to(b2[None], (K,) + b2.shape).copy()
        if act3[2]:
            a1 = A1(X2n[None, :], X3n[:, None])                           # (k node, j node)
            b2f += ((a1[1:, :] - a1[:-1, :]) / dx3)[:, :, None]
            b3 = -(a1[:, 1:] - a1[:, :-1]) / dx2                           # (k node, j)
            b3f = np.broadcast_to(b3[:, :, None], (nb[2] + 1, nb[1], nb[0])).copy() + 0.3
        else:
            b3f = np.full((2, b2.shape[0] - (1 if act3[1] else 1), nb[0]), 0.3) if act3[1] else np.full((2, 1, nb[0]), 0.3 + 0.0)
            if not act3[1]:
                b3f = 0.3 + 0.1 * np.cos(ph(ce[0], 0))[None, None, :] * np.ones((2, 1, 1))
        sl = lambda d, extra=0: slice(s[d], s[d] + nb[d] + extra) if act3[d] else slice(0, 1 + extra)
        b[0][m, sl(2), sl(1), sl(0, 1)] = b1f
        b[1][m, sl(2), sl(1, 1), sl(0)] = b2f if act3[1] else b2f[:, :2, :].reshape(K, 2, nb[0])
        b[2][m, sl(2, 1), sl(1), sl(0)] = b3f
        c1 = ph(ce[0], 0)[None, None, :]
        c2 = ph(ce[1], 1)[None, :, None] if act3[1] else np.zeros((1, 1, 1))
        c3 = ph(ce[2], 2)[:, None, None] if act3[2] else np.zeros((1, 1, 1))
        rho = 1.0 + 0.2 * np.sin(c1) * np.cos(c2) * np.cos(c3)
        u[m, 0, sl(2), sl(1), sl(0)] = rho
        u[m, 1, sl(2), sl(1), sl(0)] = rho * 0.3 * np.cos(c1 + c2)
        u[m, 2, sl(2), sl(1), sl(0)] = rho * 0.2 * np.sin(c1 - c3)
        u[m, 3, sl(2), sl(1), sl(0)] = rho * 0.1 * np.cos(c2 + c3)
        u[m, 4, sl(2), sl(1), sl(0)] = 2.5 + 0.3 * np.cos(c1) * np.cos(c2)
    return u, b
