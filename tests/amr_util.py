"""AMR parity harness (dynamic refinement).  TEST INFRASTRUCTURE ONLY.
Follows an unmodified reference AMR run that dumps a restart file every cycle: from the reference's state of cycle k
(the dump is written BEFORE that cycle's AMR step, driver.cpp:432-450) the oracle evaluates the refinement criteria,
adapts the mesh (C++ host mesh, ak_mesh_adapt), moves / restricts / prolongates the data (oracle/amr_oracle.c), rebuilds
ghost zones and primitives, repairs the interior faces (RepairAMRFC), takes the new time step and advances ONE cycle
with the multilevel MHD oracle; the result must equal the reference's dump of cycle k+1 bitwise, ghost zones included."""
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
DECK = os.path.join(HERE, "decks", "blast2d_amr.athinput")
HYDRO_DECK = os.path.join(HERE, "decks", "blast2d_hydro_amr.athinput")
METHOD = {"min_max": 0, "slope": 1, "second_deriv": 2, "location": 3}


def reference_dumps(overrides=(), threads=4, deck=DECK, physics="mhd"):
    """Restart objects of cycle 0, 1, ... (the final dump repeats the last cycle and is dropped)."""
    with tempfile.TemporaryDirectory() as tmp:
        O.run_reference("athena_blast", deck, tmp, list(overrides), threads=threads)
        files = sorted(f for f in os.listdir(os.path.join(tmp, "rst")) if f.endswith(".rst"))
        out, seen = [], set()
        for f in files:
            r = read_rst(os.path.join(tmp, "rst", f), physics=physics)
            if r.ncycle not in seen:
                seen.add(r.ncycle)
                out.append(r)
    return out


def params(overrides=(), deck=DECK):
    p = H.deck_params(open(deck).read())
    for o in overrides:
        k, v = o.split("=", 1)
        p[k] = v
    return p


def make_mesh(p):
    g = lambda fmt: tuple(type_(p[fmt.format(d)]) for d in (1, 2, 3))
    type_ = int
    nmesh, nblock = g("mesh/nx{}"), g("meshblock/nx{}")
    type_ = float
    xmin, xmax = g("mesh/x{}min"), g("mesh/x{}max")
    mesh = Mesh(nmesh, nblock, xmin=xmin, xmax=xmax)
    mesh.set_amr(int(p["mesh_refinement/num_levels"]), int(p.get("mesh_refinement/refinement_interval", 5)),
                 int(p.get("mesh_refinement/ncycle_check", 1)))
    return mesh


def ocfg(mesh, p, nmb=None):
    return O.OCfg(mesh.nmb_total if nmb is None else nmb, int(p["mesh/nghost"]), *mesh.nblock, O.RECON[p["mhd/reconstruct"]],
                  O.RSOLVER[p["mhd/rsolver"]], float(p["mhd/gamma"]), O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MAX, 0, 0, 1.0,
                  1 if p.get("mhd/fofc", "false") == "true" else 0)


def criteria_flags(L, cfg, p, w0, mbs):
    flag = np.zeros(cfg.nmb, np.int32)
    for blk in sorted({k.split("/")[0] for k in p if k.startswith("amr_criterion")}):
        loc = np.array([float(p.get(f"{blk}/location_x{d}", 0.0)) for d in (1, 2, 3)] + [float(p.get(f"{blk}/location_rad", 0.0))])
        assert p[f"{blk}/variable"] in ("mhd_w_d", "mhd_u_d")
        L.oracle_amr_criterion(C.byref(cfg), METHOD[p[f"{blk}/method"]], O.p(w0), w0.shape[1], 0,
                               C.c_double(float(p.get(f"{blk}/value_min", -O.FLT_MAX))), C.c_double(float(p.get(f"{blk}/value_max", O.FLT_MAX))),
                               O.p(loc), O.p(mbs), O.ip(flag))
    return flag


class Stepper:
    """Oracle state that follows the reference one cycle at a time."""

    def __init__(self, p):
        self.p, self.mesh, self.L = p, make_mesh(p), O.lib()
        L = self.L
        L.oracle_smr_coarse_cells.restype = C.c_size_t
        L.oracle_smr_coarse_faces.restype = C.c_size_t
        L.oracle_mhd_smr_run.restype = C.c_int
        L.oracle_new_dt.restype = C.c_double
        self.nadapt = self.ncreated = self.ndeleted = 0

    def tables(self):
        mbs, _, ngh = self.mesh.rank_tables(0)
        return np.ascontiguousarray(mbs), np.ascontiguousarray(ngh, np.int32), np.ascontiguousarray(self.mesh.lloc[:, 3], np.int32)

    def coarse(self, cfg, nmb):
        L = self.L
        return (np.zeros((nmb, 5, L.oracle_smr_coarse_cells(C.byref(cfg)))),
                [np.zeros((nmb, L.oracle_smr_coarse_faces(C.byref(cfg), v))) for v in range(3)])

    def move(self, cfg, A, what, n2o, o2n, lloc_old, lloc_new, refined):
        """RedistAndRefineMeshBlocks steps 5-9 with the oracle (in place, the reference's copy order)."""
        L, mesh = self.L, self.mesh
        old_nmb, new_nmb = cfg.nmb, mesh.nmb_total
        cap = max(old_nmb, new_nmb)
        pad = lambda a: np.concatenate([a, np.zeros((cap - old_nmb,) + a.shape[1:])]) if cap > old_nmb else a.copy()
        u = pad(A.u0)
        b = [pad(x) for x in A.b0]
        cu, cb = self.coarse(cfg, cap)
        L.oracle_smr_restrict_cc(C.byref(cfg), O.p(u), O.p(cu), 5)       # coarse_u0 as the last stage's RestrictU left it
        L.oracle_amr_redistribute(C.byref(cfg), new_nmb, O.p(u), O.p(cu), 5, *[O.p(x) for x in b], *[O.p(x) for x in cb],
                                  O.ip(what), O.ip(n2o), O.ip(o2n), O.ip(lloc_old), O.ip(lloc_new))
        cfgn = ocfg(mesh, self.p)
        L.oracle_amr_refine_fc(C.byref(cfgn), *[O.p(x) for x in b], *[O.p(x) for x in cb], O.ip(refined), 1)
        return u, b

    def repair(self, cfg, A, cbn, repair):
        self.L.oracle_amr_refine_fc(C.byref(cfg), *[O.p(x) for x in A.b0], *[O.p(x) for x in cbn], O.ip(repair), 2)

    def flags(self, cfg, A, mbs):
        return criteria_flags(self.L, cfg, self.p, A.w0, mbs)

    def advance(self, r, ncycle):
        """r = the reference's state of cycle `ncycle` (before its AMR step).  Returns (Arrays, time, dt_used) after the
        AMR step and one more cycle."""
        p, L, mesh = self.p, self.L, self.mesh
        assert np.array_equal(mesh.lloc, r.lloc), f"block list differs at cycle {ncycle}"
        cfg = ocfg(mesh, p)
        A = O.Arrays(cfg)
        A.u0[:] = r.u0
        for dst, src in zip(A.b0, (r.b1f, r.b2f, r.b3f)):
            dst[:] = src
        A.c2p_all()
        mbs, ngh, lev = self.tables()
        flag = self.flags(cfg, A, mbs)
        lloc_old = np.ascontiguousarray(mesh.lloc, np.int32).copy()
        nnew, ndel, what = mesh.adapt(ncycle, flag)
        if nnew or ndel:
            self.nadapt += 1
            self.ncreated += nnew
            self.ndeleted += ndel
            old_nmb, new_nmb = cfg.nmb, mesh.nmb_total
            lloc_new = np.ascontiguousarray(mesh.lloc, np.int32)
            n2o, o2n = np.ascontiguousarray(mesh.newtoold, np.int32), np.ascontiguousarray(mesh.oldtonew, np.int32)
            refined = np.ascontiguousarray(what[n2o] > 0, np.int32)
            repair = np.ascontiguousarray(what[n2o] != 0, np.int32)
            u, b = self.move(cfg, A, what, n2o, o2n, lloc_old, lloc_new, refined)
            cfg = ocfg(mesh, p)
            A = O.Arrays(cfg)
            A.u0[:] = u[:new_nmb]
            for dst, src in zip(A.b0, b):
                dst[:] = src[:new_nmb]
            mbs, ngh, lev = self.tables()
            cu0, cbn = self.coarse(cfg, new_nmb)
            st = A.state()
            L.oracle_mhd_smr_ghosts_and_prims(C.byref(cfg), C.byref(st), O.p(cu0), *[O.p(x) for x in cbn], O.ip(ngh), O.ip(lev))
            self.repair(cfg, A, cbn, repair)                                                                             # RepairAMRFC
            L.oracle_mhd_smr_ghosts_and_prims(C.byref(cfg), C.byref(st), O.p(cu0), *[O.p(x) for x in cbn], O.ip(ngh), O.ip(lev))
        else:
            cu0, cbn = self.coarse(cfg, cfg.nmb)
        st = A.state()
        cfl, tlim = float(p["time/cfl_number"]), float(p["time/tlim"])
        dtnew = L.oracle_new_dt(C.byref(cfg), O.p(A.w0), O.p(A.bcc0), O.p(mbs), None)
        dt = min(2.0 * r.dt, cfl * dtnew)                                   # Mesh::NewTimeStep (mesh.cpp:573-640)
        if r.time < tlim and (r.time + dt) > tlim:
            dt = tlim - r.time
        t, d = C.c_double(r.time), C.c_double(dt)
        n = L.oracle_mhd_smr_run(C.byref(cfg), C.byref(st), O.p(cu0), *[O.p(x) for x in cbn], O.p(mbs), O.ip(ngh), O.ip(lev),
                                 H.INTEGRATOR[p["time/integrator"]], C.c_double(cfl), C.c_double(tlim), 1, C.byref(t), C.byref(d))
        assert n == 1
        return A, t.value, dt


class HydroStepper(Stepper):
    """The same harness for the hydro twin (reference: phydro branch of AdaptiveMeshRefinement, one ghost-zone rebuild, no
    RepairAMRFC).  The data movement is the subclass hook `move_cc`; default = the oracle's in-place restatement."""

    def __init__(self, p):
        super().__init__(p)
        L = self.L
        L.oracle_hydro_new_dt.restype = C.c_double
        L.oracle_hydro_smr_init.restype = C.c_double
        L.oracle_hydro_smr_run.restype = C.c_int

    def hcfg(self, nmb=None):
        p, mesh = self.p, self.mesh
        return O.OCfg(mesh.nmb_total if nmb is None else nmb, int(p["mesh/nghost"]), *mesh.nblock, O.RECON[p["hydro/reconstruct"]],
                      O.HYD_RSOLVER[p["hydro/rsolver"]], float(p["hydro/gamma"]), O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MAX, 0, 0, 1.0, 0)

    def move_cc(self, cfg, u0, what, n2o, o2n, lloc_old, lloc_new):
        L, mesh = self.L, self.mesh
        old_nmb, new_nmb = cfg.nmb, mesh.nmb_total
        cap = max(old_nmb, new_nmb)
        u = np.concatenate([u0, np.zeros((cap - old_nmb,) + u0.shape[1:])]) if cap > old_nmb else u0.copy()
        cu, cb = self.coarse(cfg, cap)
        fb = O.Arrays(ocfg(mesh, {**self.p, "mhd/reconstruct": "plm", "mhd/rsolver": "hlld", "mhd/gamma": "1.6"}, nmb=cap)).b0      # unused face fields
        L.oracle_smr_restrict_cc(C.byref(cfg), O.p(u), O.p(cu), 5)
        L.oracle_amr_redistribute(C.byref(cfg), new_nmb, O.p(u), O.p(cu), 5, *[O.p(x) for x in fb], *[O.p(x) for x in cb],
                                  O.ip(what), O.ip(n2o), O.ip(o2n), O.ip(lloc_old), O.ip(lloc_new))
        return u[:new_nmb]

    def advance(self, r, ncycle):
        p, L, mesh = self.p, self.L, self.mesh
        assert np.array_equal(mesh.lloc, r.lloc), f"block list differs at cycle {ncycle}"
        cfg = self.hcfg()
        u0 = np.ascontiguousarray(r.u0).copy()
        w0 = np.zeros_like(u0)
        n3, n2, n1 = u0.shape[2:]
        L.oracle_hydro_cons2prim(C.byref(cfg), O.p(u0), O.p(w0), 0, n1 - 1, 0, n2 - 1, 0, n3 - 1, None)
        mbs, ngh, lev = self.tables()
        flag = criteria_flags_hydro(L, cfg, p, w0, mbs)
        lloc_old = np.ascontiguousarray(mesh.lloc, np.int32).copy()
        nnew, ndel, what = mesh.adapt(ncycle, flag)
        cfl, tlim = float(p["time/cfl_number"]), float(p["time/tlim"])
        if nnew or ndel:
            self.nadapt += 1
            self.ncreated += nnew
            self.ndeleted += ndel
            n2o, o2n = np.ascontiguousarray(mesh.newtoold, np.int32), np.ascontiguousarray(mesh.oldtonew, np.int32)
            u0 = np.ascontiguousarray(self.move_cc(cfg, u0, what, n2o, o2n, lloc_old, np.ascontiguousarray(mesh.lloc, np.int32)))
            cfg = self.hcfg()
            w0 = np.zeros_like(u0)
            mbs, ngh, lev = self.tables()
            cu0 = np.zeros((cfg.nmb, 5, L.oracle_smr_coarse_cells(C.byref(cfg))))
            L.oracle_hydro_smr_init(C.byref(cfg), O.p(u0), O.p(cu0), O.p(w0), O.p(mbs), O.ip(ngh), O.ip(lev), C.c_double(cfl), C.c_double(tlim),
                                    C.c_double(r.time))          # ghost zones + primitives (its dt is not used)
        else:
            cu0 = np.zeros((cfg.nmb, 5, L.oracle_smr_coarse_cells(C.byref(cfg))))
        dtnew = L.oracle_hydro_new_dt(C.byref(cfg), O.p(w0), O.p(mbs), None)
        dt = min(2.0 * r.dt, cfl * dtnew)
        if r.time < tlim and (r.time + dt) > tlim:
            dt = tlim - r.time
        u1 = np.zeros_like(u0)
        t, d = C.c_double(r.time), C.c_double(dt)
        n = L.oracle_hydro_smr_run(C.byref(cfg), O.p(u0), O.p(u1), O.p(cu0), O.p(w0), O.p(mbs), O.ip(ngh), O.ip(lev),
                                   H.INTEGRATOR[p["time/integrator"]], C.c_double(cfl), C.c_double(tlim), 1, C.byref(t), C.byref(d))
        assert n == 1
        return u0, t.value, dt


def criteria_flags_hydro(L, cfg, p, w0, mbs):
    flag = np.zeros(cfg.nmb, np.int32)
    for blk in sorted({k.split("/")[0] for k in p if k.startswith("amr_criterion")}):
        loc = np.array([float(p.get(f"{blk}/location_x{d}", 0.0)) for d in (1, 2, 3)] + [float(p.get(f"{blk}/location_rad", 0.0))])
        assert p[f"{blk}/variable"] in ("hydro_w_d", "hydro_u_d")
        L.oracle_amr_criterion(C.byref(cfg), METHOD[p[f"{blk}/method"]], O.p(w0), w0.shape[1], 0,
                               C.c_double(float(p.get(f"{blk}/value_min", -O.FLT_MAX))), C.c_double(float(p.get(f"{blk}/value_max", O.FLT_MAX))),
                               O.p(loc), O.p(mbs), O.ip(flag))
    return flag

