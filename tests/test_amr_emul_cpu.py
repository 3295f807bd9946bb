"""AMR device path, host side (dynamic refinement): the per-element bodies of the AMR kernels
(csrc/amr_body.h -- out-of-place gathers: copy / on-the-fly restriction of the old siblings / direct prolongation from the
old parent, RefineFC, RepairAMRFC, criterion extrema; compiled here with g++ by tests/emul/amr_emul.cpp) and the host plan
(ak_amr_plan) replace the oracle's data movement inside the cycle-by-cycle harness of tests/amr_util.py: from the
reference's state of cycle k they must reproduce the reference's state of cycle k+1 BITWISE (ghost zones included), i.e.
the gather formulation equals the reference's in-place move + coarse-array staging, and the zero-initialised ghost zones
of changed blocks never reach a result.  No GPU needed; tests/test_gpu_zamr.py runs the same bodies as CUDA kernels."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import amr_util as AU
import golden_io as G
import hydro_util as H
import oracle_lib as O
from athenak_b200 import lib as _lib
from test_amr_data_cpu import CASES, check_step

HERE = os.path.dirname(os.path.abspath(__file__))
vp = C.c_void_p


@pytest.fixture(scope="module")
def emul():
    out = os.path.join(HERE, "emul", "_amr_emul.so")
    src = os.path.join(HERE, "emul", "amr_emul.cpp")
    hdrs = [os.path.join(HERE, "..", "athenak_b200", "csrc", h) for h in ("amr_body.h", "smr_fc_body.h")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", out, src])
    return C.CDLL(out)


def P(a):
    return a.ctypes.data_as(vp)


class EmulStepper(AU.Stepper):
    def __init__(self, p, emul):
        super().__init__(p)
        self.E, self.K = emul, _lib.load()

    def engine_cfg(self, nmb):
        ng = int(self.p["mesh/nghost"])
        return _lib.MhdCfg(nmb, ng, *self.mesh.nblock, 1, 2, 5.0 / 3.0, O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MAX, 0, 0, 1.0, 0)

    def fcdims(self, nmb):
        """FcDims of the block shape (from an SMR plan of the current mesh)."""
        _, ngh, lev = self.tables()
        cfg = self.engine_cfg(nmb)
        plan, ipp = vp(), C.POINTER(C.c_int)
        assert self.K.ak_smr_create(C.byref(cfg), ngh.ctypes.data_as(ipp), lev.ctypes.data_as(ipp), 0, 0, C.byref(plan)) == 0
        ptr, cnt = vp(), C.c_long()
        assert self.K.ak_smr_fc_tables(plan, 0, C.byref(ptr), C.byref(cnt)) == 0
        raw = C.string_at(ptr.value, 512)          # FcDims is a POD of < 512 bytes; the harness reads only its own fields
        self.K.ak_smr_destroy(plan)
        return C.create_string_buffer(raw, 512)

    def move(self, cfg, A, what, n2o, o2n, lloc_old, lloc_new, refined):
        new_nmb = self.mesh.nmb_total
        ecfg = self.engine_cfg(new_nmb)
        plan = np.zeros((new_nmb, 5), np.int32)
        ipp = C.POINTER(C.c_int)
        assert self.K.ak_amr_plan(C.byref(ecfg), cfg.nmb, n2o.ctypes.data_as(ipp), what.ctypes.data_as(ipp), lloc_new.ctypes.data_as(ipp),
                                  plan.ctypes.data_as(ipp)) == 0
        D = self.fcdims(new_nmb)
        un = np.zeros((new_nmb,) + A.u0.shape[1:])
        bn = [np.zeros((new_nmb,) + x.shape[1:]) for x in A.b0]
        self.E.emul_amr_move(D, new_nmb, P(plan), P(A.u0), P(un), 5, *[P(x) for x in A.b0], *[P(x) for x in bn])
        self.E.emul_amr_interior(D, new_nmb, P(refined), *[P(x) for x in bn])
        return un, bn

    def repair(self, cfg, A, cbn, repair):
        self.E.emul_amr_interior(self.fcdims(cfg.nmb), cfg.nmb, P(repair), *[P(x) for x in A.b0])

    def flags(self, cfg, A, mbs):
        """ak_amr_criterion's host rules on top of the emulated block extrema; must equal the oracle's verdict."""
        p = self.p
        flag = np.zeros(cfg.nmb, np.int32)
        D = self.fcdims(cfg.nmb)
        for blk in sorted({k.split("/")[0] for k in p if k.startswith("amr_criterion")}):
            method = AU.METHOD[p[f"{blk}/method"]]
            vmax, vmin = float(p.get(f"{blk}/value_max", O.FLT_MAX)), float(p.get(f"{blk}/value_min", -O.FLT_MAX))
            out = np.zeros((cfg.nmb, 2))
            self.E.emul_amr_extrema(D, cfg.nmb, method, 5, 0, P(A.w0), P(out))
            for m in range(cfg.nmb):
                if vmax < O.FLT_MAX:
                    if out[m, 0] > vmax:
                        flag[m] = 1
                    if out[m, 0] < vmax and flag[m] == 0:
                        flag[m] = -1
                if method == 0 and vmin > -O.FLT_MAX:
                    if out[m, 1] < vmin:
                        flag[m] = 1
                    if out[m, 1] > vmin and flag[m] == 0:
                        flag[m] = -1
        assert np.array_equal(flag, super().flags(cfg, A, mbs))
        return flag


def test_amr_emulated_device_path_matches_fixture(emul):
    d = G.load("amr_blast2d_events")
    p = H.deck_params(str(d["deck"]))
    S = EmulStepper(p, emul)

    class R:
        pass

    def state(k):
        r = R()
        r.lloc, r.u0, r.b1f, r.b2f, r.b3f = d[f"lloc_{k}"], d[f"u0_{k}"], d[f"b1f_{k}"], d[f"b2f_{k}"], d[f"b3f_{k}"]
        r.time, r.dt = float(d[f"time_{k}"]), float(d[f"dt_{k}"])
        return r

    events = [int(k) for k in d["events"]]
    for k in range(1, int(d["ncycles"]) - 1):
        if k in events:
            check_step(S, state(k), state(k + 1), k)
        else:
            cfg = AU.ocfg(S.mesh, p)
            flag = AU.criteria_flags(S.L, cfg, p, np.ascontiguousarray(d[f"dens_{k}"][:, None]), S.tables()[0])
            S.mesh.adapt(k, flag)
        assert np.array_equal(S.mesh.lloc, d[f"lloc_{k + 1}"])
    assert S.ncreated > 0 and S.ndeleted > 0


@pytest.mark.parametrize("name", ["2d_3levels", "3d", "3d_deref"])
def test_amr_emulated_device_path_matches_live_reference(name, emul):
    if not O.have_reference("athena_blast"):
        pytest.skip("oracle/_ref/athena_blast not built")
    ov = CASES[name]
    dumps = AU.reference_dumps(ov)
    S = EmulStepper(AU.params(ov), emul)
    for k in range(1, len(dumps) - 1):
        check_step(S, dumps[k], dumps[k + 1], k)
    assert S.nadapt >= 1 and S.ncreated > 0


class EmulHydroStepper(AU.HydroStepper):
    """Hydro: the cell-centred gather alone (face arrays are dummies)."""

    def __init__(self, p, emul):
        super().__init__(p)
        self.helper = EmulStepper.__new__(EmulStepper)
        self.helper.__dict__.update(p=p, mesh=self.mesh, L=self.L, E=emul, K=_lib.load())

    def move_cc(self, cfg, u0, what, n2o, o2n, lloc_old, lloc_new):
        h = self.helper
        new_nmb = self.mesh.nmb_total
        ecfg = h.engine_cfg(new_nmb)
        plan = np.zeros((new_nmb, 5), np.int32)
        ipp = C.POINTER(C.c_int)
        assert h.K.ak_amr_plan(C.byref(ecfg), cfg.nmb, n2o.ctypes.data_as(ipp), what.ctypes.data_as(ipp), lloc_new.ctypes.data_as(ipp),
                               plan.ctypes.data_as(ipp)) == 0
        D = h.fcdims(new_nmb)
        n3, n2, n1 = u0.shape[2:]
        fo = [np.zeros((cfg.nmb, n3, n2, n1 + 1)), np.zeros((cfg.nmb, n3, n2 + 1, n1)), np.zeros((cfg.nmb, n3 + 1, n2, n1))]
        fn = [np.zeros((new_nmb,) + x.shape[1:]) for x in fo]
        un = np.zeros((new_nmb,) + u0.shape[1:])
        h.E.emul_amr_move(D, new_nmb, P(plan), P(u0), P(un), 5, *[P(x) for x in fo], *[P(x) for x in fn])
        return un


def test_hydro_amr_emulated_device_path_matches_live_reference(emul):
    if not O.have_reference("athena_blast"):
        pytest.skip("oracle/_ref/athena_blast not built")
    dumps = AU.reference_dumps(deck=AU.HYDRO_DECK, physics="hydro")
    S = EmulHydroStepper(AU.params(deck=AU.HYDRO_DECK), emul)
    for k in range(1, len(dumps) - 1):
        r, rn = dumps[k], dumps[k + 1]
        u, t, dt = S.advance(r, k)
        assert np.array_equal(S.mesh.lloc, rn.lloc) and t == rn.time and dt == rn.dt, f"cycle {k}"
        assert np.array_equal(u, rn.u0), f"cycle {k}: max|dU| = {np.abs(u - rn.u0).max()}"
    assert S.nadapt >= 2 and S.ncreated > 0 and S.ndeleted > 0

