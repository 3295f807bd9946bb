"""Face-centred SMR path, host side (MHD): the tables ak_smr_create builds (winner maps of the
gather-form SendB/RecvB, fill / shared-face / interior boxes with the ownership rule, EMF pack boxes and the
per-block sum / zero / average op lists) and the per-element kernel bodies (csrc/smr_fc_body.h, compiled here with
g++ by tests/emul/smr_fc_emul.cpp) against oracle/smr_fc_oracle.c on RANDOM data: every array element, ghost and
coarse included, holds different bits, so a wrong owner, order or box shows up; walking the elements backwards must
give the same bits (the device launches have no defined order).  No GPU needed; the GPU tests run the same bodies
as CUDA kernels."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import smr_util as S
from athenak_b200 import lib as _lib

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = dict(S.MHD_LIVE_CASES)
vp = C.c_void_p


@pytest.fixture(scope="module")
def emul():
    out = os.path.join(HERE, "emul", "_smr_fc_emul.so")
    src = os.path.join(HERE, "emul", "smr_fc_emul.cpp")
    hdr = os.path.join(HERE, "..", "athenak_b200", "csrc", "smr_fc_body.h")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", out, src])
    return C.CDLL(out)


class Setup:
    def __init__(self, name):
        self.mesh, p = S.mesh_of(open(S.MHD_DECK).read(), overrides=CASES[name])
        mesh = self.mesh
        _, _, ngh = mesh.rank_tables(0)
        self.ngh = np.ascontiguousarray(ngh, np.int32)
        self.lev = np.ascontiguousarray(mesh.lloc[:, 3], np.int32)
        ng = int(p["mesh/nghost"])
        self.ocfg = O.OCfg(mesh.nmb_total, ng, *mesh.nblock, O.RECON["plm"], O.RSOLVER["hlld"], 5.0 / 3.0, O.FLT_MIN, O.FLT_MIN,
                           O.FLT_MIN, O.FLT_MIN, O.FLT_MAX, 0, 0, 1.0, 0)
        self.L = _lib.load()
        cfg = _lib.MhdCfg(mesh.nmb_total, ng, *mesh.nblock, 1, 2, 5.0 / 3.0, O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MAX, 0, 0, 1.0, 0)
        self.plan = vp()
        ipp = C.POINTER(C.c_int)
        rc = self.L.ak_smr_create(C.byref(cfg), self.ngh.ctypes.data_as(ipp), self.lev.ctypes.data_as(ipp), 0, 0, C.byref(self.plan))
        assert rc == 0, self.L.ak_smr_last_error(None).decode()
        self.A = O.Arrays(self.ocfg)
        OL = O.lib()
        OL.oracle_smr_coarse_faces.restype = C.c_size_t
        self.ncf = [OL.oracle_smr_coarse_faces(C.byref(self.ocfg), v) for v in range(3)]
        assert self.ncf == [self.L.ak_smr_coarse_faces(self.plan, v) for v in range(3)]

    def table(self, which):
        ptr, cnt = vp(), C.c_long()
        assert self.L.ak_smr_fc_tables(self.plan, which, C.byref(ptr), C.byref(cnt)) == 0
        return ptr, cnt.value

    def random_fields(self, seed):
        rng = np.random.default_rng(seed)
        b = [rng.standard_normal(a.shape) for a in self.A.b0]
        cb = [rng.standard_normal((self.mesh.nmb_total, n)) for n in self.ncf]
        return b, cb

    def __del__(self):
        try:
            self.L.ak_smr_destroy(self.plan)
        except Exception:
            pass


_SETUPS = {}


def setup_of(name):
    if name not in _SETUPS:
        _SETUPS[name] = Setup(name)
    return _SETUPS[name]


def P(a):
    return a.ctypes.data_as(vp)


@pytest.mark.parametrize("name", sorted(CASES))
def test_restrict_exchange_fc(name, emul):
    s = setup_of(name)
    OL = O.lib()
    b, cb = s.random_fields(1)
    ob, ocb = [x.copy() for x in b], [x.copy() for x in cb]
    OL.oracle_smr_restrict_fc(C.byref(s.ocfg), *[O.p(x) for x in ob], *[O.p(x) for x in ocb])
    OL.oracle_smr_halo_fc(C.byref(s.ocfg), *[O.p(x) for x in ob], *[O.p(x) for x in ocb], O.ip(s.ngh), O.ip(s.lev))
    D = s.table(0)[0]
    maps = (vp * 6)(*[s.table(w)[0] for w in range(1, 7)])
    for rev in (0, 1):
        eb, ecb = [x.copy() for x in b], [x.copy() for x in cb]
        emul.emul_restrict_fc(D, *[P(x) for x in eb], *[P(x) for x in ecb])
        emul.emul_exchange_fc(D, maps, s.table(7)[0], *[P(x) for x in eb], *[P(x) for x in ecb], rev)
        for v in range(3):
            assert np.array_equal(eb[v], ob[v]), f"fine x{v + 1}f differs at {np.argwhere(eb[v] != ob[v])[:5]}"
            assert np.array_equal(ecb[v], ocb[v]), f"coarse x{v + 1}f differs ({(ecb[v] != ocb[v]).sum()} faces)"
    assert any((x != y).any() for x, y in zip(ob, b))


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("steps", [1, 2, 3])
def test_prolongate_fc(name, steps, emul):
    s = setup_of(name)
    OL = O.lib()
    b, cb = s.random_fields(2)
    ob, ocb = [x.copy() for x in b], [x.copy() for x in cb]
    OL.oracle_smr_prolongate_fc_steps(C.byref(s.ocfg), *[O.p(x) for x in ob], *[O.p(x) for x in ocb], O.ip(s.ngh), O.ip(s.lev), steps)
    D = s.table(0)[0]
    (fill, nfill), (sh, nsh), (it, nit) = s.table(8), s.table(9), s.table(10)
    for rev in (0, 1):
        eb, ecb = [x.copy() for x in b], [x.copy() for x in cb]
        emul.emul_prolongate_fc(D, fill, C.c_long(nfill), sh, C.c_long(nsh), it, C.c_long(nit), steps, *[P(x) for x in eb], *[P(x) for x in ecb], rev)
        for v in range(3):
            assert np.array_equal(eb[v], ob[v]), f"fine x{v + 1}f differs at {np.argwhere(eb[v] != ob[v])[:5]}"
            assert np.array_equal(ecb[v], ocb[v]), f"coarse x{v + 1}f differs"
    if steps & 2:
        assert any((x != y).any() for x, y in zip(ob, b))


@pytest.mark.parametrize("name", sorted(CASES))
def test_emf_correct(name, emul):
    s = setup_of(name)
    OL = O.lib()
    rng = np.random.default_rng(3)
    e = [rng.standard_normal(a.shape) for a in s.A.e]
    oe = [x.copy() for x in e]
    OL.oracle_smr_emf_correct(C.byref(s.ocfg), *[O.p(x) for x in oe], O.ip(s.ngh), O.ip(s.lev))
    D = s.table(0)[0]
    (pk, npk), (ops, _), (rg, _), (_, nbuf) = s.table(11), s.table(12), s.table(13), s.table(14)
    for rev in (0, 1):
        ee = [x.copy() for x in e]
        buf = np.full(max(nbuf, 1), np.nan)
        emul.emul_emf_correct(D, pk, C.c_long(npk), ops, rg, P(buf), *[P(x) for x in ee], rev)
        for v in range(3):
            assert np.array_equal(ee[v], oe[v]), f"e{v + 1} differs at {np.argwhere(ee[v] != oe[v])[:5]}"
    assert any((x != y).any() for x, y in zip(oe, e))
