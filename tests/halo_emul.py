"""numpy interpreter of the halo copy descriptors exported by ak_halo_descs (TEST ONLY): the same
semantics as halo_copy_kernel in athenak_b200/csrc/halo.cu, used to check the exchange plan and
the per-peer message layout on CPU (single process and world_size-2 gloo)."""
import ctypes as C

import numpy as np

from athenak_b200 import lib as _lib


def make_plan(cfg, nghbr, rank, gids_start):
    L = _lib.load()
    plan = C.c_void_p()
    nghbr = np.ascontiguousarray(nghbr, np.int32)
    _lib.check(L.ak_halo_create(C.byref(cfg), nghbr.ctypes.data_as(C.POINTER(C.c_int)), rank, gids_start, C.byref(plan)))
    return plan


def descs(plan, kind, ipeer=0):
    L = _lib.load()
    n = L.ak_halo_descs(plan, kind, ipeer, None, 0)
    assert n >= 0
    out = np.zeros((max(n, 1), 13), np.int64)
    L.ak_halo_descs(plan, kind, ipeer, out.ctypes.data_as(C.POINTER(C.c_long)), n)
    return out[:n]


def peers(plan):
    L = _lib.load()
    res = []
    for ip in range(L.ak_halo_npeers(plan)):
        r, a, b = C.c_int(), C.c_long(), C.c_long()
        _lib.check(L.ak_halo_peer(plan, ip, C.byref(r), C.byref(a), C.byref(b)))
        res.append((r.value, a.value, b.value))
    return res


def _box(d):
    lo, hi = d[3:6], d[6:9]
    return (slice(lo[2], hi[2] + 1), slice(lo[1], hi[1] + 1), slice(lo[0], hi[0] + 1)), tuple(int(h - l + 1) for l, h in zip(lo, hi))


def local(ds, arrays_by_comp):
    """arrays_by_comp: {-1: u (nmb,nv,n3,n2,n1)} or {0: x1f, 1: x2f, 2: x3f} ((nmb,k,j,i))."""
    for d in ds:
        m, sm, comp = int(d[0]), int(d[1]), int(d[2])
        (sk, sj, si), _ = _box(d)
        sh = d[9:12]
        src = (slice(sk.start + sh[2], sk.stop + sh[2]), slice(sj.start + sh[1], sj.stop + sh[1]), slice(si.start + sh[0], si.stop + sh[0]))
        a = arrays_by_comp[comp]
        if comp < 0:
            a[m][(slice(None),) + (sk, sj, si)] = a[sm][(slice(None),) + src]
        else:
            a[m][(sk, sj, si)] = a[sm][src]


def pack(ds, arrays_by_comp, nvar, buf):
    for d in ds:
        m, comp, boff = int(d[0]), int(d[2]), int(d[12])
        (sk, sj, si), (ni, nj, nk) = _box(d)
        per = ni * nj * nk
        a = arrays_by_comp[comp]
        if comp < 0:
            buf[boff * nvar: boff * nvar + nvar * per] = a[m][(slice(None), sk, sj, si)].reshape(-1)
        else:
            buf[boff: boff + per] = a[m][(sk, sj, si)].reshape(-1)


def unpack(ds, arrays_by_comp, nvar, buf):
    for d in ds:
        m, comp, boff = int(d[0]), int(d[2]), int(d[12])
        (sk, sj, si), (ni, nj, nk) = _box(d)
        per = ni * nj * nk
        a = arrays_by_comp[comp]
        if comp < 0:
            a[m][(slice(None), sk, sj, si)] = buf[boff * nvar: boff * nvar + nvar * per].reshape(nvar, nk, nj, ni)
        else:
            a[m][(sk, sj, si)] = buf[boff: boff + per].reshape(nk, nj, ni)
