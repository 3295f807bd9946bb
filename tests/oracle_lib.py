"""ctypes binding of oracle/liboracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

RECON = {"dc": 0, "plm": 1, "ppm4": 2, "ppmx": 3, "wenoz": 4, "teno": 5}
RSOLVER = {"llf": 0, "hlle": 1, "hlld": 2}
HYD_RSOLVER = {"advect": 10, "llf": 11, "hlle": 12, "hllc": 13, "roe": 14}
FLT_MIN = float(np.finfo(np.float32).tiny)
FLT_MAX = float(np.finfo(np.float32).max)


class OCfg(C.Structure):
    _fields_ = [("nmb", C.c_int), ("ng", C.c_int), ("nx1", C.c_int), ("nx2", C.c_int),
                ("nx3", C.c_int), ("recon", C.c_int), ("rsolver", C.c_int),
                ("gamma", C.c_double), ("dfloor", C.c_double), ("pfloor", C.c_double),
                ("tfloor", C.c_double), ("sfloor", C.c_double), ("sigma_max", C.c_double), ("nscalars", C.c_int),
                ("is_iso", C.c_int), ("iso_cs", C.c_double), ("fofc", C.c_int)]


class OState(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_double)) for n in
                ("u0", "u1", "w0", "bcc0", "b0x1f", "b0x2f", "b0x3f", "b1x1f", "b1x2f", "b1x3f")]


def make_cfg(nmb, ng, nx, recon="ppm4", rsolver="hlld", gamma=1.6666667, dfloor=FLT_MIN,
             pfloor=FLT_MIN, tfloor=FLT_MIN, sfloor=FLT_MIN, sigma_max=FLT_MAX, nscalars=0):
    nx1, nx2, nx3 = nx
    return OCfg(nmb, ng, nx1, nx2, nx3, RECON[recon], RSOLVER[rsolver], gamma, dfloor, pfloor,
                tfloor, sfloor, sigma_max, nscalars)


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        srcs = [os.path.join(ORACLE_DIR, f) for f in ("mhd_oracle.c", "hydro_oracle.c")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle.so"])
        _lib = C.CDLL(so)
        _lib.oracle_new_dt.restype = C.c_double
        _lib.oracle_initial_dt.restype = C.c_double
        _lib.oracle_run.restype = C.c_int
        _lib.oracle_hydro_new_dt.restype = C.c_double
        _lib.oracle_hydro_init.restype = C.c_double
        _lib.oracle_hydro_run.restype = C.c_int
    return _lib


def p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def ip(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Arrays:
    """All arrays of one MeshBlockPack in the reference layout (host numpy)."""

    def __init__(self, cfg: OCfg):
        nmb, ng = cfg.nmb, cfg.ng
        # inactive dimensions (nx = 1): extent 1, no ghost cells (mesh.cpp:285-333)
        n1, n2, n3 = cfg.nx1 + 2 * ng, (cfg.nx2 + 2 * ng if cfg.nx2 > 1 else 1), (cfg.nx3 + 2 * ng if cfg.nx3 > 1 else 1)
        self.cfg, self.n = cfg, (n1, n2, n3)
        z = np.zeros
        nv = (4 if cfg.is_iso else 5) + cfg.nscalars
        self.u0, self.u1, self.w0 = (z((nmb, nv, n3, n2, n1)) for _ in range(3))
        self.bcc0 = z((nmb, 3, n3, n2, n1))
        self.b0 = [z((nmb, n3, n2, n1 + 1)), z((nmb, n3, n2 + 1, n1)), z((nmb, n3 + 1, n2, n1))]
        self.b1 = [np.zeros_like(a) for a in self.b0]
        self.flx = [z((nmb, nv, n3, n2, n1 + 1)), z((nmb, nv, n3, n2 + 1, n1)), z((nmb, nv, n3 + 1, n2, n1))]
        # face EMFs in the order e3x1,e2x1,e1x2,e3x2,e2x3,e1x3 (mhd.cpp:349-354)
        self.fe = [z((nmb, n3, n2, n1)) for _ in range(6)]
        self.e = [z((nmb, n3 + 1, n2 + 1, n1)), z((nmb, n3 + 1, n2, n1 + 1)), z((nmb, n3, n2 + 1, n1 + 1))]

    def state(self) -> OState:
        return OState(p(self.u0), p(self.u1), p(self.w0), p(self.bcc0), *[p(a) for a in self.b0],
                      *[p(a) for a in self.b1])

    def c2p_all(self):
        n1, n2, n3 = self.n
        nf = (C.c_int * 3)()
        lib().oracle_cons2prim(C.byref(self.cfg), p(self.u0), *[p(a) for a in self.b0], p(self.w0),
                               p(self.bcc0), 0, n1 - 1, 0, n2 - 1, 0, n3 - 1, nf)
        return list(nf)


def uniform_mesh(nmesh, nblock, xmin=(-0.5, -0.5, -0.5), xmax=(0.5, 0.5, 0.5), periodic=True):
    """Z-ordered block list + neighbour gids + mb_size for a uniform mesh, written directly
    from the definition (meshblock_tree.cpp:60-90,336-350; meshblock.cpp:25-131,142-430).
    Independent of the product's host mesh code, which it is used to check."""
    nmbr = [nmesh[d] // nblock[d] for d in range(3)]
    root_level = 0
    while (1 << root_level) < max(nmbr):
        root_level += 1
    llocs = []

    def walk(lx, ly, lz, lev):
        if lev == root_level:
            llocs.append((lx, ly, lz))
            return
        levfac = 1 << (root_level - lev - 1)
        for n in range(8):
            i, j, k = n & 1, (n >> 1) & 1, (n >> 2) & 1
            if ((lz * 2 + k) * levfac < nmbr[2] and (ly * 2 + j) * levfac < nmbr[1]
                    and (lx * 2 + i) * levfac < nmbr[0]):
                walk(lx * 2 + i, ly * 2 + j, lz * 2 + k, lev + 1)

    walk(0, 0, 0, 0)
    gid_of = {l: g for g, l in enumerate(llocs)}
    nmb = len(llocs)
    nghbr = -np.ones((nmb, 56), np.int32)
    L = lib()
    for g, (lx, ly, lz) in enumerate(llocs):
        for oz in (-1, 0, 1):
            for oy in (-1, 0, 1):
                for ox in (-1, 0, 1):
                    if ox == oy == oz == 0:
                        continue
                    if (oy and nmesh[1] == 1) or (oz and nmesh[2] == 1):
                        continue      # meshblock.cpp:170-425: neighbours only in active dimensions
                    q = [lx + ox, ly + oy, lz + oz]
                    ok = True
                    for d in range(3):
                        if q[d] < 0 or q[d] >= nmbr[d]:
                            if (periodic[d] if isinstance(periodic, (tuple, list)) else periodic):
                                q[d] %= nmbr[d]
                            else:
                                ok = False
                    if ok:
                        nghbr[g, L.oracle_neighbor_index(ox, oy, oz, 0, 0)] = gid_of[tuple(q)]
    mb_size = np.zeros((nmb, 9))

    for g, l in enumerate(llocs):
        for d in range(3):
            n = nmbr[d]
            lo, hi = np.float64(xmin[d]), np.float64(xmax[d])
            bmin = lo if l[d] == 0 else _left_edge(l[d], n, lo, hi)
            bmax = hi if l[d] == n - 1 else _left_edge(l[d] + 1, n, lo, hi)
            mb_size[g, d], mb_size[g, 3 + d] = bmin, bmax
            mb_size[g, 6 + d] = (bmax - bmin) / np.float64(nblock[d])
    return np.array(llocs, np.int32), nghbr, mb_size, root_level


BC = {"block": 0, "reflect": 1, "inflow": 2, "outflow": 3, "diode": 4, "user": 5, "periodic": 6,
      "shear_periodic": 7, "vacuum": 8}          # BoundaryFlag, bvals.hpp:19


def block_bcs(lloc, nmesh, nblock, mesh_bcs):
    """mb_bcs[nmb][6] (meshblock.cpp:25-131): the mesh flag on faces lying on the mesh edge,
    BoundaryFlag::block elsewhere.  mesh_bcs = 6 flags ix1,ox1,ix2,ox2,ix3,ox3."""
    nmbr = [nmesh[d] // nblock[d] for d in range(3)]
    out = np.zeros((len(lloc), 6), np.int32)
    for g, l in enumerate(lloc):
        for d in range(3):
            out[g, 2 * d] = mesh_bcs[2 * d] if l[d] == 0 else BC["block"]
            out[g, 2 * d + 1] = mesh_bcs[2 * d + 1] if l[d] == nmbr[d] - 1 else BC["block"]
    return out


def _left_edge(ith, n, xmin, xmax):
    """LeftEdgeX, coordinates/cell_locations.hpp:24-39."""
    x = np.float64(ith) / np.float64(n)
    return (x * xmax - x * xmin) - (0.5 * xmax - 0.5 * xmin) + (0.5 * xmin + 0.5 * xmax)


def run_reference(binary, deck, workdir, overrides=(), threads=1, timeout=3600):
    """Run an oracle/_ref binary; returns stdout."""
    exe = os.path.join(REF_DIR, binary)
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="false")
    os.makedirs(workdir, exist_ok=True)
    out = subprocess.run([exe, "-i", deck, *overrides], cwd=workdir, env=env, timeout=timeout,
                         capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError(f"reference run failed:\n{out.stdout[-2000:]}\n{out.stderr[-2000:]}")
    return out.stdout


def have_reference(binary="athena_blast"):
    return os.path.exists(os.path.join(REF_DIR, binary))
