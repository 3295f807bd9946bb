"""Reader/writer for the reference's restart (.rst) dump layout.

Layout (reference src/outputs/restart.cpp:216-294,312-469):
  input-deck text ending in "<par_end>\\n", int nmb_total, int root_level,
  RegionSize mesh_size (9 f64), RegionIndcs mesh_indcs (19 i32), RegionIndcs mb_indcs (19 i32),
  f64 time, f64 dt, int ncycle, LogicalLocation[nmb] (4 i32 each), float cost[nmb],
  u64 bytes-per-block, then per MeshBlock: u0[5][n3][n2][n1], b0.x1f[n3][n2][n1+1],
  b0.x2f[n3][n2+1][n1], b0.x3f[n3+1][n2][n1]  (ghost zones included).
This is the parity fixture format: tests compare the engine's u0/b0 with dumps of the
unmodified reference.
"""
from __future__ import annotations

import dataclasses
import struct

import numpy as np

_INDCS = ("ng", "nx1", "nx2", "nx3", "is", "ie", "js", "je", "ks", "ke",
          "cnx1", "cnx2", "cnx3", "cis", "cie", "cjs", "cje", "cks", "cke")


@dataclasses.dataclass
class Restart:
    deck: str
    nmb: int
    root_level: int
    mesh_size: np.ndarray      # 9 f64: x1min,x2min,x3min,x1max,x2max,x3max,dx1,dx2,dx3
    mesh_indcs: dict
    mb_indcs: dict
    time: float
    dt: float
    ncycle: int
    lloc: np.ndarray           # (nmb,4) i32: lx1,lx2,lx3,level
    cost: np.ndarray           # (nmb,) f32
    u0: np.ndarray             # (nmb,5,n3,n2,n1)
    b1f: np.ndarray            # (nmb,n3,n2,n1+1)
    b2f: np.ndarray            # (nmb,n3,n2+1,n1)
    b3f: np.ndarray            # (nmb,n3+1,n2,n1)


def read_rst(path: str, nmhd: int = 5, physics: str = "mhd") -> Restart:
    """physics="hydro": the dump of a hydro-only run holds just u0 per block (restart.cpp:320-353);
    b1f/b2f/b3f come back as None."""
    raw = open(path, "rb").read()
    end = raw.index(b"<par_end>\n") + len(b"<par_end>\n")
    deck = raw[:end].decode()
    off = end
    nmb, root_level = struct.unpack_from("<ii", raw, off); off += 8
    mesh_size = np.frombuffer(raw, "<f8", 9, off).copy(); off += 72
    mesh_indcs = dict(zip(_INDCS, struct.unpack_from("<19i", raw, off))); off += 76
    mb_indcs = dict(zip(_INDCS, struct.unpack_from("<19i", raw, off))); off += 76
    time, dt = struct.unpack_from("<dd", raw, off); off += 16
    (ncycle,) = struct.unpack_from("<i", raw, off); off += 4
    lloc = np.frombuffer(raw, "<i4", 4 * nmb, off).reshape(nmb, 4).copy(); off += 16 * nmb
    cost = np.frombuffer(raw, "<f4", nmb, off).copy(); off += 4 * nmb
    (per_block,) = struct.unpack_from("<Q", raw, off); off += 8
    ng = mb_indcs["ng"]
    n1 = mb_indcs["nx1"] + 2 * ng
    n2 = mb_indcs["nx2"] + 2 * ng if mb_indcs["nx2"] > 1 else 1
    n3 = mb_indcs["nx3"] + 2 * ng if mb_indcs["nx3"] > 1 else 1
    ncc, nf1, nf2, nf3 = n3 * n2 * n1, n3 * n2 * (n1 + 1), n3 * (n2 + 1) * n1, (n3 + 1) * n2 * n1
    if physics == "hydro":
        if per_block != 8 * nmhd * ncc:
            raise ValueError(f"{path}: block payload {per_block} B, expected {8 * nmhd * ncc} B (hydro only)")
        u0 = np.empty((nmb, nmhd, n3, n2, n1))
        for m in range(nmb):
            u0[m] = np.frombuffer(raw, "<f8", nmhd * ncc, off).reshape(nmhd, n3, n2, n1); off += 8 * nmhd * ncc
        if off != len(raw):
            raise ValueError(f"{path}: {len(raw) - off} trailing bytes")
        return Restart(deck, nmb, root_level, mesh_size, mesh_indcs, mb_indcs, time, dt, ncycle,
                       lloc, cost, u0, None, None, None)
    expect = 8 * (nmhd * ncc + nf1 + nf2 + nf3)
    if per_block != expect:
        raise ValueError(f"{path}: block payload {per_block} B, expected {expect} B (MHD only)")
    # per block (restart.cpp:312-469): u0, then b0.x1f, b0.x2f, b0.x3f
    u0 = np.empty((nmb, nmhd, n3, n2, n1)); b1f = np.empty((nmb, n3, n2, n1 + 1))
    b2f = np.empty((nmb, n3, n2 + 1, n1)); b3f = np.empty((nmb, n3 + 1, n2, n1))
    for m in range(nmb):
        u0[m] = np.frombuffer(raw, "<f8", nmhd * ncc, off).reshape(nmhd, n3, n2, n1); off += 8 * nmhd * ncc
        b1f[m] = np.frombuffer(raw, "<f8", nf1, off).reshape(n3, n2, n1 + 1); off += 8 * nf1
        b2f[m] = np.frombuffer(raw, "<f8", nf2, off).reshape(n3, n2 + 1, n1); off += 8 * nf2
        b3f[m] = np.frombuffer(raw, "<f8", nf3, off).reshape(n3 + 1, n2, n1); off += 8 * nf3
    if off != len(raw):
        raise ValueError(f"{path}: {len(raw) - off} trailing bytes")
    return Restart(deck, nmb, root_level, mesh_size, mesh_indcs, mb_indcs, time, dt, ncycle,
                   lloc, cost, u0, b1f, b2f, b3f)


def write_rst(path: str, r: Restart) -> None:
    """Writes a dump the reference can restart from (`athena -r file`): same byte layout as
    src/outputs/restart.cpp:216-294,312-469 (single file; MHD, or hydro when r.b1f is None)."""
    nmb = int(r.nmb)
    with open(path, "wb") as f:
        f.write(r.deck.encode())
        f.write(struct.pack("<ii", nmb, int(r.root_level)))
        f.write(np.ascontiguousarray(r.mesh_size, "<f8").tobytes())
        f.write(struct.pack("<19i", *[int(r.mesh_indcs[k]) for k in _INDCS]))
        f.write(struct.pack("<19i", *[int(r.mb_indcs[k]) for k in _INDCS]))
        f.write(struct.pack("<dd", float(r.time), float(r.dt)))
        f.write(struct.pack("<i", int(r.ncycle)))
        f.write(np.ascontiguousarray(r.lloc, "<i4").tobytes())
        f.write(np.ascontiguousarray(r.cost, "<f4").tobytes())
        arrays = [a for a in (r.u0, r.b1f, r.b2f, r.b3f) if a is not None]
        per_block = 8 * sum(a[0].size for a in arrays)
        f.write(struct.pack("<Q", per_block))
        for m in range(nmb):
            for a in arrays:
                f.write(np.ascontiguousarray(a[m], "<f8").tobytes())
