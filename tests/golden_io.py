"""Loads the committed golden fixtures (tests/golden/*.npz, made by tests/gen_golden.py from the
unmodified reference) into full ghost-inclusive arrays.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

import oracle_lib as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def fill_active(A: O.Arrays, d, cyc, block=0):
    """Put the active zones of golden state `cyc` into block `block` of A (ghosts untouched)."""
    ng = A.cfg.ng
    s = slice(ng, -ng)
    A.u0[block, :, s, s, s] = d[f"u0_c{cyc}"][0]
    A.b0[0][block, s, s, ng:-ng] = d[f"b1f_c{cyc}"][0]
    A.b0[1][block, s, ng:-ng, s] = d[f"b2f_c{cyc}"][0]
    A.b0[2][block, ng:-ng, s, s] = d[f"b3f_c{cyc}"][0]


def active_of(A: O.Arrays, block=0):
    ng = A.cfg.ng
    s = slice(ng, -ng)
    return (A.u0[block, :, s, s, s], A.b0[0][block, s, s, ng:-ng], A.b0[1][block, s, ng:-ng, s],
            A.b0[2][block, ng:-ng, s, s])


def oracle_init(A: O.Arrays, nghbr_gid):
    """Driver::Initialize on the oracle: ghost exchange of U and B, c2p everywhere."""
    L = O.lib()
    L.oracle_halo_cc(C.byref(A.cfg), O.p(A.u0), 5, O.ip(nghbr_gid))
    L.oracle_halo_fc(C.byref(A.cfg), *[O.p(a) for a in A.b0], O.ip(nghbr_gid))
    return A.c2p_all()


def oracle_init_bcs(A: O.Arrays, nghbr_gid, mesh_bcs, mb_bcs):
    """Driver::Initialize with physical boundaries: ghost exchange, ApplyPhysicalBCs, c2p."""
    L = O.lib()
    L.oracle_halo_cc(C.byref(A.cfg), O.p(A.u0), 5, O.ip(nghbr_gid))
    L.oracle_halo_fc(C.byref(A.cfg), *[O.p(a) for a in A.b0], O.ip(nghbr_gid))
    assert L.oracle_apply_bcs(C.byref(A.cfg), O.ip(mesh_bcs), O.ip(mb_bcs), O.p(A.u0), *[O.p(a) for a in A.b0]) == 0
    return A.c2p_all()
