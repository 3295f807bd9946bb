"""Python handle on the C++ host mesh (athenak_b200/csrc/host_mesh.cpp): the uniform-mesh part
of the reference's Mesh / MeshBlockTree / MeshBlock metadata (src/mesh), kept on the host."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _lib

BC = _lib.BC          # BoundaryFlag values (src/bvals/bvals.hpp:19)


class Mesh:
    def __init__(self, nmesh, nblock, xmin=(-0.5, -0.5, -0.5), xmax=(0.5, 0.5, 0.5), bcs=("periodic",) * 6, nranks=1, refined_regions=None):
        """bcs: ix1,ox1,ix2,ox2,ix3,ox3 as BoundaryFlag names or values.
        refined_regions: None (uniform mesh) or a list of (level, x1min, x1max, x2min, x2max, x3min, x3max), the
        <refined_region> blocks of a static-mesh-refinement deck (host maps only, see ak_mesh_create_smr)."""
        self.L = _lib.load()
        bcs = tuple(BC[b] if isinstance(b, str) else int(b) for b in bcs)
        self.bcs = bcs
        self.nmesh, self.nblock, self.nranks = tuple(nmesh), tuple(nblock), int(nranks)
        self.xmin, self.xmax = tuple(xmin), tuple(xmax)
        i3, d3, i6 = (C.c_int * 3), (C.c_double * 3), (C.c_int * 6)
        self.h = C.c_void_p()
        if refined_regions is None:
            rc = self.L.ak_mesh_create(i3(*nmesh), i3(*nblock), d3(*xmin), d3(*xmax), i6(*bcs), nranks, C.byref(self.h))
        else:
            reg = np.ascontiguousarray(np.array(refined_regions, np.float64).reshape(-1, 7))
            rc = self.L.ak_mesh_create_smr(i3(*nmesh), i3(*nblock), d3(*xmin), d3(*xmax), i6(*bcs), nranks, len(reg),
                                           reg.ctypes.data_as(C.POINTER(C.c_double)), C.byref(self.h))
        if rc:
            raise _lib.EngineError(self.L.ak_mesh_last_error().decode())
        self._refresh()

    def _refresh(self):
        a, b = C.c_int(), C.c_int()
        self.L.ak_mesh_info(self.h, C.byref(a), C.byref(b))
        self.nmb_total, self.root_level = a.value, b.value
        self.lloc = np.zeros((self.nmb_total, 4), np.int32)
        self.rank_of = np.zeros(self.nmb_total, np.int32)
        self.gids_start = np.zeros(self.nranks, np.int32)
        self.nmb_rank = np.zeros(self.nranks, np.int32)
        ip = lambda x: x.ctypes.data_as(C.POINTER(C.c_int))
        self.L.ak_mesh_tables(self.h, ip(self.lloc), ip(self.rank_of), ip(self.gids_start), ip(self.nmb_rank))

    # -- adaptive mesh refinement, host maps (src/mesh/mesh_refinement.cpp:196-510) ---------------------------
    def set_amr(self, num_levels, refinement_interval=5, ncycle_check=1):
        """<mesh_refinement> refinement=adaptive: num_levels, refinement_interval, ncycle_check."""
        if self.L.ak_mesh_set_amr(self.h, int(num_levels), int(refinement_interval), int(ncycle_check)):
            raise _lib.EngineError("ak_mesh_set_amr failed")
        self.ncycle_check = int(ncycle_check)

    def adapt(self, ncycle, refine_flag):
        """One AdaptiveMeshRefinement bookkeeping step.  refine_flag: the criteria's verdict per block (+1 / -1 / 0).
        Returns (nnew, ndel, what happened to each OLD block); when the mesh changed, self.lloc etc. describe the
        new mesh and self.newtoold / self.oldtonew map the block indices."""
        flag = np.ascontiguousarray(refine_flag, np.int32).copy()
        assert flag.shape == (self.nmb_total,)
        ip = lambda x: x.ctypes.data_as(C.POINTER(C.c_int))
        nnew, ndel = C.c_int(), C.c_int()
        if self.L.ak_mesh_adapt(self.h, int(ncycle), ip(flag), C.byref(nnew), C.byref(ndel)):
            raise _lib.EngineError(self.L.ak_mesh_last_error().decode())
        if nnew.value or ndel.value:
            old_nmb = self.nmb_total
            self._refresh()
            self.newtoold = np.zeros(self.nmb_total, np.int32)
            self.oldtonew = np.zeros(old_nmb, np.int32)
            self.L.ak_mesh_amr_maps(self.h, ip(self.newtoold), ip(self.oldtonew), None, None)
        return nnew.value, ndel.value, flag

    def __del__(self):
        try:
            self.L.ak_mesh_destroy(self.h)
        except Exception:
            pass

    def rank_tables(self, rank):
        n = int(self.nmb_rank[rank])
        mb_size = np.zeros((n, 9))
        mb_bcs = np.zeros((n, 6), np.int32)
        nghbr = np.zeros((n, 56, 4), np.int32)
        self.L.ak_mesh_rank_tables(self.h, rank, mb_size.ctypes.data_as(C.POINTER(C.c_double)),
                                   mb_bcs.ctypes.data_as(C.POINTER(C.c_int)), nghbr.ctypes.data_as(C.POINTER(C.c_int)))
        return mb_size, mb_bcs, nghbr

    def pgen_orszag_tang(self, rank, ng, gamma, pinned=False):
        """Orszag-Tang ICs for the rank's blocks as host arrays (u0, b1f, b2f, b3f)."""
        import torch
        n = int(self.nmb_rank[rank])
        n1, n2, n3 = (self.nblock[d] + 2 * ng for d in range(3))
        mk = lambda *s: torch.zeros(*s, dtype=torch.float64, pin_memory=pinned)
        u0, b1, b2, b3 = mk(n, 5, n3, n2, n1), mk(n, n3, n2, n1 + 1), mk(n, n3, n2 + 1, n1), mk(n, n3 + 1, n2, n1)
        dp = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_double))
        if self.L.ak_pgen_orszag_tang(self.h, rank, ng, gamma, dp(u0), dp(b1), dp(b2), dp(b3)):
            raise _lib.EngineError("ak_pgen_orszag_tang failed")
        return u0, b1, b2, b3

    def pgen_blast(self, rank, ng, gamma, pi_amb=0.1, di_amb=1.0, prat=100.0, drat=1.0, b_amb=1.0,
                   inner_radius=0.1, outer_radius=0.1, pinned=False):
        """Blast ICs for the rank's blocks as host arrays (u0, b1f, b2f, b3f), active zones filled."""
        import torch
        n = int(self.nmb_rank[rank])
        n1, n2, n3 = (self.nblock[d] + 2 * ng for d in range(3))
        mk = lambda *s: torch.zeros(*s, dtype=torch.float64, pin_memory=pinned)
        u0, b1, b2, b3 = mk(n, 5, n3, n2, n1), mk(n, n3, n2, n1 + 1), mk(n, n3, n2 + 1, n1), mk(n, n3 + 1, n2, n1)
        prm = (C.c_double * 7)(pi_amb, di_amb, prat, drat, b_amb, inner_radius, outer_radius)
        dp = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_double))
        rc = self.L.ak_pgen_blast(self.h, rank, ng, gamma, prm, dp(u0), dp(b1), dp(b2), dp(b3))
        if rc:
            raise _lib.EngineError("ak_pgen_blast failed")
        return u0, b1, b2, b3
