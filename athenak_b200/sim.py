"""Python handle on the C++ host driver (athenak_b200/csrc/host_driver.cpp): TaskList + Driver +
MHD task registration in C++, owning the device arrays; Python only feeds ICs and reads results."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _lib
from .mhd import FLT_MAX, FLT_MIN

INTEGRATOR = {"rk1": 1, "rk2": 2, "rk3": 3}


class Simulation:
    def __init__(self, mesh, rank=0, *, ng=4, recon="ppm4", rsolver="hlld", gamma=1.6666667, integrator="rk3",
                 cfl_number=0.3, tlim=1.0, nlim=-1, fused=True, dfloor=FLT_MIN, pfloor=FLT_MIN, tfloor=FLT_MIN,
                 sfloor=FLT_MIN, sigma_max=FLT_MAX, physics="mhd", nscalars=0, eos="ideal", iso_sound_speed=1.0, fofc=False, multilevel=False):
        """physics="hydro": the hydro twin (rsolver in llf/hlle/hllc/roe; 1-D/2-D/3-D blocks); the state is u0 only."""
        self.L = _lib.load()
        self.mesh, self.rank, self.ng = mesh, rank, ng
        nmb = int(mesh.nmb_rank[rank])
        nx = mesh.nblock
        mb_size, mb_bcs, nghbr = mesh.rank_tables(rank)
        self.hydro = physics == "hydro"
        rs = _lib.HYD_RSOLVER[rsolver] if self.hydro else _lib.RSOLVER[rsolver]
        mcfg = _lib.MhdCfg(nmb, ng, nx[0], nx[1], nx[2], _lib.RECON[recon], rs, gamma,
                           dfloor, pfloor, tfloor, sfloor, sigma_max, int(nscalars), _lib.EOS[eos], iso_sound_speed, 1 if fofc else 0)
        self.cfg = _lib.SimCfg(mcfg, INTEGRATOR[integrator], cfl_number, tlim, nlim, rank, mesh.nranks,
                               int(mesh.gids_start[rank]), mesh.nmb_total, 1 if fused else 0)
        self.h = C.c_void_p()
        rc = self.L.ak_sim_create(C.byref(self.cfg), mb_size.ctypes.data_as(C.POINTER(C.c_double)),
                                  nghbr.ctypes.data_as(C.POINTER(C.c_int)), C.byref(self.h))
        if rc:
            raise _lib.EngineError("ak_sim_create failed: " + (self.L.ak_last_error(None) or b"").decode())
        ipp = C.POINTER(C.c_int)
        mesh_bcs = np.asarray(mesh.bcs, np.int32)
        if self.L.ak_sim_set_bcs(self.h, mesh_bcs.ctypes.data_as(ipp), mb_bcs.ctypes.data_as(ipp)):
            raise _lib.EngineError(self.L.ak_bcs_last_error(None).decode())
        if multilevel:       # statically refined mesh: levels from the mesh (hydro, one rank)
            lev = np.ascontiguousarray(mesh.lloc[int(mesh.gids_start[rank]):int(mesh.gids_start[rank]) + nmb, 3], np.int32)
            if self.L.ak_sim_set_levels(self.h, lev.ctypes.data_as(ipp)):
                raise _lib.EngineError("ak_sim_set_levels failed: " + (self.L.ak_smr_last_error(None) or b"").decode())
        n1, n2, n3 = (nx[d] + 2 * ng if (d == 0 or nx[d] > 1) else 1 for d in range(3))
        self.shapes = [(nmb, (4 if eos == "isothermal" else 5) + int(nscalars), n3, n2, n1), (nmb, n3, n2, n1 + 1), (nmb, n3, n2 + 1, n1), (nmb, n3 + 1, n2, n1)]
        if self.hydro:
            self.shapes = self.shapes[:1]

    def __del__(self):
        try:
            self.L.ak_sim_destroy(self.h)
        except Exception:
            pass

    def comm_init(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        if self.L.ak_sim_comm_init(self.h, buf):
            raise _lib.EngineError("ak_sim_comm_init failed")

    def unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        if self.L.ak_nccl_unique_id(buf):
            raise _lib.EngineError("NCCL not available")
        return buf.raw

    @staticmethod
    def _ptr(t):
        if t is None:
            return C.c_void_p(None)
        return C.c_void_p(t.data_ptr() if hasattr(t, "data_ptr") else t.ctypes.data)

    def _check_buffers(self, arrays, what):
        """The C side copies exactly prod(shape) doubles per array: anything else is an out-of-bounds host access."""
        for t, shape in zip(arrays, self.shapes):
            tshape = tuple(t.shape)
            dt = str(t.dtype)
            contig = t.is_contiguous() if hasattr(t, "is_contiguous") else t.flags["C_CONTIGUOUS"]
            on_host = (not hasattr(t, "is_cuda")) or (not t.is_cuda)
            if tshape != tuple(shape) or "float64" not in dt or not contig or not on_host:
                raise _lib.EngineError(f"{what}: expected a C-contiguous float64 host array of shape {tuple(shape)}, got {tshape} {dt}")

    def set_state(self, u0, b1f=None, b2f=None, b3f=None):
        """Host arrays (numpy or torch CPU tensors, C-contiguous float64, full ghost-inclusive shape)."""
        if not self.hydro and (b1f is None or b2f is None or b3f is None):
            raise _lib.EngineError("set_state: MHD needs the three face fields")
        self._check_buffers([u0] if self.hydro else [u0, b1f, b2f, b3f], "set_state")
        if self.L.ak_sim_set_state(self.h, *[self._ptr(t) for t in (u0, b1f, b2f, b3f)]):
            raise _lib.EngineError("ak_sim_set_state failed")

    def get_state(self, out=None):
        out = out or [np.empty(s) for s in self.shapes]
        self._check_buffers(out, "get_state")
        ptrs = [self._ptr(t) for t in out] + [self._ptr(None)] * (4 - len(out))
        if self.L.ak_sim_get_state(self.h, *ptrs):
            raise _lib.EngineError("ak_sim_get_state failed")
        return out

    def get_prim(self):
        """(w0, bcc0) of the current state as numpy arrays, ghost zones included (bcc0 is None for the hydro twin)."""
        w0 = np.empty(self.shapes[0])
        bcc = None if self.hydro else np.empty((self.shapes[0][0], 3) + tuple(self.shapes[0][2:]))
        if self.L.ak_sim_get_prim(self.h, self._ptr(w0), self._ptr(bcc)):
            raise _lib.EngineError("ak_sim_get_prim failed")
        return w0, bcc

    def initialize(self):
        if self.L.ak_sim_initialize(self.h):
            raise _lib.EngineError("ak_sim_initialize failed")

    def execute(self, ncycles=-1):
        return self.L.ak_sim_execute(self.h, ncycles)

    # ---- adaptive mesh refinement inside the C++ driver (one rank) ----
    @staticmethod
    def _crit_rows(criteria):
        from .amr import METHOD
        rows = []
        for c in criteria:
            if c.method != "location" and c.variable not in ("mhd_w_d", "mhd_u_d", "hydro_w_d", "hydro_u_d"):
                raise _lib.EngineError(f"athenak_b200: refinement variable {c.variable!r} not supported")
            rows.append([float(METHOD[c.method]), 1.0 if c.variable.endswith("_w_d") else 0.0, c.value_min, c.value_max, *c.location])
        return np.ascontiguousarray(rows, np.float64).reshape(-1, 8)

    def _after_mesh_change(self):
        cnt = (C.c_long * 4)()
        self.L.ak_sim_amr_status(self.h, cnt)
        self.shapes = [(int(cnt[0]),) + tuple(s[1:]) for s in self.shapes]
        return dict(nmb=int(cnt[0]), nadapt=int(cnt[1]), ncreated=int(cnt[2]), ndeleted=int(cnt[3]))

    def restart_from(self, time, dt, ncycle):
        """After set_state: the loaded arrays (ghost zones included) are the state of (time, dt, ncycle); primitives and dtnew follow."""
        self.L.ak_sim_set_time(self.h, float(time), float(dt), int(ncycle))
        self.L.ak_sim_prims_newdt(self.h)

    def adapt(self, criteria):
        """One AdaptiveMeshRefinement call followed by Mesh::NewTimeStep; returns (changed, status dict)."""
        rows = self._crit_rows(criteria)
        ch = C.c_int()
        if self.L.ak_sim_adapt(self.h, self.mesh.h, rows.ctypes.data_as(C.POINTER(C.c_double)), len(rows), C.byref(ch)):
            raise _lib.EngineError("ak_sim_adapt failed: " + (self.L.ak_amr_last_error() or b"").decode())
        self.L.ak_sim_mesh_new_dt(self.h)
        self.mesh._refresh()
        return bool(ch.value), self._after_mesh_change()

    def execute_amr(self, criteria, ncycles=-1):
        """Driver::Execute with AdaptiveMeshRefinement after every cycle; self.mesh must be adaptive (Mesh.set_amr) and the
        simulation multilevel.  Returns (cycles done, status dict)."""
        rows = self._crit_rows(criteria)
        n = self.L.ak_sim_execute_amr(self.h, self.mesh.h, rows.ctypes.data_as(C.POINTER(C.c_double)), len(rows), int(ncycles))
        if n < 0:
            raise _lib.EngineError("ak_sim_execute_amr failed: " + (self.L.ak_amr_last_error() or b"").decode())
        self.mesh._refresh()
        return n, self._after_mesh_change()

    def sync(self):
        self.L.ak_sim_sync(self.h)

    def status(self):
        st, cnt = (C.c_double * 3)(), (C.c_long * 8)()
        self.L.ak_sim_status(self.h, st, cnt)
        return dict(time=st[0], dt=st[1], dtnew=st[2], ncycle=cnt[0], launches=cnt[1], nfloor=list(cnt[2:5]), stream=cnt[5],
                    peer_push=bool(cnt[6]), peer_exchanges=cnt[7])


class HostPipeline:
    """Pipelined host-buffer stepping (ak_pipe_*, include/athenak_b200.h): a stream of independent jobs, each
    `set_state(host in) -> initialize -> execute(ncycles) -> get_state(host out)`, over several Simulation instances on one
    device.  One upload and one download stream are shared by all instances, so consecutive jobs' H2D, step and D2H overlap and
    both PCIe directions stay busy; every job's result equals the serial sequence on one instance bit for bit."""

    def __init__(self, sims):
        self.sims = list(sims)          # keeps the engine instances alive
        self.L = self.sims[0].L
        arr = (C.c_void_p * len(self.sims))(*[s.h for s in self.sims])
        self.h = C.c_void_p()
        if self.L.ak_pipe_create(arr, len(self.sims), C.byref(self.h)):
            raise _lib.EngineError("ak_pipe_create failed (instances must live on one device)")
        self._keep = {}

    def streams(self):
        a, b = C.c_void_p(), C.c_void_p()
        self.L.ak_pipe_streams(self.h, C.byref(a), C.byref(b))
        return int(a.value or 0), int(b.value or 0)

    def submit(self, state_in, state_out, ncycles=1):
        s0 = self.sims[0]
        s0._check_buffers(list(state_in), "HostPipeline.submit (in)")
        s0._check_buffers(list(state_out), "HostPipeline.submit (out)")
        pin = [s0._ptr(t) for t in state_in] + [s0._ptr(None)] * (4 - len(state_in))
        pout = [s0._ptr(t) for t in state_out] + [s0._ptr(None)] * (4 - len(state_out))
        t = self.L.ak_pipe_submit(self.h, *pin, *pout, int(ncycles))
        if t < 0:
            raise _lib.EngineError("ak_pipe_submit failed")
        self._keep[t] = (state_in, state_out)       # the buffers must outlive the job
        return t

    def wait(self, ticket=-1):
        if self.L.ak_pipe_wait(self.h, int(ticket)):
            raise _lib.EngineError("a pipelined job failed: " + (self.L.ak_last_error(None) or b"").decode())
        if ticket < 0:
            self._keep.clear()
        else:
            self._keep.pop(ticket, None)

    def close(self):
        if self.h:
            self.L.ak_pipe_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
