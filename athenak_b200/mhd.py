"""Host-side mirror of the reference's `mhd::MHD` module for the Newtonian ideal-gas path.

Reference: src/mhd/mhd.hpp:87-207 (state owner), src/mhd/mhd_tasks.cpp:38-84 (task list).
The object owns every array of one MeshBlockPack as FP64 device tensors in the reference
layout and exposes one method per task with the reference's task signature
`(driver, stage) -> TaskStatus`; each method body is one call through the C ABI
(include/athenak_b200.h).  PyTorch is used for device memory, streams and (multi-rank)
torch.distributed point-to-point only.
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np
import torch

from . import lib as _lib

FLT_MIN = float(np.finfo(np.float32).tiny)
FLT_MAX = float(np.finfo(np.float32).max)


class _DevPtr:
    """Zero-copy view of library-owned device memory as a torch tensor (CUDA array interface)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class TaskStatus(enum.Enum):  # src/tasklist/task_list.hpp:30
    fail = 0
    complete = 1
    incomplete = 2


class MHD:
    def __init__(self, nmb, ng, nx, mb_size, nghbr, *, recon="ppm4", rsolver="hlld", gamma,
                 dfloor=FLT_MIN, pfloor=FLT_MIN, tfloor=FLT_MIN, sfloor=FLT_MIN, sigma_max=FLT_MAX,
                 my_rank=0, gids_start=0, device="cuda", process_group=None, fused=True,
                 mesh_bcs=None, mb_bcs=None, nscalars=0, eos="ideal", iso_sound_speed=1.0,
                 fofc=False, mb_lev=None):
        """mb_lev: logical level of every block (Mesh.lloc[:, 3]) on a statically refined mesh; enables the fine/coarse
        boundary path for U, B and the edge EMFs (ak_smr_*; one rank)."""
        if not torch.cuda.is_available():
            raise _lib.EngineError("athenak_b200: no CUDA device; the engine has no CPU path")
        self.L = _lib.load()
        self.device = torch.device(device)
        self.nmb, self.ng, self.nx = int(nmb), int(ng), tuple(int(v) for v in nx)
        nx1, nx2, nx3 = self.nx
        self.multi_d, self.three_d = nx2 > 1, nx3 > 1      # src/mesh/mesh.cpp:70-79
        self.n = (nx1 + 2 * ng, nx2 + 2 * ng if self.multi_d else 1, nx3 + 2 * ng if self.three_d else 1)
        self.nsweeps = 1 + int(self.multi_d) + int(self.three_d)
        self.nfl = 4 if eos == "isothermal" else 5       # src/mhd/mhd.cpp:100-108: nmhd
        self.nvar = self.nfl + int(nscalars)      # src/mhd/mhd.cpp: nmhd + nscalars
        self.cfg = _lib.MhdCfg(self.nmb, ng, nx1, nx2, nx3, _lib.RECON[recon], _lib.RSOLVER[rsolver],
                               gamma, dfloor, pfloor, tfloor, sfloor, sigma_max, int(nscalars), _lib.EOS[eos], iso_sound_speed, 1 if fofc else 0)
        self.fofc, self.nfofc = bool(fofc), 0     # <mhd>/fofc, EventCounters::nfofc (src/mesh/mesh.hpp:70-74)
        self.fused = fused
        self.my_rank, self.gids_start, self.pg = my_rank, gids_start, process_group
        n1, n2, n3 = self.n
        z = lambda *s: torch.zeros(*s, dtype=torch.float64, device=self.device)
        # src/mhd/mhd.cpp:148-161,329-354
        nv = self.nvar
        self.u0, self.u1, self.w0 = z(nmb, nv, n3, n2, n1), z(nmb, nv, n3, n2, n1), z(nmb, nv, n3, n2, n1)
        self.bcc0 = z(nmb, 3, n3, n2, n1)
        self.b0 = [z(nmb, n3, n2, n1 + 1), z(nmb, n3, n2 + 1, n1), z(nmb, n3 + 1, n2, n1)]
        self.b1 = [torch.zeros_like(a) for a in self.b0]
        self.uflx = [z(nmb, nv, n3, n2, n1 + 1), z(nmb, nv, n3, n2 + 1, n1), z(nmb, nv, n3 + 1, n2, n1)]
        self.femf = [z(nmb, n3, n2, n1) for _ in range(6)]   # e3x1,e2x1,e1x2,e3x2,e2x3,e1x3
        self.efld = [z(nmb, n3 + 1, n2 + 1, n1), z(nmb, n3 + 1, n2, n1 + 1), z(nmb, n3, n2 + 1, n1 + 1)]
        self.mb_size_h = np.ascontiguousarray(mb_size, dtype=np.float64).reshape(nmb, 9)
        self.mb_size = torch.from_numpy(self.mb_size_h).to(self.device)
        self.arrays = _lib.MhdArrays()
        self._bind()
        self.h = C.c_void_p()
        _lib.check(self.L.ak_mhd_create(C.byref(self.cfg), C.byref(self.h)))
        # neighbour table -> halo plan
        self.nghbr = np.ascontiguousarray(nghbr, dtype=np.int32).reshape(nmb, 56, 4)
        self.plan = C.c_void_p()
        _lib.check(self.L.ak_halo_create(C.byref(self.cfg), self.nghbr.ctypes.data_as(C.POINTER(C.c_int)),
                                         my_rank, gids_start, C.byref(self.plan)))
        self.smr = None
        if mb_lev is not None:       # multilevel mesh: coarse arrays (mhd.cpp:163-176) + SMR plan
            self.mb_lev = np.ascontiguousarray(mb_lev, dtype=np.int32).reshape(nmb)
            self.smr = C.c_void_p()
            ipp = C.POINTER(C.c_int)
            if self.L.ak_smr_create(C.byref(self.cfg), self.nghbr.ctypes.data_as(ipp), self.mb_lev.ctypes.data_as(ipp),
                                    my_rank, gids_start, C.byref(self.smr)):
                self.smr = None
                raise _lib.EngineError(self.L.ak_smr_last_error(None).decode())
            self.coarse_u0 = z(nmb, nv, self.L.ak_smr_coarse_cells(self.smr))
            self.coarse_b0 = [z(nmb, self.L.ak_smr_coarse_faces(self.smr, v)) for v in range(3)]
        self.peers = []
        for ip in range(self.L.ak_halo_npeers(self.plan) if self.smr is None else 0):
            r, ncc, nfc = C.c_int(), C.c_long(), C.c_long()
            _lib.check(self.L.ak_halo_peer(self.plan, ip, C.byref(r), C.byref(ncc), C.byref(nfc)))
            self.peers.append(dict(rank=r.value, ncc=ncc.value, nfc=nfc.value,
                                   scc=z(nv * ncc.value), rcc=z(nv * ncc.value), sfc=z(nfc.value), rfc=z(nfc.value)))
        # physical BCs (task 26); None = strictly periodic mesh
        self.bcs = None
        if mesh_bcs is not None and any(int(b) != _lib.BC["periodic"] for b in mesh_bcs):
            mesh_bcs = np.ascontiguousarray(mesh_bcs, dtype=np.int32).reshape(6)
            mb_bcs = np.ascontiguousarray(mb_bcs, dtype=np.int32).reshape(nmb, 6)
            ipp = C.POINTER(C.c_int)
            self.bcs = C.c_void_p()
            if self.L.ak_bcs_create(C.byref(self.cfg), mesh_bcs.ctypes.data_as(ipp), mb_bcs.ctypes.data_as(ipp),
                                    C.byref(self.bcs)):
                self.bcs = None
                raise _lib.EngineError(self.L.ak_bcs_last_error(None).decode())
            act = (True, self.multi_d, self.three_d)
            self.nbc = sum(act[d] and int(mesh_bcs[2 * d]) != _lib.BC["periodic"] for d in range(3))
            self.bcs_coarse = None
            if self.smr is not None:     # HydroBCsCoarse / BFieldBCsCoarse (bfield_bcs.cpp:44-63): the same fill on the coarse index space
                ccfg = _lib.MhdCfg.from_buffer_copy(self.cfg)
                ccfg.nx1, ccfg.nx2, ccfg.nx3 = nx1 // 2, max(1, nx2 // 2), max(1, nx3 // 2)
                self.bcs_coarse = C.c_void_p()
                if self.L.ak_bcs_create(C.byref(ccfg), mesh_bcs.ctypes.data_as(ipp), mb_bcs.ctypes.data_as(ipp),
                                        C.byref(self.bcs_coarse)):
                    self.bcs_coarse = None
                    raise _lib.EngineError(self.L.ak_bcs_last_error(None).decode())
        self.dtnew = FLT_MAX
        self.nfloor = [0, 0, 0]       # EventCounters neos_dfloor/efloor/tfloor (src/mesh/mesh.hpp:70-74)
        self.launches = 0             # kernels this object has launched (bench: gpu_launches)

    def _bind(self):
        a = self.arrays
        t = dict(u0=self.u0, u1=self.u1, w0=self.w0, bcc0=self.bcc0,
                 b0x1f=self.b0[0], b0x2f=self.b0[1], b0x3f=self.b0[2],
                 b1x1f=self.b1[0], b1x2f=self.b1[1], b1x3f=self.b1[2],
                 flx1=self.uflx[0], flx2=self.uflx[1], flx3=self.uflx[2],
                 e3x1=self.femf[0], e2x1=self.femf[1], e1x2=self.femf[2], e3x2=self.femf[3],
                 e2x3=self.femf[4], e1x3=self.femf[5], e1=self.efld[0], e2=self.efld[1], e3=self.efld[2],
                 mb_size=self.mb_size)
        for k, v in t.items():
            setattr(a, k, v.data_ptr())

    def __del__(self):
        try:
            if getattr(self, "bcs_coarse", None):
                self.L.ak_bcs_destroy(self.bcs_coarse)
            if getattr(self, "bcs", None):
                self.L.ak_bcs_destroy(self.bcs)
            if getattr(self, "smr", None):
                self.L.ak_smr_destroy(self.smr)
            if getattr(self, "plan", None):
                self.L.ak_halo_destroy(self.plan)
            if getattr(self, "h", None):
                self.L.ak_mhd_destroy(self.h)
        except Exception:
            pass

    # -- helpers -------------------------------------------------------------------------
    def _st(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ck(self, rc):
        _lib.check(rc, self.h)

    def load_state(self, u0, b1f, b2f, b3f):
        """Copy host (numpy) conserved state into the device arrays (ICs / restart)."""
        self.u0.copy_(torch.from_numpy(np.ascontiguousarray(u0)))
        for dst, src in zip(self.b0, (b1f, b2f, b3f)):
            dst.copy_(torch.from_numpy(np.ascontiguousarray(src)))

    # -- tasks (src/mhd/mhd_tasks.cpp:48-75), uniform-mesh subset -------------------------
    def CopyCons(self, driver, stage):
        if stage == 1:
            self._ck(self.L.ak_mhd_copy_cons(self.h, C.byref(self.arrays), self._st()))
            self.launches += 4
        return TaskStatus.complete

    def Fluxes(self, driver, stage):
        """mhd_tasks.cpp:177-218: CalculateFluxes, then FOFC when enabled."""
        self._ck(self.L.ak_mhd_fluxes(self.h, C.byref(self.arrays), self._st()))
        self.launches += self.nsweeps
        if self.fofc:
            g0, g1, bdt = driver.gam0[stage - 1], driver.gam1[stage - 1], driver.beta[stage - 1] * driver.dt
            n = C.c_int()
            self._ck(self.L.ak_mhd_fofc(self.h, C.byref(self.arrays), g0, g1, bdt, C.byref(n), self._st()))
            self.nfofc += n.value
            self.launches += 1 + self.nsweeps
        return TaskStatus.complete

    def RKUpdate(self, driver, stage):
        g0, g1, bdt = driver.gam0[stage - 1], driver.gam1[stage - 1], driver.beta[stage - 1] * driver.dt
        self._ck(self.L.ak_mhd_rk_update(self.h, C.byref(self.arrays), g0, g1, bdt, self._st()))
        self.launches += 1
        return TaskStatus.complete

    def EField(self, driver, stage):
        self._ck(self.L.ak_mhd_corner_e(self.h, C.byref(self.arrays), self._st()))
        self.launches += 1
        return TaskStatus.complete

    def CT(self, driver, stage):
        g0, g1, bdt = driver.gam0[stage - 1], driver.gam1[stage - 1], driver.beta[stage - 1] * driver.dt
        self._ck(self.L.ak_mhd_ct(self.h, C.byref(self.arrays), g0, g1, bdt, self._st()))
        self.launches += 1
        return TaskStatus.complete

    def StageFluxEmf(self, driver, stage):
        """Tasks 2 + 14 in one ABI call (ak_mhd_stage_flux_emf)."""
        if self.fofc:       # the FOFC estimate needs CopyCons over the whole arrays and the RK coefficients
            self.CopyCons(driver, stage)
            g0, g1, bdt = driver.gam0[stage - 1], driver.gam1[stage - 1], driver.beta[stage - 1] * driver.dt
            self._ck(self.L.ak_mhd_set_stage(self.h, g0, g1, bdt))
            self.launches += 1 + self.nsweeps
        self._ck(self.L.ak_mhd_stage_flux_emf(self.h, C.byref(self.arrays), self._st()))
        self.launches += self.nsweeps + 1
        return TaskStatus.complete

    def StageUpdate(self, driver, stage):
        """Tasks 1 + 5 + 17 in one ABI call (ak_mhd_stage_update)."""
        g0, g1, bdt = driver.gam0[stage - 1], driver.gam1[stage - 1], driver.beta[stage - 1] * driver.dt
        self._ck(self.L.ak_mhd_stage_update(self.h, C.byref(self.arrays), stage, g0, g1, bdt, self._st()))
        self.launches += 1
        return TaskStatus.complete

    def _smr_ck(self, rc):
        if rc:
            raise _lib.EngineError(self.L.ak_smr_last_error(self.smr).decode())

    def FluxCorrect(self, driver, stage):
        """SendFlux/RecvFlux (mhd_tasks.cpp:52-53 -> flux_correct_cc.cpp): multilevel meshes only."""
        if self.smr is not None:
            self._smr_ck(self.L.ak_smr_flux_correct_cc(self.smr, *[f.data_ptr() for f in self.uflx], self.nvar, self._st()))
            self.launches += 1
        return TaskStatus.complete

    def Prolongate(self, driver, stage):
        """mhd_tasks.cpp:527-551: FillCoarseInBndryCC/FC, then ProlongateCC/FC (multilevel meshes only)."""
        if self.smr is None:
            return TaskStatus.complete
        st, L = self._st(), self.L
        u, cu = self.u0.data_ptr(), self.coarse_u0.data_ptr()
        b, cb = [a.data_ptr() for a in self.b0], [a.data_ptr() for a in self.coarse_b0]
        self._smr_ck(L.ak_smr_prolongate_cc_steps(self.smr, u, cu, self.nvar, 1, st))
        self._smr_ck(L.ak_smr_prolongate_fc_steps(self.smr, *b, *cb, 1, st))
        if self.bcs is not None:     # physical BCs on the coarse arrays before the prolongation reads them
            if L.ak_bcs_apply(self.bcs_coarse, cu, self.nvar, *cb, st):
                raise _lib.EngineError(L.ak_bcs_last_error(self.bcs_coarse).decode())
            self.launches += self.nbc
        self._smr_ck(L.ak_smr_prolongate_cc_steps(self.smr, u, cu, self.nvar, 2, st))
        self._smr_ck(L.ak_smr_prolongate_fc_steps(self.smr, *b, *cb, 2, st))
        self.launches += 5
        return TaskStatus.complete

    def SendRecvE(self, driver, stage):
        """Tasks 15/16 SendE/RecvE: EMF flux correction (src/bvals/flux_correct_fc.cpp)."""
        st, L = self._st(), self.L
        e = [t.data_ptr() for t in self.efld]
        if self.smr is not None:     # with coarser / finer neighbours: pack (restricted) + gather-form sum / zero / average
            self._smr_ck(L.ak_smr_emf_correct(self.smr, *e, st))
            self.launches += 2
            return TaskStatus.complete
        reqs = []
        if self.peers:
            import torch.distributed as dist
            ops = []
            for ip, p in enumerate(self.peers):
                if "semf" not in p:
                    cnt, ptr = C.c_long(), C.c_void_p()
                    _lib.check(L.ak_halo_emf_peer(self.plan, ip, C.byref(cnt), C.byref(ptr)))
                    p["semf"] = torch.zeros(cnt.value, dtype=torch.float64, device=self.device)
                    # the peer's message is received straight into the plan's receive segment
                    p["remf"] = torch.as_tensor(_DevPtr(ptr.value, cnt.value), device=self.device)
                _lib.check(L.ak_halo_emf_pack(self.plan, *e, ip, p["semf"].data_ptr(), st))
                self.launches += 1
                ops.append(dist.P2POp(dist.isend, p["semf"], p["rank"], group=self.pg))
                ops.append(dist.P2POp(dist.irecv, p["remf"], p["rank"], group=self.pg))
            reqs = dist.batch_isend_irecv(ops)
        _lib.check(L.ak_halo_emf_pack_local(self.plan, *e, st))
        self.launches += 1
        for rq in reqs:
            rq.wait()
        _lib.check(L.ak_halo_emf_sumavg(self.plan, *e, st))
        self.launches += 1
        return TaskStatus.complete

    def _exchange(self, which):
        """SendU/RecvU (which='u') or SendB/RecvB (which='b'): src/bvals/bvals_cc.cpp:42-447,
        bvals_fc.cpp:63-436.  Off-rank slabs: pack -> one message per peer -> unpack."""
        st = self._st()
        L = self.L
        if self.smr is not None:     # RestrictU/B + SendU/B + RecvU/B across levels (mhd_tasks.cpp:55-57, 63-65)
            if which == "u":
                u, cu = self.u0.data_ptr(), self.coarse_u0.data_ptr()
                self._smr_ck(L.ak_smr_restrict_cc(self.smr, u, cu, self.nvar, st))
                self._smr_ck(L.ak_smr_exchange_cc(self.smr, u, cu, self.nvar, st))
            else:
                b, cb = [a.data_ptr() for a in self.b0], [a.data_ptr() for a in self.coarse_b0]
                self._smr_ck(L.ak_smr_restrict_fc(self.smr, *b, *cb, st))
                self._smr_ck(L.ak_smr_exchange_fc(self.smr, *b, *cb, st))
            self.launches += 2
            return
        if self.peers:
            import torch.distributed as dist
            ops = []
            for ip, p in enumerate(self.peers):
                if which == "u":
                    _lib.check(L.ak_halo_pack_cc(self.plan, self.u0.data_ptr(), self.nvar, ip, p["scc"].data_ptr(), st))
                    s, r = p["scc"], p["rcc"]
                else:
                    _lib.check(L.ak_halo_pack_fc(self.plan, *[a.data_ptr() for a in self.b0], ip, p["sfc"].data_ptr(), st))
                    s, r = p["sfc"], p["rfc"]
                self.launches += 1
                ops.append(dist.P2POp(dist.isend, s, p["rank"], group=self.pg))
                ops.append(dist.P2POp(dist.irecv, r, p["rank"], group=self.pg))
            reqs = dist.batch_isend_irecv(ops)
        if which == "u":
            _lib.check(L.ak_halo_local_cc(self.plan, self.u0.data_ptr(), self.nvar, st))
        else:
            _lib.check(L.ak_halo_local_fc(self.plan, *[a.data_ptr() for a in self.b0], st))
        self.launches += 1
        if self.peers:
            for rq in reqs:
                rq.wait()
            for ip, p in enumerate(self.peers):
                if which == "u":
                    _lib.check(L.ak_halo_unpack_cc(self.plan, self.u0.data_ptr(), self.nvar, ip, p["rcc"].data_ptr(), st))
                else:
                    _lib.check(L.ak_halo_unpack_fc(self.plan, *[a.data_ptr() for a in self.b0], ip, p["rfc"].data_ptr(), st))
                self.launches += 1

    def HistorySums(self):
        """History-output sums over this pack's active cells (src/outputs/history.cpp:78-160,272-372):
        mass, 1/2/3-mom, tot-E, 1/2/3-KE (+ 1/2/3-ME for MHD)."""
        out, n = (C.c_double * 11)(), C.c_int()
        self._ck(self.L.ak_hst_sums(self.h, C.byref(self.arrays), out, C.byref(n), self._st()))
        self.launches += 2
        return np.array(out[:n.value])

    def SendRecvU(self, driver, stage):
        self._exchange("u")
        return TaskStatus.complete

    def SendRecvB(self, driver, stage):
        self._exchange("b")
        return TaskStatus.complete

    def ApplyPhysicalBCs(self, driver, stage):
        """Task 26 (src/mhd/mhd_tasks.cpp:504-520): HydroBCs + BFieldBCs; no-op when strictly periodic."""
        if self.bcs is None:
            return TaskStatus.complete
        if self.L.ak_bcs_apply(self.bcs, self.u0.data_ptr(), self.nvar, *[a.data_ptr() for a in self.b0], self._st()):
            raise _lib.EngineError(self.L.ak_bcs_last_error(self.bcs).decode())
        self.launches += self.nbc
        return TaskStatus.complete

    def ConToPrim(self, driver, stage, count=True):
        n1, n2, n3 = self.n
        nf = (C.c_int * 3)() if count else None
        self._ck(self.L.ak_mhd_cons2prim(self.h, C.byref(self.arrays), 0, n1 - 1, 0, n2 - 1, 0, n3 - 1, nf, self._st()))
        self.launches += 1
        if count:
            for q in range(3):
                self.nfloor[q] += nf[q]
        return TaskStatus.complete

    def NewTimeStep(self, driver, stage):
        if stage != driver.nexp_stages:      # src/mhd/mhd_newdt.cpp:32-34
            return TaskStatus.complete
        dt3 = (C.c_double * 3)()
        self._ck(self.L.ak_mhd_new_dt(self.h, C.byref(self.arrays), dt3, self._st()))
        self.launches += 1
        self._set_dtnew(dt3)
        return TaskStatus.complete

    def _set_dtnew(self, dt3):
        self.dtnew = dt3[0]                        # src/mhd/mhd_newdt.cpp:152-155
        if self.multi_d:
            self.dtnew = min(self.dtnew, dt3[1])
        if self.three_d:
            self.dtnew = min(self.dtnew, dt3[2])

    def ConToPrimNewDt(self, driver, stage):
        """Tasks 27+28 fused on the last stage (ak_mhd_cons2prim_newdt)."""
        if stage != driver.nexp_stages:
            return self.ConToPrim(driver, stage, count=False)   # counters accumulate on the device
        nf, dt3 = (C.c_int * 3)(), (C.c_double * 3)()
        self._ck(self.L.ak_mhd_cons2prim_newdt(self.h, C.byref(self.arrays), nf, dt3, self._st()))
        self.launches += 1
        for q in range(3):
            self.nfloor[q] += nf[q]
        self._set_dtnew(dt3)
        return TaskStatus.complete
