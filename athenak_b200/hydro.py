"""Host-side mirror of the reference's `hydro::Hydro` module (Newtonian ideal gas), the hydro twin
of athenak_b200.mhd.MHD.

Reference: src/hydro/hydro.hpp:44-161 (state owner), src/hydro/hydro_tasks.cpp:47-79 (task list).
One method per task with the reference's task signature `(driver, stage) -> TaskStatus`; each
body is one call through the C ABI (include/athenak_b200.h, ak_hydro_*).  MeshBlocks may be 1-D,
2-D or 3-D: inactive dimensions have extent 1 and no ghost cells (src/mesh/mesh.cpp:285-333).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import lib as _lib
from .mhd import FLT_MAX, FLT_MIN, TaskStatus


class Hydro:
    def __init__(self, nmb, ng, nx, mb_size, nghbr, *, recon="plm", rsolver="hllc", gamma,
                 dfloor=FLT_MIN, pfloor=FLT_MIN, tfloor=FLT_MIN, sfloor=FLT_MIN,
                 my_rank=0, gids_start=0, device="cuda", process_group=None, mesh_bcs=None, mb_bcs=None, nscalars=0,
                 eos="ideal", iso_sound_speed=1.0, mb_lev=None):
        """mb_lev: logical level of every block (Mesh.lloc[:, 3]) on a statically refined mesh; enables the
        fine/coarse boundary path (ak_smr_*: restriction, multi-level exchange, prolongation, flux correction)."""
        if not torch.cuda.is_available():
            raise _lib.EngineError("athenak_b200: no CUDA device; the engine has no CPU path")
        self.L = _lib.load()
        self.device = torch.device(device)
        self.nmb, self.ng, self.nx = int(nmb), int(ng), tuple(int(v) for v in nx)
        nx1, nx2, nx3 = self.nx
        self.multi_d, self.three_d = nx2 > 1, nx3 > 1
        self.n = (nx1 + 2 * ng, nx2 + 2 * ng if self.multi_d else 1, nx3 + 2 * ng if self.three_d else 1)
        self.nfl = 4 if eos == "isothermal" else 5
        self.nvar = nv = self.nfl + int(nscalars)      # src/hydro/hydro.cpp:75,117: nhydro + nscalars
        self.cfg = _lib.MhdCfg(self.nmb, ng, nx1, nx2, nx3, _lib.RECON[recon], _lib.HYD_RSOLVER[rsolver],
                               gamma, dfloor, pfloor, tfloor, sfloor, FLT_MAX, int(nscalars), _lib.EOS[eos], iso_sound_speed)
        self.my_rank, self.gids_start, self.pg = my_rank, gids_start, process_group
        n1, n2, n3 = self.n
        z = lambda *s: torch.zeros(*s, dtype=torch.float64, device=self.device)
        # src/hydro/hydro.cpp:100-120,268-290
        self.u0, self.u1, self.w0 = z(nmb, nv, n3, n2, n1), z(nmb, nv, n3, n2, n1), z(nmb, nv, n3, n2, n1)
        self.uflx = [z(nmb, nv, n3, n2, n1 + 1), z(nmb, nv, n3, n2 + 1, n1), z(nmb, nv, n3 + 1, n2, n1)]
        self.mb_size_h = np.ascontiguousarray(mb_size, dtype=np.float64).reshape(nmb, 9)
        self.mb_size = torch.from_numpy(self.mb_size_h).to(self.device)
        self.arrays = _lib.MhdArrays()
        for k, v in dict(u0=self.u0, u1=self.u1, w0=self.w0, flx1=self.uflx[0], flx2=self.uflx[1],
                         flx3=self.uflx[2], mb_size=self.mb_size).items():
            setattr(self.arrays, k, v.data_ptr())
        self.h = C.c_void_p()
        _lib.check(self.L.ak_hydro_create(C.byref(self.cfg), C.byref(self.h)))
        self.nghbr = np.ascontiguousarray(nghbr, dtype=np.int32).reshape(nmb, 56, 4)
        self.plan = C.c_void_p()
        self.smr = None
        if mb_lev is not None:       # multilevel mesh (src/mesh/mesh.hpp multilevel): coarse array + SMR plan
            self.mb_lev = np.ascontiguousarray(mb_lev, dtype=np.int32).reshape(nmb)
            self.smr = C.c_void_p()
            ipp = C.POINTER(C.c_int)
            if self.L.ak_smr_create(C.byref(self.cfg), self.nghbr.ctypes.data_as(ipp), self.mb_lev.ctypes.data_as(ipp),
                                    my_rank, gids_start, C.byref(self.smr)):
                self.smr = None
                raise _lib.EngineError(self.L.ak_smr_last_error(None).decode())
            self.coarse_u0 = z(nmb, nv, self.L.ak_smr_coarse_cells(self.smr))      # hydro.cpp:121-128
        else:
            _lib.check(self.L.ak_halo_create(C.byref(self.cfg), self.nghbr.ctypes.data_as(C.POINTER(C.c_int)),
                                             my_rank, gids_start, C.byref(self.plan)))
        self.peers = []
        for ip in range(self.L.ak_halo_npeers(self.plan) if self.smr is None else 0):
            r, ncc, nfc = C.c_int(), C.c_long(), C.c_long()
            _lib.check(self.L.ak_halo_peer(self.plan, ip, C.byref(r), C.byref(ncc), C.byref(nfc)))
            self.peers.append(dict(rank=r.value, ncc=ncc.value, scc=z(nv * ncc.value), rcc=z(nv * ncc.value)))
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
            if self.smr is not None:     # HydroBCsCoarse (hydro_bcs.cpp:51-67): the same fill on the coarse index space
                ccfg = _lib.MhdCfg.from_buffer_copy(self.cfg)
                ccfg.nx1, ccfg.nx2, ccfg.nx3 = nx1 // 2, max(1, nx2 // 2), max(1, nx3 // 2)
                self.bcs_coarse = C.c_void_p()
                if self.L.ak_bcs_create(C.byref(ccfg), mesh_bcs.ctypes.data_as(ipp), mb_bcs.ctypes.data_as(ipp),
                                        C.byref(self.bcs_coarse)):
                    raise _lib.EngineError(self.L.ak_bcs_last_error(None).decode())
        self.dtnew = FLT_MAX
        self.nfloor = [0, 0, 0]
        self.launches = 0
        self.nsweeps = 1 + int(self.multi_d) + int(self.three_d)

    def __del__(self):
        try:
            if getattr(self, "bcs", None):
                self.L.ak_bcs_destroy(self.bcs)
            if getattr(self, "bcs_coarse", None):
                self.L.ak_bcs_destroy(self.bcs_coarse)
            if getattr(self, "plan", None):
                self.L.ak_halo_destroy(self.plan)
            if getattr(self, "smr", None):
                self.L.ak_smr_destroy(self.smr)
            if getattr(self, "h", None):
                self.L.ak_hydro_destroy(self.h)
        except Exception:
            pass

    def _st(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ck(self, rc):
        _lib.check(rc, self.h)

    def load_state(self, u0):
        self.u0.copy_(torch.from_numpy(np.ascontiguousarray(u0)))

    # -- tasks (src/hydro/hydro_tasks.cpp:55-71), uniform-mesh subset -----------------------
    def CopyCons(self, driver, stage):
        if stage == 1:
            self._ck(self.L.ak_hydro_copy_cons(self.h, C.byref(self.arrays), self._st()))
            self.launches += 1
        return TaskStatus.complete

    def Fluxes(self, driver, stage):
        self._ck(self.L.ak_hydro_fluxes(self.h, C.byref(self.arrays), self._st()))
        self.launches += self.nsweeps
        return TaskStatus.complete

    def RKUpdate(self, driver, stage):
        g0, g1, bdt = driver.gam0[stage - 1], driver.gam1[stage - 1], driver.beta[stage - 1] * driver.dt
        self._ck(self.L.ak_hydro_rk_update(self.h, C.byref(self.arrays), g0, g1, bdt, self._st()))
        self.launches += 1
        return TaskStatus.complete

    def StageFused(self, driver, stage):
        """CopyCons + Fluxes + RKUpdate in one ABI call (ak_hydro_stage_fused)."""
        g0, g1, bdt = driver.gam0[stage - 1], driver.gam1[stage - 1], driver.beta[stage - 1] * driver.dt
        self._ck(self.L.ak_hydro_stage_fused(self.h, C.byref(self.arrays), stage, g0, g1, bdt, self._st()))
        self.launches += self.nsweeps + 1
        return TaskStatus.complete

    def _smr_ck(self, rc):
        if rc:
            raise _lib.EngineError(self.L.ak_smr_last_error(self.smr).decode())

    def FluxCorrect(self, driver, stage):
        """SendFlux/RecvFlux (hydro_tasks.cpp:58-59 -> flux_correct_cc.cpp): multilevel meshes only."""
        if self.smr is not None:
            self._smr_ck(self.L.ak_smr_flux_correct_cc(self.smr, *[f.data_ptr() for f in self.uflx], self.nvar, self._st()))
            self.launches += 1
        return TaskStatus.complete

    def SendRecvU(self, driver, stage):
        """SendU/RecvU (src/bvals/bvals_cc.cpp:42-447): same-rank gather + one message per peer.  On a multilevel
        mesh: RestrictU, SendU/RecvU across levels and Prolongate (hydro_tasks.cpp:63-68)."""
        st, L = self._st(), self.L
        if self.smr is not None:
            u, cu = self.u0.data_ptr(), self.coarse_u0.data_ptr()
            self._smr_ck(L.ak_smr_restrict_cc(self.smr, u, cu, self.nvar, st))
            self._smr_ck(L.ak_smr_exchange_cc(self.smr, u, cu, self.nvar, st))
            if self.bcs is None:
                self._smr_ck(L.ak_smr_prolongate_cc(self.smr, u, cu, self.nvar, st))
            else:       # Prolongate (hydro_tasks.cpp:378-398): coarse-array BCs between the coarse fill and the prolongation
                self._smr_ck(L.ak_smr_prolongate_cc_steps(self.smr, u, cu, self.nvar, 1, st))
                if L.ak_bcs_hydro(self.bcs_coarse, cu, self.nvar, st):
                    raise _lib.EngineError(L.ak_bcs_last_error(self.bcs_coarse).decode())
                self._smr_ck(L.ak_smr_prolongate_cc_steps(self.smr, u, cu, self.nvar, 2, st))
                self.launches += self.nbc
            self.launches += 4
            return TaskStatus.complete
        reqs = []
        if self.peers:
            import torch.distributed as dist
            ops = []
            for ip, p in enumerate(self.peers):
                _lib.check(L.ak_halo_pack_cc(self.plan, self.u0.data_ptr(), self.nvar, ip, p["scc"].data_ptr(), st))
                self.launches += 1
                ops.append(dist.P2POp(dist.isend, p["scc"], p["rank"], group=self.pg))
                ops.append(dist.P2POp(dist.irecv, p["rcc"], p["rank"], group=self.pg))
            reqs = dist.batch_isend_irecv(ops)
        _lib.check(L.ak_halo_local_cc(self.plan, self.u0.data_ptr(), self.nvar, st))
        self.launches += 1
        for rq in reqs:
            rq.wait()
        for ip, p in enumerate(self.peers):
            _lib.check(L.ak_halo_unpack_cc(self.plan, self.u0.data_ptr(), self.nvar, ip, p["rcc"].data_ptr(), st))
            self.launches += 1
        return TaskStatus.complete

    def HistorySums(self):
        """History-output sums over this pack's active cells (src/outputs/history.cpp:78-160,272-372):
        mass, 1/2/3-mom, tot-E, 1/2/3-KE (+ 1/2/3-ME for MHD)."""
        out, n = (C.c_double * 11)(), C.c_int()
        self._ck(self.L.ak_hst_sums(self.h, C.byref(self.arrays), out, C.byref(n), self._st()))
        self.launches += 2
        return np.array(out[:n.value])

    def SendRecvB(self, driver, stage):     # no magnetic field: lets Driver.Initialize treat both modules alike
        return TaskStatus.complete

    def ApplyPhysicalBCs(self, driver, stage):
        """hydro_tasks.cpp:357-371: HydroBCs on u0; no-op when strictly periodic."""
        if self.bcs is None:
            return TaskStatus.complete
        if self.L.ak_bcs_hydro(self.bcs, self.u0.data_ptr(), self.nvar, self._st()):
            raise _lib.EngineError(self.L.ak_bcs_last_error(self.bcs).decode())
        self.launches += self.nbc
        return TaskStatus.complete

    def ConToPrim(self, driver, stage, count=True):
        n1, n2, n3 = self.n
        nf = (C.c_int * 3)() if count else None
        self._ck(self.L.ak_hydro_cons2prim(self.h, C.byref(self.arrays), 0, n1 - 1, 0, n2 - 1, 0, n3 - 1, nf, self._st()))
        self.launches += 1
        if count:
            for q in range(3):
                self.nfloor[q] += nf[q]
        return TaskStatus.complete

    def _set_dtnew(self, dt3):
        # hydro_newdt.cpp:122-125
        self.dtnew = dt3[0]
        if self.multi_d:
            self.dtnew = min(self.dtnew, dt3[1])
        if self.three_d:
            self.dtnew = min(self.dtnew, dt3[2])

    def NewTimeStep(self, driver, stage):
        if stage != driver.nexp_stages:      # hydro_newdt.cpp:32-34
            return TaskStatus.complete
        dt3 = (C.c_double * 3)()
        self._ck(self.L.ak_hydro_new_dt(self.h, C.byref(self.arrays), dt3, self._st()))
        self.launches += 1
        self._set_dtnew(dt3)
        return TaskStatus.complete

    def ConToPrimNewDt(self, driver, stage):
        if stage != driver.nexp_stages:
            return self.ConToPrim(driver, stage, count=False)
        nf, dt3 = (C.c_int * 3)(), (C.c_double * 3)()
        self._ck(self.L.ak_hydro_cons2prim_newdt(self.h, C.byref(self.arrays), nf, dt3, self._st()))
        self.launches += 1
        for q in range(3):
            self.nfloor[q] += nf[q]
        self._set_dtnew(dt3)
        return TaskStatus.complete


def assemble_hydro_tasks(hyd: Hydro, fused=True):
    """Hydro::AssembleHydroTasks (hydro_tasks.cpp:47-79) restricted to the tasks that do work on a
    uniform mesh without source terms; `fused` uses the fused ABI calls."""
    from .driver import TaskList
    tl = {n: TaskList() for n in ("before_timeintegrator", "before_stagen", "stagen",
                                  "after_stagen", "after_timeintegrator")}
    s = tl["stagen"]
    if hyd.smr is not None:     # flux correction sits between Fluxes and RKUpdate: no fused flux+update call
        t = s.AddTask(hyd.CopyCons, 0)
        t = s.AddTask(hyd.Fluxes, t)
        t = s.AddTask(hyd.FluxCorrect, t)
        t = s.AddTask(hyd.RKUpdate, t)
        t = s.AddTask(hyd.SendRecvU, t)
        t = s.AddTask(hyd.ApplyPhysicalBCs, t)
        if fused:
            t = s.AddTask(hyd.ConToPrimNewDt, t)
        else:
            t = s.AddTask(hyd.ConToPrim, t)
            t = s.AddTask(hyd.NewTimeStep, t)
    elif fused:
        t = s.AddTask(hyd.StageFused, 0)
        t = s.AddTask(hyd.SendRecvU, t)
        t = s.AddTask(hyd.ApplyPhysicalBCs, t)
        t = s.AddTask(hyd.ConToPrimNewDt, t)
    else:
        t = s.AddTask(hyd.CopyCons, 0)
        t = s.AddTask(hyd.Fluxes, t)
        t = s.AddTask(hyd.RKUpdate, t)
        t = s.AddTask(hyd.SendRecvU, t)
        t = s.AddTask(hyd.ApplyPhysicalBCs, t)
        t = s.AddTask(hyd.ConToPrim, t)
        t = s.AddTask(hyd.NewTimeStep, t)
    return tl
