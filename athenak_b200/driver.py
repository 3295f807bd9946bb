"""Mirror of the reference's TaskList / Driver operator API for the MHD path.

Reference: src/tasklist/task_list.hpp:37-236 (TaskID bitmask dependencies, DoAvailable),
src/driver/driver.cpp:93-130 (RK tables), :290-307 (ExecuteTaskList), :380-459 (Execute),
src/mesh/mesh.cpp:573-643 (Mesh::NewTimeStep).  The task bodies are the C-ABI calls in
athenak_b200.mhd.MHD; this file only sequences them the way the reference does.
"""
from __future__ import annotations

import numpy as np

from .mhd import FLT_MAX, MHD, TaskStatus


class TaskList:
    """Bitmask-dependency task list (task_list.hpp:121-236)."""

    def __init__(self):
        self.tasks = []          # [id_bit, dep_mask, func, complete]
        self.completed = 0

    def AddTask(self, func, dep=0):
        tid = 1 << len(self.tasks)
        if len(self.tasks) >= 64:
            raise RuntimeError("TaskList: more than 64 tasks (NUMBER_TASKID_BITS)")
        self.tasks.append([tid, dep, func, False])
        return tid

    def Reset(self):
        self.completed = 0
        for t in self.tasks:
            t[3] = False

    def Empty(self):
        return not self.tasks

    def IsComplete(self):
        return all(self.completed & t[0] for t in self.tasks)

    def DoAvailable(self, driver, stage):
        for t in self.tasks:
            if (self.completed & t[1]) == t[1] and not t[3]:
                if t[2](driver, stage) == TaskStatus.complete:
                    t[3] = True
                    self.completed |= t[0]
        return self.IsComplete()


def assemble_mhd_tasks(mhd: MHD, fused=True):
    """MHD::AssembleMHDTasks (mhd_tasks.cpp:38-84) restricted to the tasks that do work on a
    uniform mesh; `fused` swaps tasks 1,2,14 / 5,17 / 27,28 for the fused ABI calls."""
    tl = {n: TaskList() for n in ("before_timeintegrator", "before_stagen", "stagen",
                                  "after_stagen", "after_timeintegrator")}
    s = tl["stagen"]
    if getattr(mhd, "smr", None) is not None:
        # multilevel mesh (mhd_tasks.cpp:48-75): flux correction sits between Fluxes and the corner EMFs (CornerE
        # upwinds with the CORRECTED mass fluxes), so the fused flux+EMF call does not apply.  `fused`: the corner
        # EMFs and their correction move ahead of the update (they read w0 / bcc0 / fluxes only), then one
        # kernel does CopyCons + RKUpdate + CT (ak_mhd_stage_update); the new time step rides on the last c2p.
        if fused and not mhd.fofc:      # FOFC estimates the update from u1 / b1 over the whole arrays: it needs CopyCons first
            t = s.AddTask(mhd.Fluxes, 0)
            t = s.AddTask(mhd.FluxCorrect, t)
            t = s.AddTask(mhd.EField, t)
            t = s.AddTask(mhd.SendRecvE, t)
            t = s.AddTask(mhd.StageUpdate, t)
            t = s.AddTask(mhd.SendRecvU, t)
            t = s.AddTask(mhd.SendRecvB, t)
            t = s.AddTask(mhd.Prolongate, t)
            t = s.AddTask(mhd.ApplyPhysicalBCs, t)
            t = s.AddTask(mhd.ConToPrimNewDt, t)
        else:
            t = s.AddTask(mhd.CopyCons, 0)
            t = s.AddTask(mhd.Fluxes, t)
            t = s.AddTask(mhd.FluxCorrect, t)
            t = s.AddTask(mhd.RKUpdate, t)
            t = s.AddTask(mhd.SendRecvU, t)
            t = s.AddTask(mhd.EField, t)
            t = s.AddTask(mhd.SendRecvE, t)
            t = s.AddTask(mhd.CT, t)
            t = s.AddTask(mhd.SendRecvB, t)
            t = s.AddTask(mhd.Prolongate, t)
            t = s.AddTask(mhd.ApplyPhysicalBCs, t)
            if fused:
                t = s.AddTask(mhd.ConToPrimNewDt, t)
            else:
                t = s.AddTask(mhd.ConToPrim, t)
                t = s.AddTask(mhd.NewTimeStep, t)
    elif fused:
        t = s.AddTask(mhd.StageFluxEmf, 0)
        t = s.AddTask(mhd.SendRecvE, t)
        t = s.AddTask(mhd.StageUpdate, t)
        t = s.AddTask(mhd.SendRecvU, t)
        t = s.AddTask(mhd.SendRecvB, t)
        t = s.AddTask(mhd.ApplyPhysicalBCs, t)
        t = s.AddTask(mhd.ConToPrimNewDt, t)
    else:
        t = s.AddTask(mhd.CopyCons, 0)
        t = s.AddTask(mhd.Fluxes, t)
        t = s.AddTask(mhd.RKUpdate, t)
        t = s.AddTask(mhd.SendRecvU, t)
        t = s.AddTask(mhd.EField, t)
        t = s.AddTask(mhd.SendRecvE, t)
        t = s.AddTask(mhd.CT, t)
        t = s.AddTask(mhd.SendRecvB, t)
        t = s.AddTask(mhd.ApplyPhysicalBCs, t)
        t = s.AddTask(mhd.ConToPrim, t)
        t = s.AddTask(mhd.NewTimeStep, t)
    return tl


class Driver:
    """Time-integration driver (driver.cpp:60-282, 380-459)."""

    def __init__(self, integrator="rk3", cfl_number=0.3, tlim=1.0, nlim=-1):
        if integrator == "rk1":
            self.gam0, self.gam1, self.beta = [0.0], [1.0], [1.0]
        elif integrator == "rk2":
            self.gam0, self.gam1, self.beta = [0.0, 0.5], [1.0, 0.5], [1.0, 0.5]
        elif integrator == "rk3":
            self.gam0, self.gam1, self.beta = [0.0, 0.25, 2.0 / 3.0], [1.0, 0.75, 1.0 / 3.0], [1.0, 0.25, 2.0 / 3.0]
        else:
            raise ValueError(f"integrator {integrator!r} not supported on this path")
        self.nexp_stages = len(self.beta)
        self.cfl_no, self.tlim, self.nlim = float(cfl_number), float(tlim), int(nlim)
        self.time, self.dt, self.ncycle = 0.0, FLT_MAX, 0
        self.nmb_updated = 0

    def ExecuteTaskList(self, tl, name, stage):
        lst = tl[name]
        if lst.Empty():
            return
        lst.Reset()
        while not lst.DoAvailable(self, stage):
            pass

    def NewTimeStep(self, mhd: MHD):
        """Mesh::NewTimeStep (mesh.cpp:573-643)."""
        dt = 2.0 * self.dt
        dt = min(dt, self.cfl_no * mhd.dtnew)
        if mhd.pg is not None or _dist_ready():
            import torch
            import torch.distributed as dist
            t = torch.tensor([dt], dtype=torch.float64, device=mhd.device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=mhd.pg)
            dt = float(t.item())
        if self.time < self.tlim and (self.time + dt) > self.tlim:
            dt = self.tlim - self.time
        self.dt = dt

    def Initialize(self, mhd: MHD):
        """Driver::Initialize -> InitBoundaryValuesAndPrimitives (driver.cpp:314-371,569-653):
        fill ghost zones of U and B, c2p everywhere, first time step."""
        mhd.SendRecvU(self, 0)
        mhd.SendRecvB(self, 0)
        if getattr(mhd, "smr", None) is not None and hasattr(mhd, "Prolongate"):     # driver.cpp:607-653 on a multilevel mesh
            mhd.Prolongate(self, 0)
        mhd.ApplyPhysicalBCs(self, 0)
        mhd.ConToPrim(self, 0)
        save = self.nexp_stages
        mhd.NewTimeStep(self, save)
        self.NewTimeStep(mhd)

    def Execute(self, mhd: MHD, tl, ncycles=None):
        done = 0
        while self.time < self.tlim and (self.nlim < 0 or self.ncycle < self.nlim):
            if ncycles is not None and done >= ncycles:
                break
            self.ExecuteTaskList(tl, "before_timeintegrator", 0)
            for stage in range(1, self.nexp_stages + 1):
                self.ExecuteTaskList(tl, "before_stagen", stage)
                self.ExecuteTaskList(tl, "stagen", stage)
                self.ExecuteTaskList(tl, "after_stagen", stage)
            self.ExecuteTaskList(tl, "after_timeintegrator", 1)
            self.time += self.dt
            self.ncycle += 1
            done += 1
            self.nmb_updated += mhd.nmb
            self.NewTimeStep(mhd)
        return done


def _dist_ready():
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        return False
