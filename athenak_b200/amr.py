"""Adaptive mesh refinement around the MHD mirror (reference: MeshRefinement::AdaptiveMeshRefinement,
src/mesh/mesh_refinement.cpp:145-190, called from Driver::Execute after the outputs, src/driver/driver.cpp:447-450).

The host (athenak_b200.mesh.Mesh = the C++ host mesh) owns the tree and the block lists; the device evaluates the
refinement criteria (ak_amr_criterion) and moves the data (ak_amr_plan + ak_amr_move: every element of the NEW arrays is
gathered from the OLD ones -- copy / on-the-fly restriction / direct prolongation -- and ak_amr_repair_fc).  Because the
arrays of the mirror are sized by the number of MeshBlocks, a mesh change builds a new MHD object for the new mesh
(`make_mhd(mesh)`), which is also what re-creates the ghost-exchange and refinement plans for the new neighbour table.

ONE RANK ONLY: the flag array, ak_mesh_adapt, ak_amr_plan and the new-to-old maps all index the GLOBAL block list, and no block
migrates between ranks after the load balance (reference: src/mesh/load_balance.cpp:97-1024, not built); adaptive_mesh_refinement
raises EngineError for a mesh sharded over several ranks."""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

from . import lib as _lib
from .mhd import FLT_MAX

METHOD = {"min_max": 0, "slope": 1, "second_deriv": 2, "location": 3}


@dataclasses.dataclass
class Criterion:      # one <amr_criterion> block (src/mesh/refinement_criteria.cpp:33-71)
    method: str
    variable: str = "mhd_w_d"
    value_min: float = -FLT_MAX
    value_max: float = FLT_MAX
    location: tuple = (0.0, 0.0, 0.0, 0.0)      # location_x1, _x2, _x3, _rad


def criteria_from_params(p):
    out = []
    for blk in sorted({k.split("/")[0] for k in p if k.startswith("amr_criterion")}):
        out.append(Criterion(p[f"{blk}/method"], p.get(f"{blk}/variable", "mhd_w_d"), float(p.get(f"{blk}/value_min", -FLT_MAX)),
                             float(p.get(f"{blk}/value_max", FLT_MAX)),
                             tuple(float(p.get(f"{blk}/location_{q}", 0.0)) for q in ("x1", "x2", "x3", "rad"))))
    return out


def refinement_flags(mhd, mesh, criteria):
    """CheckForRefinement's criteria loop (mesh_refinement.cpp:219-245) on the device."""
    L = mhd.L
    flag = np.zeros(mhd.nmb, np.int32)
    mbs = np.ascontiguousarray(mhd.mb_size_h)
    dp, ipp = C.POINTER(C.c_double), C.POINTER(C.c_int)
    for c in criteria:
        if c.method not in METHOD:
            raise _lib.EngineError(f"athenak_b200: refinement method {c.method!r} not supported")
        if c.method != "location" and c.variable not in ("mhd_w_d", "mhd_u_d", "hydro_w_d", "hydro_u_d"):
            raise _lib.EngineError(f"athenak_b200: refinement variable {c.variable!r} not supported")
        q = mhd.w0 if c.variable.endswith("_w_d") else mhd.u0
        loc = (C.c_double * 4)(*c.location)
        if L.ak_amr_criterion(C.byref(mhd.cfg), METHOD[c.method], q.data_ptr(), mhd.nvar, 0, c.value_min, c.value_max, loc,
                              mbs.ctypes.data_as(dp), flag.ctypes.data_as(ipp), mhd._st()):
            raise _lib.EngineError(L.ak_amr_last_error().decode())
        mhd.launches += 1
    return flag


def init_boundary_values_and_primitives(mhd, driver):
    """Driver::InitBoundaryValuesAndPrimitives (driver.cpp:569-653) for the MHD mirror, or for the hydro twin (whose SendRecvU
    already restricts, exchanges and prolongates)."""
    mhd.SendRecvU(driver, 0)
    if hasattr(mhd, "Prolongate"):
        mhd.SendRecvB(driver, 0)
        mhd.Prolongate(driver, 0)
    mhd.ApplyPhysicalBCs(driver, 0)
    mhd.ConToPrim(driver, 0)


def adaptive_mesh_refinement(mhd, mesh, driver, criteria, make_mhd):
    """One AdaptiveMeshRefinement call.  Returns (mhd, changed): the same object when the mesh did not change, otherwise
    a new MHD object (from make_mhd(mesh)) holding the moved data, rebuilt ghost zones / primitives and its new dtnew."""
    L = mhd.L
    if getattr(mesh, "nranks", 1) != 1 or mhd.nmb != mesh.nmb_total:
        raise _lib.EngineError("athenak_b200: adaptive mesh refinement is built for one rank holding every MeshBlock "
                               f"(mesh.nranks = {getattr(mesh, 'nranks', 1)}, blocks here {mhd.nmb} of {mesh.nmb_total})")
    # CheckForRefinement returns before evaluating any criterion on a non-check cycle (mesh_refinement.cpp:196-215): zero flags
    # still go through ak_mesh_adapt so that its cycle counters advance
    ncheck = getattr(mesh, "ncycle_check", 1)
    if ncheck > 1 and driver.ncycle % ncheck != 0:
        flag = np.zeros(mhd.nmb, np.int32)
    else:
        flag = refinement_flags(mhd, mesh, criteria)
    old_nmb = mhd.nmb
    nnew, ndel, what = mesh.adapt(driver.ncycle, flag)
    if not (nnew or ndel):
        return mhd, False
    new = make_mhd(mesh)
    ipp = C.POINTER(C.c_int)
    n2o = np.ascontiguousarray(mesh.newtoold, np.int32)
    lloc = np.ascontiguousarray(mesh.lloc, np.int32)
    plan = np.zeros((new.nmb, 5), np.int32)
    if L.ak_amr_plan(C.byref(new.cfg), old_nmb, n2o.ctypes.data_as(ipp), what.ctypes.data_as(ipp), lloc.ctypes.data_as(ipp),
                     plan.ctypes.data_as(ipp)):
        raise _lib.EngineError(L.ak_amr_last_error().decode())
    st = new._st()
    is_mhd = hasattr(mhd, "b0")
    bo = [a.data_ptr() for a in mhd.b0] if is_mhd else [None] * 3
    bn = [a.data_ptr() for a in new.b0] if is_mhd else [None] * 3
    if L.ak_amr_move(C.byref(new.cfg), plan.ctypes.data_as(ipp), mhd.u0.data_ptr(), new.u0.data_ptr(), new.nvar, *bo, *bn, st):
        raise _lib.EngineError(L.ak_amr_last_error().decode())
    new.launches = mhd.launches + (4 if is_mhd else 1)
    new.nfloor = list(mhd.nfloor)
    if hasattr(mhd, "nfofc"):
        new.nfofc = mhd.nfofc              # EventCounters accumulate over the whole run (src/mesh/mesh.hpp:70-74)
    init_boundary_values_and_primitives(new, driver)
    if is_mhd:       # mesh_refinement.cpp:164-171: only MHD repairs the interior faces and rebuilds the ghost zones a second time
        repair = np.ascontiguousarray(what[n2o] != 0, np.int32)
        if L.ak_amr_repair_fc(C.byref(new.cfg), repair.ctypes.data_as(ipp), *bn, st):
            raise _lib.EngineError(L.ak_amr_last_error().decode())
        new.launches += 1
        init_boundary_values_and_primitives(new, driver)
    new.NewTimeStep(driver, driver.nexp_stages)
    return new, True


def _assemble(obj, fused):
    if hasattr(obj, "b0"):
        from .driver import assemble_mhd_tasks
        return assemble_mhd_tasks(obj, fused=fused)
    from .hydro import assemble_hydro_tasks
    return assemble_hydro_tasks(obj, fused=fused)


def execute(mhd, mesh, driver, criteria, make_mhd, ncycles=None, fused=True):
    """Driver::Execute with AMR (driver.cpp:380-459): cycle, then AdaptiveMeshRefinement, then Mesh::NewTimeStep.  Returns the
    MHD object of the final mesh and the number of mesh changes."""
    tl = _assemble(mhd, fused)
    done = nchanged = 0
    while driver.time < driver.tlim and (driver.nlim < 0 or driver.ncycle < driver.nlim):
        if ncycles is not None and done >= ncycles:
            break
        for stage in range(1, driver.nexp_stages + 1):
            driver.ExecuteTaskList(tl, "stagen", stage)
        driver.time += driver.dt
        driver.ncycle += 1
        driver.nmb_updated += mhd.nmb
        done += 1
        mhd, changed = adaptive_mesh_refinement(mhd, mesh, driver, criteria, make_mhd)
        if changed:
            nchanged += 1
            tl = _assemble(mhd, fused)      # the task bodies are bound to the MHD / Hydro object
        driver.NewTimeStep(mhd)
    return mhd, nchanged

