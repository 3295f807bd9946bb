"""ctypes binding of libathenak_b200.so (the C ABI in include/athenak_b200.h).

The library is the product: there is no Python or CPU implementation behind these calls.
If the shared object is missing it is built in-tree with nvcc (athenak_b200/build.py); if it
cannot be loaded, or no CUDA device is present, the calls raise -- nothing falls back.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

RECON = {"dc": 0, "plm": 1, "ppm4": 2, "ppmx": 3, "wenoz": 4, "teno": 5}
RSOLVER = {"llf": 0, "hlle": 1, "hlld": 2}
HYD_RSOLVER = {"advect": 10, "llf": 11, "hlle": 12, "hllc": 13, "roe": 14}   # AK_RSOLVER_HYD_*

_dp = C.POINTER(C.c_double)


class MhdCfg(C.Structure):
    """ak_mhd_cfg"""
    _fields_ = [("nmb", C.c_int), ("ng", C.c_int), ("nx1", C.c_int), ("nx2", C.c_int),
                ("nx3", C.c_int), ("recon", C.c_int), ("rsolver", C.c_int),
                ("gamma", C.c_double), ("dfloor", C.c_double), ("pfloor", C.c_double),
                ("tfloor", C.c_double), ("sfloor", C.c_double), ("sigma_max", C.c_double), ("nscalars", C.c_int),
                ("eos", C.c_int), ("iso_sound_speed", C.c_double), ("fofc", C.c_int)]


EOS = {"ideal": 0, "isothermal": 1}      # AK_EOS_*


ARRAY_FIELDS = ("u0", "u1", "w0", "bcc0", "b0x1f", "b0x2f", "b0x3f", "b1x1f", "b1x2f", "b1x3f",
                "flx1", "flx2", "flx3", "e3x1", "e2x1", "e1x2", "e3x2", "e2x3", "e1x3",
                "e1", "e2", "e3", "mb_size")


class MhdArrays(C.Structure):
    """ak_mhd_arrays: borrowed device pointers"""
    _fields_ = [(n, C.c_void_p) for n in ARRAY_FIELDS]


class SimCfg(C.Structure):
    """ak_sim_cfg"""
    _fields_ = [("mhd", MhdCfg), ("integrator", C.c_int), ("cfl_number", C.c_double), ("tlim", C.c_double),
                ("nlim", C.c_int), ("my_rank", C.c_int), ("nranks", C.c_int), ("gids_start", C.c_int),
                ("nmb_total", C.c_int), ("fused", C.c_int)]


class EngineError(RuntimeError):
    pass


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path):
        path = _build.build()
    lib = C.CDLL(path)
    lib.ak_last_error.restype = C.c_char_p
    lib.ak_last_error.argtypes = [C.c_void_p]
    vp, ip, d = C.c_void_p, C.c_int, C.c_double
    ap = C.POINTER(MhdArrays)
    sig = {
        "ak_mhd_create": [C.POINTER(MhdCfg), C.POINTER(vp)],
        "ak_mhd_destroy": [vp],
        "ak_mhd_copy_cons": [vp, ap, vp],
        "ak_mhd_fluxes": [vp, ap, vp],
        "ak_mhd_rk_update": [vp, ap, d, d, d, vp],
        "ak_mhd_corner_e": [vp, ap, vp],
        "ak_mhd_ct": [vp, ap, d, d, d, vp],
        "ak_mhd_cons2prim": [vp, ap, ip, ip, ip, ip, ip, ip, C.POINTER(C.c_int), vp],
        "ak_mhd_new_dt": [vp, ap, _dp, vp],
        "ak_mhd_stage_fused": [vp, ap, ip, d, d, d, vp, vp],
        "ak_mhd_stage_flux_emf": [vp, ap, vp],
        "ak_mhd_fofc": [vp, ap, d, d, d, C.POINTER(C.c_int), vp],
        "ak_mhd_set_stage": [vp, d, d, d],
        "ak_mhd_stage_begin": [vp, ip, d, d, d],
        "ak_mhd_aux_stream": [vp, C.POINTER(vp)],
        "ak_mhd_stage_update": [vp, ap, ip, d, d, d, vp],
        "ak_mhd_cons2prim_newdt": [vp, ap, C.POINTER(C.c_int), _dp, vp],
        "ak_hst_sums": [vp, ap, _dp, C.POINTER(C.c_int), vp],
        "ak_halo_push_setup": [vp, ip, ip, vp, vp, vp, vp],
        "ak_halo_push_flags": [vp, ip, ip, vp, vp],
        "ak_halo_push_commit": [vp],
        "ak_halo_push": [vp, vp, ip, vp, vp, vp, vp],
        "ak_halo_signal": [vp, ip, C.c_ulonglong, vp],
        "ak_halo_wait": [vp, ip, C.c_ulonglong, vp],
        "ak_hydro_create": [C.POINTER(MhdCfg), C.POINTER(vp)],
        "ak_hydro_destroy": [vp],
        "ak_hydro_copy_cons": [vp, ap, vp],
        "ak_hydro_fluxes": [vp, ap, vp],
        "ak_hydro_rk_update": [vp, ap, d, d, d, vp],
        "ak_hydro_cons2prim": [vp, ap, ip, ip, ip, ip, ip, ip, C.POINTER(C.c_int), vp],
        "ak_hydro_new_dt": [vp, ap, _dp, vp],
        "ak_hydro_stage_fused": [vp, ap, ip, d, d, d, vp],
        "ak_hydro_cons2prim_newdt": [vp, ap, C.POINTER(C.c_int), _dp, vp],
        "ak_mhd_tail_supported": [vp],
        "ak_mhd_fluxes_update_supported": [vp],
        "ak_mhd_ct_save": [vp, ap, ip, d, d, d, vp],
        "ak_mhd_fluxes_update": [vp, ap, vp, ip, d, d, d, vp],
        "ak_mhd_stage_tail": [vp, ap, ap, ip, d, d, d, ip, vp],
        "ak_mhd_stage_shell_ct": [vp, ap, ap, ip, d, d, d, vp],
        "ak_mhd_stage_shell_c2p": [vp, ap, C.POINTER(C.c_int), _dp, vp],
        "ak_halo_push_set_faces": [vp, ip, vp, vp, vp],
        "ak_halo_create": [C.POINTER(MhdCfg), C.POINTER(C.c_int), ip, ip, C.POINTER(vp)],
        "ak_halo_destroy": [vp],
        "ak_halo_npeers": [vp],
        "ak_halo_peer": [vp, ip, C.POINTER(C.c_int), C.POINTER(C.c_long), C.POINTER(C.c_long)],
        "ak_halo_descs": [vp, ip, ip, C.POINTER(C.c_long), ip],
        "ak_halo_local_cc": [vp, vp, ip, vp],
        "ak_halo_local_fc": [vp, vp, vp, vp, vp],
        "ak_halo_pack_cc": [vp, vp, ip, ip, vp, vp],
        "ak_halo_unpack_cc": [vp, vp, ip, ip, vp, vp],
        "ak_halo_pack_fc": [vp, vp, vp, vp, ip, vp, vp],
        "ak_halo_unpack_fc": [vp, vp, vp, vp, ip, vp, vp],
    }
    ipp, dpp = C.POINTER(C.c_int), _dp
    sig.update({
        "ak_mesh_create": [ipp, ipp, dpp, dpp, ipp, ip, C.POINTER(vp)],
        "ak_mesh_create_smr": [ipp, ipp, dpp, dpp, ipp, ip, ip, dpp, C.POINTER(vp)],
        "ak_mesh_destroy": [vp],
        "ak_mesh_set_amr": [vp, ip, ip, ip],
        "ak_mesh_adapt": [vp, ip, ipp, ipp, ipp],
        "ak_mesh_amr_maps": [vp, ipp, ipp, ipp, ipp],
        "ak_amr_criterion": [C.POINTER(MhdCfg), ip, vp, ip, ip, d, d, dpp, dpp, ipp, vp],
        "ak_amr_plan": [C.POINTER(MhdCfg), ip, ipp, ipp, ipp, ipp],
        "ak_amr_move": [C.POINTER(MhdCfg), ipp, vp, vp, ip, vp, vp, vp, vp, vp, vp, vp],
        "ak_amr_repair_fc": [C.POINTER(MhdCfg), ipp, vp, vp, vp, vp],
        "ak_smr_create": [C.POINTER(MhdCfg), ipp, ipp, ip, ip, C.POINTER(vp)],
        "ak_smr_destroy": [vp],
        "ak_smr_counts": [vp, C.POINTER(C.c_long)],
        "ak_smr_restrict_cc": [vp, vp, vp, ip, vp],
        "ak_smr_exchange_cc": [vp, vp, vp, ip, vp],
        "ak_smr_prolongate_cc": [vp, vp, vp, ip, vp],
        "ak_smr_prolongate_cc_steps": [vp, vp, vp, ip, ip, vp],
        "ak_smr_flux_correct_cc": [vp, vp, vp, vp, ip, vp],
        "ak_smr_restrict_fc": [vp, vp, vp, vp, vp, vp, vp, vp],
        "ak_smr_exchange_fc": [vp, vp, vp, vp, vp, vp, vp, vp],
        "ak_smr_prolongate_fc_steps": [vp, vp, vp, vp, vp, vp, vp, ip, vp],
        "ak_smr_emf_correct": [vp, vp, vp, vp, vp],
        "ak_smr_fc_tables": [vp, ip, C.POINTER(vp), C.POINTER(C.c_long)],
        "ak_mesh_info": [vp, ipp, ipp],
        "ak_mesh_tables": [vp, ipp, ipp, ipp, ipp],
        "ak_mesh_rank_tables": [vp, ip, dpp, ipp, ipp],
        "ak_neighbor_index": [ip, ip, ip, ip, ip],
        "ak_pgen_blast": [vp, ip, ip, d, dpp, dpp, dpp, dpp, dpp],
        "ak_halo_emf_peer": [vp, ip, C.POINTER(C.c_long), C.POINTER(vp)],
        "ak_halo_emf_pack_local": [vp, vp, vp, vp, vp],
        "ak_halo_emf_pack": [vp, vp, vp, vp, ip, vp, vp],
        "ak_halo_emf_sumavg": [vp, vp, vp, vp, vp],
        "ak_pgen_orszag_tang": [vp, ip, ip, d, dpp, dpp, dpp, dpp],
        "ak_selftest_divsqrt": [C.c_long, C.c_ulonglong, ip, C.POINTER(C.c_long)],
        "ak_peak_fp64": [ip, ip, ip, ip, dpp, dpp],
        "ak_sim_create": [C.POINTER(SimCfg), dpp, ipp, C.POINTER(vp)],
        "ak_sim_destroy": [vp],
        "ak_sim_set_levels": [vp, ipp],
        "ak_sim_adapt": [vp, vp, dpp, ip, ipp],
        "ak_sim_execute_amr": [vp, vp, dpp, ip, ip],
        "ak_sim_amr_status": [vp, C.POINTER(C.c_long)],
        "ak_sim_set_time": [vp, d, d, ip],
        "ak_sim_prims_newdt": [vp],
        "ak_sim_mesh_new_dt": [vp],
        "ak_small_d2h_sync": [vp, vp, C.c_long, vp],
        "ak_mhd_set_march_tma": [ip],
        "ak_pipe_create": [C.POINTER(vp), ip, C.POINTER(vp)],
        "ak_pipe_wait": [vp, C.c_long],
        "ak_pipe_streams": [vp, C.POINTER(vp), C.POINTER(vp)],
        "ak_pipe_destroy": [vp],
        "ak_nccl_unique_id": [vp],
        "ak_sim_comm_init": [vp, vp],
        "ak_sim_set_state": [vp, vp, vp, vp, vp],
        "ak_sim_get_state": [vp, vp, vp, vp, vp],
        "ak_sim_get_prim": [vp, vp, vp],
        "ak_sim_initialize": [vp],
        "ak_sim_execute": [vp, ip],
        "ak_sim_sync": [vp],
        "ak_sim_status": [vp, dpp, C.POINTER(C.c_long)],
        "ak_sim_arrays": [vp, C.POINTER(MhdArrays)],
        "ak_sim_set_bcs": [vp, ipp, ipp],
        "ak_bcs_create": [C.POINTER(MhdCfg), ipp, ipp, C.POINTER(vp)],
        "ak_bcs_destroy": [vp],
        "ak_bcs_strictly_periodic": [vp],
        "ak_bcs_hydro": [vp, vp, ip, vp],
        "ak_bcs_bfield": [vp, vp, vp, vp, vp],
        "ak_bcs_apply": [vp, vp, ip, vp, vp, vp, vp],
    })
    lib.ak_bcs_last_error.restype = C.c_char_p
    lib.ak_bcs_last_error.argtypes = [vp]
    lib.ak_mesh_last_error.restype = C.c_char_p
    lib.ak_smr_last_error.restype = C.c_char_p
    lib.ak_smr_last_error.argtypes = [vp]
    lib.ak_smr_coarse_cells.restype = C.c_long
    lib.ak_smr_coarse_cells.argtypes = [vp]
    lib.ak_amr_last_error.restype = C.c_char_p
    lib.ak_amr_last_error.argtypes = []
    lib.ak_smr_coarse_faces.restype = C.c_long
    lib.ak_smr_coarse_faces.argtypes = [vp, ip]
    if hasattr(lib, "ak_pipe_submit"):
        lib.ak_pipe_submit.restype = C.c_long
        lib.ak_pipe_submit.argtypes = [vp] * 9 + [ip]
    lax = bool(os.environ.get("ATHENAK_B200_LIB")) and os.environ.get("AKB_LIB_LAX") == "1"   # tuning variants built from older sources
    for name, args in sig.items():
        if lax and not hasattr(lib, name):
            continue
        fn = getattr(lib, name)      # AttributeError here = header/library mismatch: fail loudly
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = lib
    return lib


EXPORTED = ("ak_mhd_create", "ak_mhd_destroy", "ak_last_error", "ak_mhd_copy_cons", "ak_mhd_fluxes",
            "ak_mhd_rk_update", "ak_mhd_corner_e", "ak_mhd_ct", "ak_mhd_cons2prim", "ak_mhd_new_dt",
            "ak_mhd_stage_fused", "ak_mhd_tail_supported", "ak_mhd_fluxes_update_supported", "ak_mhd_ct_save", "ak_mhd_fluxes_update", "ak_mhd_stage_tail", "ak_mhd_stage_shell_ct", "ak_mhd_stage_shell_c2p", "ak_halo_push_set_faces", "ak_mhd_stage_flux_emf", "ak_mhd_fofc", "ak_mhd_set_stage", "ak_mhd_stage_begin", "ak_mhd_aux_stream", "ak_mhd_stage_update", "ak_mhd_cons2prim_newdt", "ak_halo_create", "ak_halo_destroy",
            "ak_halo_npeers", "ak_halo_peer", "ak_halo_descs", "ak_halo_local_cc", "ak_halo_local_fc",
            "ak_halo_pack_cc", "ak_halo_unpack_cc", "ak_halo_pack_fc", "ak_halo_unpack_fc",
            "ak_halo_emf_peer", "ak_halo_emf_pack_local", "ak_halo_emf_pack", "ak_halo_emf_sumavg",
            "ak_mesh_create", "ak_mesh_create_smr", "ak_mesh_destroy", "ak_mesh_set_amr", "ak_mesh_adapt", "ak_mesh_amr_maps", "ak_amr_last_error", "ak_amr_criterion", "ak_amr_plan", "ak_amr_move", "ak_amr_repair_fc", "ak_smr_create", "ak_smr_destroy", "ak_smr_last_error",
            "ak_smr_coarse_cells", "ak_smr_counts", "ak_smr_restrict_cc", "ak_smr_exchange_cc", "ak_smr_prolongate_cc", "ak_smr_prolongate_cc_steps", "ak_smr_flux_correct_cc", "ak_smr_coarse_faces", "ak_smr_restrict_fc", "ak_smr_exchange_fc", "ak_smr_prolongate_fc_steps", "ak_smr_emf_correct", "ak_smr_fc_tables", "ak_mesh_last_error", "ak_mesh_info", "ak_mesh_tables",
            "ak_mesh_rank_tables", "ak_neighbor_index", "ak_pgen_blast", "ak_pgen_orszag_tang", "ak_selftest_divsqrt", "ak_peak_fp64",
            "ak_sim_create", "ak_sim_destroy", "ak_sim_set_levels", "ak_sim_adapt", "ak_sim_execute_amr", "ak_sim_amr_status", "ak_sim_set_time", "ak_sim_prims_newdt", "ak_sim_mesh_new_dt", "ak_mhd_set_march_tma", "ak_small_d2h_sync", "ak_pipe_create", "ak_pipe_submit", "ak_pipe_wait", "ak_pipe_streams", "ak_pipe_destroy", "ak_nccl_unique_id", "ak_sim_comm_init", "ak_sim_set_state",
            "ak_sim_get_state", "ak_sim_get_prim", "ak_sim_initialize", "ak_sim_execute", "ak_sim_sync", "ak_sim_status", "ak_sim_arrays",
            "ak_sim_set_bcs", "ak_bcs_create", "ak_bcs_destroy", "ak_bcs_last_error", "ak_bcs_strictly_periodic",
            "ak_bcs_hydro", "ak_bcs_bfield", "ak_bcs_apply",
            "ak_hst_sums", "ak_halo_emf_rbuf", "ak_halo_push_setup", "ak_halo_push_flags", "ak_halo_push_commit", "ak_halo_push",
            "ak_halo_signal", "ak_halo_wait", "ak_hydro_create", "ak_hydro_destroy", "ak_hydro_copy_cons", "ak_hydro_fluxes", "ak_hydro_rk_update",
            "ak_hydro_cons2prim", "ak_hydro_new_dt", "ak_hydro_stage_fused", "ak_hydro_cons2prim_newdt")

# BoundaryFlag (src/bvals/bvals.hpp:19) == enum ak_bc_flag
BC = {"undef": -1, "block": 0, "reflect": 1, "inflow": 2, "outflow": 3, "diode": 4, "user": 5, "periodic": 6,
      "shear_periodic": 7, "vacuum": 8}


def check(rc, handle=None):
    if rc != 0:
        msg = load().ak_last_error(handle)
        raise EngineError((msg or b"unknown error").decode())
