"""CPU tests of the isothermal-EOS restatement: oracle/mhd_oracle.c and
oracle/hydro_oracle.c with `is_iso` against committed fixtures made from the unmodified reference
(tests/golden/*_iso_*.npz, tests/gen_golden.py iso()), and the C ABI's handling of the EOS selector."""
import ctypes as C

import numpy as np
import pytest

import golden_io as G
import hydro_util as H
import iso_util as I
import oracle_lib as O


@pytest.mark.parametrize("name", sorted(I.CASES))
def test_iso_oracle_matches_golden(name):
    d = G.load(name)
    case = H.Case(str(d["deck"]))
    assert case.iso == ("_iso_" in name) and case.fofc == ("_fofc_" in name) and d["u0_c0"].shape[1] == case.nvar
    if case.physics == "hydro":
        u, w, t, dt, n, _ = H.oracle_run(case, d["u0_c0"])
        b = ()
    else:
        A, t, dt, n, _ = H.oracle_run_mhd(case, d["u0_c0"], (d["b1f_c0"], d["b2f_c0"], d["b3f_c0"]), ncycles=int(d["ncycle_final"]))
        u, b = A.u0, A.b0
    assert n == int(d["ncycle_final"]) and t == float(d["time_final"])
    if case.fofc:
        assert O.lib().oracle_get_nfofc(1) == int(d["nfofc"]) > 0      # the correction fired, as often as in the reference's event log
    assert np.array_equal(u, d["u0_final"]), f"max|dU| = {np.abs(u - d['u0_final']).max()}"
    for x, k in zip(b, ("b1f_final", "b2f_final", "b3f_final")):
        assert np.array_equal(x, d[k])


def test_iso_cfg_validation(engine_lib):
    """ak_*_create argument checks for the EOS selector (no device needed: they fail before any CUDA call)."""
    from athenak_b200 import lib as L
    h = C.c_void_p()
    bad_eos = L.MhdCfg(1, 4, 16, 16, 16, 2, 2, 1.6666667, 0, 0, 0, 0, 1e38, 0, 7, 1.0)
    assert engine_lib.ak_mhd_create(C.byref(bad_eos), C.byref(h)) != 0
    assert b"eos" in engine_lib.ak_last_error(None)
    # hydro.cpp:252-262: hllc cannot be used with the isothermal EOS
    hllc_iso = L.MhdCfg(1, 4, 16, 16, 16, 2, L.HYD_RSOLVER["hllc"], 1.4, 0, 0, 0, 0, 1e38, 0, L.EOS["isothermal"], 1.0)
    assert engine_lib.ak_hydro_create(C.byref(hllc_iso), C.byref(h)) != 0
    assert b"hllc" in engine_lib.ak_last_error(None)
    no_cs = L.MhdCfg(1, 4, 16, 16, 16, 2, 2, 0.0, 0, 0, 0, 0, 1e38, 0, L.EOS["isothermal"], 0.0)
    assert engine_lib.ak_mhd_create(C.byref(no_cs), C.byref(h)) != 0
    assert b"iso_sound_speed" in engine_lib.ak_last_error(None)
