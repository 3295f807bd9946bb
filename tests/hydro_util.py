"""Helpers of the hydro parity tests.  TEST INFRASTRUCTURE ONLY.

A hydro case is defined by a restart dump of the unmodified reference (oracle/_ref/athena_builtin
run live, or a committed fixture made from one by tests/gen_golden.py): its embedded input deck
gives the mesh, boundary flags, reconstruction, Riemann solver, integrator and CFL number; its u0
is the synthetic input.  `oracle_run` evolves it with oracle/hydro_oracle.c."""
from __future__ import annotations

import ctypes as C
import re

import numpy as np

import oracle_lib as O

INTEGRATOR = {"rk1": 1, "rk2": 2, "rk3": 3}


def deck_params(deck: str) -> dict:
    """{'block/key': value} of an athinput text (parameter_input.cpp:120-230: '<block>', 'key = value # comment')."""
    out, blk = {}, None
    for ln in deck.splitlines():
        ln = ln.split("#")[0].strip()
        m = re.match(r"<(\w+)>", ln)
        if m:
            blk = m.group(1)
        elif "=" in ln and blk:
            k, v = ln.split("=", 1)
            out[f"{blk}/{k.strip()}"] = v.strip()
    return out


class Case:
    def __init__(self, deck: str, overrides=()):
        p = deck_params(deck)
        for ov in overrides:
            k, v = ov.split("=", 1)
            p[k] = v
        self.p = p
        self.nmesh = tuple(int(p[f"mesh/nx{d}"]) for d in (1, 2, 3))
        self.nblock = tuple(int(p[f"meshblock/nx{d}"]) for d in (1, 2, 3))
        self.xmin = tuple(float(p[f"mesh/x{d}min"]) for d in (1, 2, 3))
        self.xmax = tuple(float(p[f"mesh/x{d}max"]) for d in (1, 2, 3))
        self.bc_names = tuple(p[f"mesh/{s}x{d}_bc"] for d in (1, 2, 3) for s in ("i", "o"))
        self.mesh_bcs = np.array([O.BC[b] for b in self.bc_names], np.int32)
        self.ng = int(float(p.get("mesh/nghost", 2)))
        self.physics = "mhd" if "mhd/rsolver" in p else "hydro"
        ph = self.physics
        self.recon, self.rsolver = p[ph + "/reconstruct"], p[ph + "/rsolver"]
        self.gamma = float(p[ph + "/gamma"])
        self.dfloor = float(p.get(ph + "/dfloor", O.FLT_MIN))      # eos.cpp:22-25 defaults
        self.nscalars = int(p.get(ph + "/nscalars", 0))
        self.iso = p.get(ph + "/eos", "ideal") == "isothermal"
        self.iso_cs = float(p.get(ph + "/iso_sound_speed", 1.0))
        self.nfl = 4 if self.iso else 5
        self.fofc = p.get(ph + "/fofc", "false").strip().lower() in ("true", "1")
        self.nvar = self.nfl + self.nscalars
        self.integrator = p["time/integrator"]
        self.cfl, self.tlim = float(p["time/cfl_number"]), float(p["time/tlim"])
        periodic = [self.mesh_bcs[2 * d] == O.BC["periodic"] for d in range(3)]
        self.lloc, self.nghbr, self.mb_size, _ = O.uniform_mesh(self.nmesh, self.nblock, self.xmin, self.xmax, periodic)
        self.mb_bcs = O.block_bcs(self.lloc, self.nmesh, self.nblock, self.mesh_bcs)
        self.nmb = len(self.lloc)
        self.strictly_periodic = all(periodic)

    def ocfg(self):
        rs = O.RSOLVER[self.rsolver] if self.physics == "mhd" else O.HYD_RSOLVER[self.rsolver]
        return O.OCfg(self.nmb, self.ng, *self.nblock, O.RECON[self.recon], rs, self.gamma,
                      self.dfloor, O.FLT_MIN, O.FLT_MIN, O.FLT_MIN, O.FLT_MAX, self.nscalars, 1 if self.iso else 0, self.iso_cs,
                      1 if self.fofc else 0)


def oracle_run(case: Case, u_init, ncycles=-1, time0=0.0):
    """Driver::Initialize + Execute on the oracle.  Returns (u0, w0, time, dt_next, ncycles, dt0)."""
    L = O.lib()
    cfg = case.ocfg()
    u0 = np.ascontiguousarray(u_init, dtype=np.float64).copy()
    u1, w0 = np.zeros_like(u0), np.zeros_like(u0)
    if not case.strictly_periodic:
        L.oracle_set_bcs(O.ip(case.mesh_bcs), O.ip(case.mb_bcs))
    try:
        dt0 = L.oracle_hydro_init(C.byref(cfg), O.p(u0), O.p(w0), O.p(case.mb_size), O.ip(case.nghbr),
                                  C.c_double(case.cfl), C.c_double(case.tlim), C.c_double(time0))
        t, dt = C.c_double(time0), C.c_double(dt0)
        n = L.oracle_hydro_run(C.byref(cfg), O.p(u0), O.p(u1), O.p(w0), O.p(case.mb_size), O.ip(case.nghbr),
                               INTEGRATOR[case.integrator], C.c_double(case.cfl), C.c_double(case.tlim), ncycles,
                               C.byref(t), C.byref(dt), None)
    finally:
        L.oracle_set_bcs(None, None)
    return u0, w0, t.value, dt.value, n, dt0


def oracle_run_mhd(case: Case, u_init, b_init, ncycles=-1, time0=0.0):
    """Driver::Initialize + Execute on the MHD oracle (oracle/mhd_oracle.c), 1-D / 2-D / 3-D.
    Returns (Arrays, time, dt_next, ncycles, dt0)."""
    L = O.lib()
    cfg = case.ocfg()
    A = O.Arrays(cfg)
    A.u0[:] = u_init
    for dst, src in zip(A.b0, b_init):
        dst[:] = src
    if not case.strictly_periodic:
        L.oracle_set_bcs(O.ip(case.mesh_bcs), O.ip(case.mb_bcs))
    try:
        L.oracle_halo_cc(C.byref(cfg), O.p(A.u0), case.nvar, O.ip(case.nghbr))
        L.oracle_halo_fc(C.byref(cfg), *[O.p(a) for a in A.b0], O.ip(case.nghbr))
        if not case.strictly_periodic:
            assert L.oracle_apply_bcs(C.byref(cfg), O.ip(case.mesh_bcs), O.ip(case.mb_bcs), O.p(A.u0), *[O.p(a) for a in A.b0]) == 0
        A.c2p_all()
        st = A.state()
        dt0 = L.oracle_initial_dt(C.byref(cfg), C.byref(st), O.p(case.mb_size), C.c_double(case.cfl), C.c_double(case.tlim), C.c_double(time0))
        t, dt = C.c_double(time0), C.c_double(dt0)
        n = L.oracle_run(C.byref(cfg), C.byref(st), O.p(case.mb_size), O.ip(case.nghbr), INTEGRATOR[case.integrator],
                         C.c_double(case.cfl), C.c_double(case.tlim), ncycles if ncycles >= 0 else 1 << 30,
                         C.byref(t), C.byref(dt), None)
    finally:
        L.oracle_set_bcs(None, None)
    return A, t.value, dt.value, n, dt0


def sod_density(x, tlim=0.25):
    """Analytic Sod density as tabulated by the reference's own test
    (tst/test_suite/nr/test_nr_sod_cpu.py:20-46): shock, contact, foot and head of the rarefaction."""
    xs, xc, xf, xh = 1.7522 * tlim, 0.92745 * tlim, -0.07027 * tlim, -1.1832 * tlim
    fan = 0.42632 * (1.0 + 0.20046 * (0.92745 - (0.92745 * (x - xh) / (xf - xh)))) ** 5
    return np.where(x > xs, 0.125, np.where(x > xc, 0.26557, np.where(x > xf, 0.42632, np.where(x > xh, fan, 1.0))))
