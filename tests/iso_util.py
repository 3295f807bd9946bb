"""Isothermal-EOS and first-order-flux-correction (FOFC) parity cases.
TEST INFRASTRUCTURE ONLY.

Two kinds of synthetic input, both produced by the UNMODIFIED reference (oracle/_ref/athena_builtin):
  * "pgen":    a deck whose problem generator supports the isothermal EOS (linear_wave), run as is;
  * "restart": the shock-tube generators write w0(IEN) unconditionally, so for strong-shock cases the
               ideal-gas initial condition is dumped (nlim=0), its energy variable is dropped, the
               embedded deck is switched to `eos = isothermal`, dt is reset to FLT_MAX (the value a
               fresh Mesh starts from, mesh.cpp:573-580) and the reference is restarted from that
               file (`athena -r`), exactly as a user would restart it from a dump of this engine.
"""
from __future__ import annotations

import dataclasses
import os
import re
import subprocess
import tempfile

import numpy as np

import oracle_lib as O
from athenak_b200.rst import read_rst, write_rst

HERE = os.path.dirname(os.path.abspath(__file__))
DECKS = os.path.join(HERE, "decks")

# name -> (kind, deck, basename, physics, overrides)
CASES = {
    "hydro_iso_lwave3d_ppm4_roe_rk3": ("pgen", "lwave3d_hydro.athinput", "LinWave", "hydro",
        ("mesh/nx1=16", "mesh/nx2=8", "mesh/nx3=8", "meshblock/nx1=8", "meshblock/nx2=8", "meshblock/nx3=4",
         "problem/amp=0.2", "time/tlim=0.4", "hydro/eos=isothermal", "hydro/rsolver=roe", "hydro/reconstruct=ppm4")),
    "mhd_iso_lwave3d_ppm4_hlld_rk3": ("pgen", "lwave3d.athinput", "LinWave", "mhd",
        ("mesh/nx1=16", "mesh/nx2=8", "mesh/nx3=8", "meshblock/nx1=8", "meshblock/nx2=8", "meshblock/nx3=4",
         "problem/amp=0.2", "time/tlim=0.4", "mhd/eos=isothermal", "mhd/rsolver=hlld", "mhd/reconstruct=ppm4")),
    "mhd_iso_bw1d_ppm4_hlld_rk3": ("restart", "bw1d.athinput", "BrioWu", "mhd",
        ("mhd/rsolver=hlld", "mhd/reconstruct=ppm4", "mesh/nghost=3", "time/integrator=rk3", "mesh/nx1=128",
         "meshblock/nx1=32", "time/tlim=0.08")),
    "mhd_iso_bw2d_wenoz_hlle_rk2": ("restart", "bw2d.athinput", "BrioWu", "mhd",
        ("mhd/rsolver=hlle", "mhd/reconstruct=wenoz", "time/integrator=rk2", "time/tlim=0.03")),
    "hydro_iso_sod2d_ppm4_roe_rk3": ("restart", "sod2d.athinput", "Sod", "hydro",
        ("hydro/rsolver=roe", "hydro/reconstruct=ppm4", "time/tlim=0.1")),
}
# FOFC (<mhd>/fofc = true): cases in which the correction actually fires (the reference's event log counts
# 230 flagged cells in the first, ~14000 in the second, where dfloor = 0.2 exceeds the right-hand density)
_STRONG2D = ("problem/pr=1.0e-7", "problem/dr=1.0e-3", "problem/vl=20.0", "problem/vr=-20.0")
CASES.update({
    "mhd_fofc_bw2d_ppm4_hlld_rk3": ("pgen", "bw2d.athinput", "BrioWu", "mhd",
        _STRONG2D + ("mhd/fofc=true", "mesh/nghost=4", "time/tlim=0.002")),
    "mhd_iso_fofc_bw2d_ppm4_hlld_rk3": ("restart", "bw2d.athinput", "BrioWu", "mhd",
        ("mhd/fofc=true", "mesh/nghost=4", "time/tlim=0.05", "time/nlim=12"), (("dfloor", "0.2"),)),
})
# more combinations, run only against the live reference (GPU tests)
LIVE_CASES = dict(CASES)
LIVE_CASES.update({
    "hydro_iso_sod1d_plm_llf_rk2": ("restart", "sod1d.athinput", "Sod", "hydro", ("time/tlim=0.1",)),
    "hydro_iso_sod2d_teno_hlle_rk3": ("restart", "sod2d.athinput", "Sod", "hydro",
        ("hydro/rsolver=hlle", "hydro/reconstruct=teno", "time/tlim=0.05")),
    "hydro_iso_sod2d_scalars": ("restart", "sod2d.athinput", "Sod", "hydro",
        ("hydro/rsolver=llf", "hydro/reconstruct=ppmx", "hydro/nscalars=2", "time/tlim=0.05")),
    "mhd_iso_bw1d_plm_llf_rk2": ("restart", "bw1d.athinput", "BrioWu", "mhd", ("mhd/rsolver=llf", "time/tlim=0.05")),
    "mhd_iso_bw2d_ppm4_hlld_scalars": ("restart", "bw2d.athinput", "BrioWu", "mhd", ("mhd/nscalars=1", "time/tlim=0.02")),
    "mhd_fofc_blast3d_ppm4_hlld_rk3": ("pgen", "blast3d.athinput", "Blast", "mhd",
        ("mesh/nx1=32", "mesh/nx2=32", "mesh/nx3=32", "meshblock/nx1=16", "meshblock/nx2=16", "meshblock/nx3=16", "mhd/fofc=true",
         "time/nlim=6", "problem/prat=100000", "problem/pi_amb=0.0001", "problem/b_amb=3.0")),
    "mhd_fofc_bw2d_plm_llf_rk2": ("pgen", "bw2d.athinput", "BrioWu", "mhd",
        _STRONG2D + ("mhd/fofc=true", "mesh/nghost=3", "mhd/reconstruct=plm", "mhd/rsolver=llf", "time/integrator=rk2", "time/tlim=0.002")),
    "mhd_fofc_bw1d_wenoz_hlle_rk3": ("pgen", "bw1d.athinput", "BrioWu", "mhd",
        ("problem/pr=1.0e-7", "problem/dr=1.0e-3", "problem/ul=20.0", "problem/ur=-20.0", "mhd/fofc=true", "mesh/nghost=4",
         "mhd/rsolver=hlle", "mhd/reconstruct=wenoz", "time/integrator=rk3", "mesh/nx1=128", "meshblock/nx1=32", "time/tlim=0.004")),
    "mhd_iso_fofc_bw2d_scalars": ("restart", "bw2d.athinput", "BrioWu", "mhd",
        ("problem/ul=-100.0", "problem/ur=100.0", "problem/vl=-100.0", "problem/vr=100.0", "mhd/fofc=true", "mhd/nscalars=1",
         "mesh/nghost=4", "time/tlim=0.05", "time/nlim=8"), (("dfloor", "0.2"),)),
    "mhd_iso_lwave3d_wenoz_hlle_rk3": ("pgen", "lwave3d.athinput", "LinWave", "mhd",
        ("mesh/nx1=16", "mesh/nx2=8", "mesh/nx3=8", "meshblock/nx1=8", "meshblock/nx2=4", "meshblock/nx3=4",
         "problem/amp=0.2", "time/tlim=0.3", "mhd/eos=isothermal", "mhd/rsolver=hlle", "mhd/reconstruct=wenoz")),
})

FLT_MAX = float(np.finfo(np.float32).max)


def _nscal(overrides, physics):
    return max([int(o.split("=")[1]) for o in overrides if o.startswith(physics + "/nscalars=")] + [0])


def _nfofc(workdir, base):
    """Cells flagged by FOFC over the run = sum of the last column of the reference's event log
    (src/outputs/eventlog.cpp:95-121; counters are zeroed after every row)."""
    path = os.path.join(workdir, base + ".log")
    if not os.path.exists(path):
        return 0
    return int(sum(int(ln.split()[-1]) for ln in open(path) if ln.strip() and not ln.startswith("#")))


def reference(name, threads=4):
    """Runs the unmodified reference on LIVE_CASES[name]; returns (r0, r1): the initial and final dumps
    (r1.nfofc = FOFC cell count from its event log)."""
    kind, deck, base, physics, ov = LIVE_CASES[name][:5]
    deck_edits = LIVE_CASES[name][5] if len(LIVE_CASES[name]) > 5 else ()     # (key, value) pairs rewritten in the restart deck
    ns = _nscal(ov, physics)
    nfl = 4 if (kind == "restart" or any(o.endswith("/eos=isothermal") for o in ov)) else 5
    with tempfile.TemporaryDirectory() as tmp:
        if kind == "pgen":
            blast = deck.startswith("blast")      # its own binary, one dump per cycle
            O.run_reference("athena_blast" if blast else "athena_builtin", os.path.join(DECKS, deck), tmp, list(ov), threads=threads)
            last = max(int(o.split("=")[1]) for o in ov if o.startswith("time/nlim=")) if blast else 1
            r0 = read_rst(os.path.join(tmp, "rst", base + ".00000.rst"), nmhd=nfl + ns, physics=physics)
            r1 = read_rst(os.path.join(tmp, "rst", base + ".%05d.rst" % last), nmhd=nfl + ns, physics=physics)
            r1.nfofc = _nfofc(tmp, base)
            return r0, r1
        # ideal-gas initial condition -> isothermal restart file
        O.run_reference("athena_builtin", os.path.join(DECKS, deck), tmp, list(ov) + ["time/nlim=0"], threads=threads)
        ri = read_rst(os.path.join(tmp, "rst", base + ".00000.rst"), nmhd=5 + ns, physics=physics)
        deck_txt, n = re.subn(r"(eos\s*=\s*)ideal", lambda m: m.group(1) + "isothermal", ri.deck, count=1)
        assert n == 1
        for key, val in deck_edits:     # e.g. the floors, which the dump lists as "added at run time" defaults
            deck_txt, n = re.subn(r"(\n%s\s*=\s*)\S+" % key, lambda m: m.group(1) + val, deck_txt, count=1)
            assert n == 1
        keep = [0, 1, 2, 3] + list(range(5, 5 + ns))       # drop IEN
        r0 = dataclasses.replace(ri, deck=deck_txt, u0=np.ascontiguousarray(ri.u0[:, keep]), dt=FLT_MAX)
        work = os.path.join(tmp, "restart")
        os.makedirs(work)
        start = os.path.join(work, "start.rst")
        write_rst(start, r0)
        exe = os.path.join(O.REF_DIR, "athena_builtin")
        env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="false")
        nlim = [o for o in ov if o.startswith("time/nlim=")] or ["time/nlim=-1"]     # the dump carries nlim=0
        out = subprocess.run([exe, "-r", start, nlim[-1]], cwd=work, env=env, capture_output=True, text=True, timeout=600)
        if out.returncode != 0:
            raise RuntimeError(f"reference restart failed:\n{out.stdout[-2000:]}\n{out.stderr[-2000:]}")
        r1 = read_rst(os.path.join(work, "rst", base + ".00001.rst"), nmhd=4 + ns, physics=physics)
        r1.nfofc = _nfofc(work, base)
        return r0, r1
