#!/usr/bin/env python
"""Side measurement (NOT the headline bench): throughput of the other solver families on the same 256^3 blast
state, one B200, C++ host driver, fused task lists.  Prints one line per variant:
  python profiles/bench_variants.py [--nx 256] [--steps 10]
The headline metric and its contract live in bench.py; this file only documents how the shared sweep kernels
behave when instantiated for other reconstructions / Riemann solvers / EOS / the hydro twin / a refined mesh."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from athenak_b200.mesh import Mesh  # noqa: E402
from athenak_b200.sim import Simulation  # noqa: E402

GAMMA, NG = 1.6666667, 4


ONLY = None


def run(tag, mesh, state, steps, **kw):
    if ONLY and ONLY not in tag:
        return
    sim = Simulation(mesh, 0, ng=NG, gamma=GAMMA, cfl_number=0.3, tlim=1.0e9, fused=True, **kw)
    sim.set_state(*state)
    sim.initialize()
    st = torch.cuda.ExternalStream(sim.status()["stream"])
    sim.execute(3)
    sim.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    sim.execute(steps)
    e1.record(st)
    sim.sync()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ncell = mesh.nmb_total * int(np.prod(mesh.nblock))
    print(f"{tag:44s} {ncell / ms / 1e3:9.1f} Mcell-updates/s  {ms:8.3f} ms/step  nmb={mesh.nmb_total}", flush=True)
    del sim


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--only", default=None, help="run only the variants whose tag contains this text")
    a = ap.parse_args()
    global ONLY
    ONLY = a.only
    n = a.nx
    mesh = Mesh((n, n, n), (n, n, n))
    small = bool(ONLY) and ("SMR" in ONLY or "config 5" in ONLY)      # the refined-mesh lines build their own states
    u0, b1, b2, b3 = mesh.pgen_blast(0, NG, GAMMA) if not small else (torch.zeros(1, 5, 1, 1, 1, dtype=torch.float64),) * 4
    mhd = (u0, b1, b2, b3)
    run("mhd ideal ppm4+hlld rk3 (headline)", mesh, mhd, a.steps, recon="ppm4", rsolver="hlld", integrator="rk3")
    run("mhd ideal plm+hlle rk2", mesh, mhd, a.steps, recon="plm", rsolver="hlle", integrator="rk2")
    run("mhd ideal wenoz+hlld rk3", mesh, mhd, a.steps, recon="wenoz", rsolver="hlld", integrator="rk3")
    run("mhd ideal ppm4+hlld rk3 + 2 scalars", mesh, (torch.cat([u0, 0.5 * u0[:, :1], 0.25 * u0[:, :1]], 1).contiguous(), b1, b2, b3),
        a.steps, recon="ppm4", rsolver="hlld", integrator="rk3", nscalars=2)
    run("mhd ideal ppm4+hlld rk3 + fofc", mesh, mhd, a.steps, recon="ppm4", rsolver="hlld", integrator="rk3", fofc=True)
    iso = (u0[:, :4].contiguous(), b1, b2, b3)
    run("mhd isothermal ppm4+hlld rk3", mesh, iso, a.steps, recon="ppm4", rsolver="hlld", integrator="rk3", eos="isothermal",
        iso_sound_speed=1.0)
    run("hydro ideal ppm4+hllc rk3", mesh, (u0,), a.steps, recon="ppm4", rsolver="hllc", integrator="rk3", physics="hydro")
    run("hydro ideal plm+roe rk2", mesh, (u0,), a.steps, recon="plm", rsolver="roe", integrator="rk2", physics="hydro")
    run("hydro isothermal ppm4+hlle rk3", mesh, (u0[:, :4].contiguous(),), a.steps, recon="ppm4", rsolver="hlle", integrator="rk3",
        physics="hydro", eos="isothermal", iso_sound_speed=1.0)
    # many small blocks: 64 blocks of 64^3 (BASELINE config 3's per-GPU layout)
    m64 = Mesh((n, n, n), (n // 4, n // 4, n // 4))
    run("mhd ideal ppm4+hlld rk3, 64 blocks of (nx/4)^3", m64, m64.pgen_blast(0, NG, GAMMA) if not small else None, a.steps, recon="ppm4", rsolver="hlld",
        integrator="rk3")
    # statically refined hydro mesh: 8 root blocks of (nx/4)^3 ... one octant refined
    q = n // 4
    ms = Mesh((2 * q, 2 * q, 2 * q), (q, q, q), refined_regions=[(1, -0.4, -0.1, -0.4, -0.1, -0.4, -0.1)])
    us = torch.zeros(ms.nmb_total, 5, q + 2 * NG, q + 2 * NG, q + 2 * NG, dtype=torch.float64)
    us[:, 0] = 1.0
    us[:, 4] = 1.5
    us[:, 4, NG + q // 4: NG + q // 2, NG + q // 4: NG + q // 2, NG + q // 4: NG + q // 2] = 15.0       # a hot spot in every block
    run("hydro SMR ppm4+hllc rk3, 2 levels", ms, (us,), a.steps, recon="ppm4", rsolver="hllc", integrator="rk3", physics="hydro",
        multilevel=True)
    # the same refined mesh for MHD: face fields, EMF correction and divergence-preserving prolongation across levels
    ub = us.clone()
    ub[:, 4] += 0.5                      # + B^2/2 of the uniform field B = (1, 0, 0)
    fb = [torch.zeros(ms.nmb_total, q + 2 * NG + (d == 2), q + 2 * NG + (d == 1), q + 2 * NG + (d == 0), dtype=torch.float64) for d in range(3)]
    fb[0][:] = 1.0
    run("mhd SMR ppm4+hlld rk3, 2 levels", ms, (ub, *fb), a.steps, recon="ppm4", rsolver="hlld", integrator="rk3", multilevel=True)

    # BASELINE config 5, static variant on ONE GPU: 3-D MHD blast, three levels, 32^3 blocks (nx root cells per dimension:
    # 1408 blocks = 46 M cells at nx = 256)
    if ONLY and "config 5" in ONLY:
        m5 = Mesh((n, n, n), (32, 32, 32), refined_regions=[(1, -0.25, 0.25, -0.25, 0.25, -0.25, 0.25), (2, -0.125, 0.125, -0.125, 0.125, -0.125, 0.125)])
        run("mhd SMR blast, config 5 static: 3 levels, 32^3 blocks", m5, m5.pgen_blast(0, NG, GAMMA), a.steps, recon="ppm4", rsolver="hlld",
            integrator="rk3", multilevel=True)


if __name__ == "__main__":
    main()
