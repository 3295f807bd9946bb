#!/usr/bin/env python
"""Sweep of the FP64 issue-peak microbenchmark (csrc/peaks.cu, ak_peak_fp64): op x ILP x resident warps per SMSP.
Prints lane-instructions/s in units of 1e12 and the implied warp-instructions per clock per SM sub-partition."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402  (initialises the CUDA context the library launches into)

from athenak_b200 import lib as _lib  # noqa: E402

torch.cuda.init()
torch.zeros(1, device="cuda")
L = _lib.load()
nsm = torch.cuda.get_device_properties(0).multi_processor_count
mhz = 1965.0
names = {0: "DFMA", 1: "DADD", 2: "DMUL", 3: "DADD/DMUL"}
print(f"# SMs {nsm}; warp-instr/clk/SMSP assumes {mhz:.0f} MHz")
print(f"{'op':10s} {'ilp':>3s} {'warps/SMSP':>10s} {'T lane-instr/s':>15s} {'warp-instr/clk/SMSP':>20s} {'ms':>8s}")
best = {}
for op in (0, 1, 2, 3):
    for ilp in (1, 2, 4, 8):
        for w in (1, 2, 4, 6, 8, 12, 16):
            r, ms = C.c_double(), C.c_double()
            iters = 4096 // ilp
            assert L.ak_peak_fp64(op, ilp, w, iters, C.byref(r), C.byref(ms)) == 0
            v = r.value
            best[op] = max(best.get(op, 0.0), v)
            print(f"{names[op]:10s} {ilp:3d} {w:10d} {v / 1e12:15.3f} {v / 32 / (nsm * 4) / (mhz * 1e6):20.3f} {ms.value:8.3f}")
print("# best:", {names[k]: round(v / 1e12, 3) for k, v in best.items()})
