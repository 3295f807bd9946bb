"""Side measurement: the pipelined host-buffer path (ak_pipe_*) alone -- slots sweep, optional AKB_PIPE_TRACE=1 timeline."""
import argparse, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from athenak_b200.mesh import Mesh
from athenak_b200.sim import HostPipeline, Simulation
ap = argparse.ArgumentParser(); ap.add_argument("--nx", type=int, default=256); ap.add_argument("--slots", type=int, default=4); ap.add_argument("--jobs", type=int, default=16)
a = ap.parse_args()
n, NI = a.nx, a.slots
mesh = Mesh((n, n, n), (n, n, n))
hin = mesh.pgen_blast(0, 4, 1.6666667, pinned=True)
sims = [Simulation(mesh, 0, ng=4, gamma=1.6666667, cfl_number=0.3, tlim=1e9, fused=True, recon="ppm4", rsolver="hlld", integrator="rk3") for _ in range(NI)]
outs = [[torch.empty_like(t).pin_memory() for t in hin] for _ in range(NI)]
pipe = HostPipeline(sims)
up, down = (torch.cuda.ExternalStream(h) for h in pipe.streams())
def run(nj):
    for j in range(nj):
        pipe.submit(hin, outs[j % NI], 1)
        if j >= NI - 1: pipe.wait(j - (NI - 1))
    pipe.wait(-1)
run(NI); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(up); run(a.jobs); e1.record(down); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.jobs
print(f"slots {NI}: {ms:.2f} ms/step  {n**3 / ms / 1e3:.1f} Mcell-updates/s e2e", flush=True)
pipe.close()
