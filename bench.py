#!/usr/bin/env python
"""bench.py -- headline benchmark: Mcell-updates/s on the 3-D MHD blast, 256^3 cells per GPU,
one MeshBlock per GPU, PPM4 + HLLD + RK3, nghost=4 (BASELINE.json configs[1]).

A "step" is one full time step (all three RK stages) of the hot path over the rank's blocks.
  python bench.py [--gpus N] [--steps K] [--warmup W] [--nx 256]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...   # the reference's own CPU build (oracle/_ref), all host threads
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GAMMA = 1.6666667
NG = 4
DECK = os.path.join(ROOT, "tests", "decks", "blast3d.athinput")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "athena_blast")
ALGO_BYTES_PER_CELL_STAGE = 256.0      # 32 FP64 words per active cell per stage
METRIC = "Mcell-updates/s on 3D MHD blast 256^3/GPU"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------------------
# reference CPU arm (unmodified reference binary built by oracle/build_ref.sh)

def run_reference_cpu(n, ncycles, threads, timeout=1500):
    """Runs oracle/_ref/athena_blast at n^3 for ncycles; returns per-cycle wall seconds (from the
    binary's own 'elapsed=' diagnostics) and its zone-cycles/cpu_second line."""
    if not os.path.exists(REF_BIN):
        return None
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="spread", OMP_PLACES="threads")
    ov = [f"mesh/nx{d}={n}" for d in (1, 2, 3)] + [f"meshblock/nx{d}={n}" for d in (1, 2, 3)]
    ov += [f"time/nlim={ncycles}", "output1/dcycle=-1"]
    with tempfile.TemporaryDirectory(prefix="akref_") as tmp:
        p = subprocess.run([REF_BIN, "-i", DECK, *ov], cwd=tmp, env=env, capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError("reference run failed: " + p.stdout[-1500:] + p.stderr[-1500:])
    el = [float(m.group(1)) for m in re.finditer(r"^elapsed=(\S+) cycle=", p.stdout, re.M)]
    zc = re.search(r"zone-cycles/cpu_second = (\S+)", p.stdout)
    per_cycle = [b - a for a, b in zip(el[:-1], el[1:])]
    return dict(per_cycle=per_cycle, zone_cycles_per_s=float(zc.group(1)) if zc else None)


def pick_reference_size(threads, budget_s, ncycles):
    """Bounded sample: largest n in {256,128,64} whose (ncycles) cycles fit the budget, judged from
    a 2-cycle probe at 64^3."""
    probe = run_reference_cpu(64, 2, threads)
    rate = max(probe["zone_cycles_per_s"] or 1.0, 1.0)      # zone-cycles per second
    for n in (256, 128, 64):
        if n ** 3 * ncycles / rate <= budget_s:
            return n, rate
    return 64, rate


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_cores()
    if not os.path.exists(REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/athena_blast not built (run oracle/build_ref.sh where /root/reference exists)"}))
        return
    K, W = args.steps, args.warmup
    n, _ = pick_reference_size(threads, budget_s=150.0, ncycles=K + W)
    r = run_reference_cpu(n, K + W, threads)
    timed = r["per_cycle"][W:W + K] if len(r["per_cycle"]) >= W + K else r["per_cycle"][-K:]
    tsum = sum(timed)
    value = n ** 3 * len(timed) / tsum / 1e6
    sample = f"{n}^3 single MeshBlock (bounded sample of the 256^3 workload), {len(timed)} timed cycles after {W} warm-up, all host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mcell-updates/s", "n_gpus": args.gpus,
        "steps": len(timed), "warmup": W, "ms_per_step": 1e3 * tsum / len(timed), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "3D MHD blast, PPM4+HLLD+RK3, nghost=4, reference Kokkos::OpenMP CPU build", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mcell-updates/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------
# native arm

class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            t0 = time.time()     # nvidia-smi needs ~0.1-1 s to come up: wait for its first line so the
            while not self.lines and time.time() - t0 < 5.0:   # samples cover the short timed region
                time.sleep(0.01)
            self.idle = len(self.lines)
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines[getattr(self, "idle", 0):]:     # drop the samples taken before the load started
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bench_native(args):
    import torch
    import torch.distributed as dist

    from athenak_b200.driver import Driver, assemble_mhd_tasks
    from athenak_b200.mesh import Mesh
    from athenak_b200.mhd import MHD

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.nx
    # weak scaling: one n^3 MeshBlock per GPU; blocks tile (2,2,2)-wise in Z-order, same dx
    nroot = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(world)
    if nroot is None:
        raise SystemExit("bench.py: --gpus must be 1, 2, 4 or 8")
    mesh = Mesh(tuple(n * r for r in nroot), (n, n, n), xmin=tuple(-0.5 * r for r in nroot),
                xmax=tuple(0.5 * r for r in nroot), nranks=world)
    assert int(mesh.nmb_rank[rank]) == 1
    mb_size, _, nghbr = mesh.rank_tables(rank)
    hu0, hb1, hb2, hb3 = mesh.pgen_blast(rank, NG, GAMMA, pinned=True)
    # headline path: the C++ host driver (TaskList/Driver/MHD tasks in csrc/host_driver.cpp) calling
    # the C ABI; NCCL communicator of its own for N > 1 (id broadcast through torch.distributed)
    from athenak_b200.sim import Simulation
    sim = Simulation(mesh, rank, ng=NG, recon="ppm4", rsolver="hlld", gamma=GAMMA, integrator="rk3",
                     cfl_number=0.3, tlim=1.0e9, fused=True)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(sim.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        sim.comm_init(bytes(idt.cpu().numpy().tobytes()))
    sim.set_state(hu0, hb1, hb2, hb3)
    sim.initialize()
    sstream = torch.cuda.ExternalStream(sim.status()["stream"], device=dev)

    def sync_all():
        sim.sync()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    K, W = args.steps, args.warmup
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # sampled from the warm-up on, so short timed regions still get samples under load
    sim.execute(W)
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = sim.status()["launches"]
    ev0.record(sstream)
    sim.execute(K)
    ev1.record(sstream)
    sync_all()
    launches = sim.status()["launches"] - l0
    ms = ev0.elapsed_time(ev1)
    halo_mode = ("same-rank gather kernel only" if world == 1 else
                 "peer-memory push kernel over NVLink (cudaIpc-mapped neighbour arrays, flag handshake); EMF correction packed straight into the neighbours' receive segments the same way; dt via ncclAllReduce(min)"
                 if sim.status()["peer_push"] else "grouped ncclSend/ncclRecv, one message per peer")
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    cells = float(n) ** 3 * world
    value = cells * K / (ms * 1e-3) / 1e6
    dt_now = sim.status()["dt"]

    # ---- end to end: host buffers in, host buffers out, every step --------------------------
    # H2D of the conserved state (u0 + 3 face fields, pinned) -> ghost exchange + c2p -> one full
    # step -> D2H of the updated conserved state, all through the C ABI (ak_sim_*).
    ho = [torch.empty_like(t).pin_memory() for t in (hu0, hb1, hb2, hb3)]
    ke2e = max(2, min(K, 5))
    h2d = sum(t.numel() * 8 for t in (hu0, hb1, hb2, hb3))
    d2h = sum(t.numel() * 8 for t in ho)

    def e2e_step():
        sim.set_state(hu0, hb1, hb2, hb3)
        sim.initialize()
        sim.execute(1)
        sim.get_state(ho)

    e2e_step()
    sync_all()
    e2e_mode = "serial: H2D -> step -> D2H, one engine instance"
    if world == 1:
        # Several engine instances (own stream + host thread each, ctypes releases the GIL): while one
        # instance computes step i, others upload step i+1's input and download step i-1's result,
        # so PCIe (both directions) and the SMs are busy at the same time.  Every step still does its own
        # H2D of the full conserved state from pinned memory, ghost exchange + c2p, one RK3 step, and a
        # D2H of the full result.
        NI = 3      # per instance the chain H2D -> step -> D2H is serial (same buffers); three chains in flight keep
                    # both PCIe directions and the SMs busy
        sims = [sim] + [Simulation(mesh, rank, ng=NG, recon="ppm4", rsolver="hlld", gamma=GAMMA, integrator="rk3",
                                   cfl_number=0.3, tlim=1.0e9, fused=True) for _ in range(NI - 1)]
        outs = [ho] + [[torch.empty_like(t).pin_memory() for t in ho] for _ in range(NI - 1)]
        per = max(2, min(K, 12) // NI)
        ke2e = per * NI

        def worker(s, out, nsteps):
            for _ in range(nsteps):
                s.set_state(hu0, hb1, hb2, hb3)
                s.initialize()
                s.execute(1)
                s.get_state(out)
            s.sync()

        for s_, o_ in zip(sims[1:], outs[1:]):
            worker(s_, o_, 1)
        sync_all()
        ths = [threading.Thread(target=worker, args=(s_, o_, per)) for s_, o_ in zip(sims, outs)]
        ev0.record(sstream)           # all streams idle here: the event time-stamps "now" on the device clock
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        ev1.record(sstream)           # every instance drained (worker ends with a stream sync): end of the last D2H
        sync_all()
        e2e_mode = f"{NI} engine instances over {NI} streams/host threads (H2D, step, D2H of consecutive steps overlap)"
        assert all(bool(torch.equal(x, y)) for o_ in outs[1:] for x, y in zip(ho, o_)), "the e2e instances disagree"
        del sims, outs
    else:
        ev0.record(sstream)
        for _ in range(ke2e):
            e2e_step()
        ev1.record(sstream)
        sync_all()
    ms_e2e = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = cells * ke2e / (ms_e2e * 1e-3) / 1e6
    del sim
    torch.cuda.empty_cache()

    # ---- per-kernel breakdown through the per-task C-ABI entry points -------------------------
    e = MHD(1, NG, (n, n, n), mb_size, nghbr, recon="ppm4", rsolver="hlld", gamma=GAMMA, my_rank=rank,
            gids_start=int(mesh.gids_start[rank]), device=dev)
    for dst, src in zip([e.u0] + e.b0, (hu0, hb1, hb2, hb3)):
        dst.copy_(src, non_blocking=True)
    drv = Driver("rk3", 0.3, tlim=1.0e9)
    tl = assemble_mhd_tasks(e, fused=True)
    drv.Initialize(e)
    drv.Execute(e, tl, ncycles=2)

    # ---- per-kernel timing (per-task ABI path, CUDA events on the launching stream) ----------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback (B200_PROFILING.md)")
    kern = {}
    if True:   # every rank runs the plan (the halo entries are collective); rank 0 reports
        nrep = 3
        class D:  # stage-2 coefficients
            gam0, gam1, beta, dt, nexp_stages = [0.25] * 3, [0.75] * 3, [0.25] * 3, dt_now, 3
        def timeit(fn):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fn(); torch.cuda.synchronize(dev)
            a.record()
            for _ in range(nrep):
                fn()
            b.record(); torch.cuda.synchronize(dev)
            return a.elapsed_time(b) / nrep
        ncell = float(n) ** 3
        # words per active cell that each task must move (DESIGN.md "kernels")
        plan = [("fluxes(x1+x2+x3)", lambda: e.Fluxes(D, 2), 3 * (8 + 7)),
                ("corner_emf", lambda: e.EField(D, 2), 6 + 3 + 6 + 3),
                ("rk_update", lambda: e.RKUpdate(D, 2), 15 + 5 + 5 + 5),
                ("ct", lambda: e.CT(D, 2), 3 + 3 + 3 + 3),
                ("cons2prim", lambda: e.ConToPrim(D, 2, count=False), 8 + 8),
                ("halo_u", lambda: e.SendRecvU(D, 2), 0), ("halo_b", lambda: e.SendRecvB(D, 2), 0)]
        for name, fn, words in plan:
            t_ms = timeit(fn)
            kern[name] = {"ms": t_ms, "algo_GBps": (words * 8 * ncell / (t_ms * 1e-3) / 1e9) if words else None}
    stage_ms = ms / K / 3.0
    ncell = float(n) ** 3
    path_achieved = ALGO_BYTES_PER_CELL_STAGE * ncell / (stage_ms * 1e-3) / 1e9   # per GPU
    dom = max(((k, v) for k, v in kern.items()), key=lambda kv: kv[1]["ms"])[0] if kern else None
    # dominant kernel = one flux sweep (x1 / x2 / x3 launches are the same work in three directions):
    # algorithmic bytes per launch = (8 words read + 7 written) per active cell
    flux_ms = kern["fluxes(x1+x2+x3)"]["ms"] / 3.0 if kern else float("nan")
    flux_bytes = 15 * 8 * ncell
    flux_achieved = flux_bytes / (flux_ms * 1e-3) / 1e9

    line = {
        "metric": METRIC, "value": value, "unit": "Mcell-updates/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"3D MHD blast {n}^3 cells per GPU, one MeshBlock per GPU, PPM4+HLLD+RK3, nghost=4, periodic, "
                               f"mesh {'x'.join(str(n * r) for r in nroot)}; step = one full RK3 time step",
                   "l2": "every field array is %.0f MB (> 126 MB L2); no explicit flush" % (8 * (n + 2 * NG) ** 3 / 1e6),
                   "fmad": False, "parity_build": True, "driver": "C++ host driver (ak_sim_*), fused task list",
                   "ghost_exchange": halo_mode},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "Mcell-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": ke2e, "ms_per_step": ms_e2e / ke2e, "mode": e2e_mode},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "flux sweep (flux_x1_kernel / flux_march_kernel<x2> / <x3>; 3 launches per stage, 63% of the stage)",
                     "achieved": flux_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": flux_achieved / hbm_peak,
                     "traffic": 2.13e9, "traffic_source": "ncu --set full, dram__bytes_read.sum+dram__bytes_write.sum per launch: 2.08/2.20/2.19 GB (profiles/r01_ncu_kernels.txt)",
                     "algorithmic_bytes_per_launch": flux_bytes, "avg_launch_ms": flux_ms, "peak_source": peak_src,
                     "note": "FP64-issue-bound, not HBM-bound: ~670 FP64 instructions per cell per sweep; ncu FP64-pipe utilisation 50-60%",
                     "path": {"what": "whole stage, 256 B algorithmic per active cell per stage (SURVEY.md 8d) / mean stage time",
                              "achieved": path_achieved, "frac": path_achieved / hbm_peak, "stage_ms": stage_ms},
                     "dominant_task": dom, "kernels": kern},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_cores()
        try:
            ncpu, rate = pick_reference_size(threads, budget_s=25.0, ncycles=3)
            r = run_reference_cpu(ncpu, 3, threads)
            v = ncpu ** 3 * len(r["per_cycle"][1:]) / sum(r["per_cycle"][1:]) / 1e6
            line["cpu_baseline"] = {"value": v, "unit": "Mcell-updates/s", "cores": threads, "kind": "reference",
                                    "sample": f"{ncpu}^3 single MeshBlock, 2 timed cycles after 1 warm-up, reference Kokkos::OpenMP build (oracle/_ref/athena_blast), all host threads"}
        except Exception as ex:  # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "Mcell-updates/s", "cores": threads, "kind": "reference",
                                    "sample": f"failed: {ex}"[:200]}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_native(args)


if __name__ == "__main__":
    main()
