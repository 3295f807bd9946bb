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
# The pipelined end-to-end leg drives several engine instances per GPU, each with its own compute, auxiliary and sweep streams
# and with spinning flag-wait kernels on some of them: with the default 8 hardware work queues two unrelated streams can share a
# queue and a spinning kernel then blocks the very signal it waits for.  One queue per stream (set before CUDA initialises).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

GAMMA = 1.6666667
NG = 4
DECK = os.path.join(ROOT, "tests", "decks", "blast3d.athinput")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "athena_blast")
REF_BIN_CUDA = os.path.join(ROOT, "oracle", "_ref", "athena_blast_cuda")     # oracle/build_ref.sh cuda (the reference's own Kokkos-CUDA build)
ALGO_BYTES_PER_CELL_STAGE = 256.0      # 32 FP64 words per active cell per stage
METRIC = "Mcell-updates/s on 3D MHD blast 256^3/GPU"
NCU_FLUX = os.path.join(ROOT, "profiles", "r02_flux_fp64.json")     # committed ncu summary the roofline's traffic / FP64 counts come from


def workload(n):
    """One string for both arms (the driver compares the two config.workload values)."""
    return (f"3D MHD blast {n}^3 cells per GPU, one MeshBlock per GPU, PPM4+HLLD+RK3, nghost=4, periodic; "
            "step = one full RK3 time step")


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------------------
# reference CPU arm (unmodified reference binary built by oracle/build_ref.sh)

def run_reference_cpu(n, ncycles, threads, timeout=1500, binary=None):
    """Runs oracle/_ref/athena_blast at n^3 for ncycles; returns per-cycle wall seconds (from the
    binary's own 'elapsed=' diagnostics) and its zone-cycles/cpu_second line.  binary: another build of the same
    unmodified reference (the Kokkos-CUDA one for the reference_gpu leg)."""
    REF_BIN = binary or globals()["REF_BIN"]
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
        "config": {"workload": workload(args.nx), "implementation": "reference Kokkos::OpenMP CPU build (oracle/_ref/athena_blast), unmodified",
                   "sample": sample},
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
        # NVML in-process (what nvidia-smi itself reads), one light query per 50 ms: a looping nvidia-smi process re-queries the
        # whole device record every period and measurably slows kernel launches of a launch-heavy step (seen: -10% on `value`).
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nvml, self.stop_flag, self.idle = pynvml, False, 0

            def poll():
                bits = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))
                while not self.stop_flag:
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        try:
                            r = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                        except Exception:
                            r = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        f = ["", str(sm), str(mx), "", ""] + ["Active" if r & b else "Not Active" for b, _ in bits]
                        self.lines.append(",".join(f))
                    except Exception:
                        pass
                    time.sleep(0.05)

            threading.Thread(target=poll, daemon=True).start()
            return
        except Exception:
            self.nvml = None
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
        if getattr(self, "nvml", None):
            self.stop_flag = True
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
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if getattr(self, "nvml", None) else "nvidia-smi"}



def measure_fp64_peak(L):
    """FP64 issue peak of this device (csrc/peaks.cu): best lane-instruction rate over the DADD/DMUL mix an --fmad=false
    build issues, ILP 4 and 8, 8 to 16 resident warps per SM sub-partition."""
    import ctypes as C
    best, how = 0.0, None
    for op, name in ((3, "DADD/DMUL"), (0, "DFMA")):
        for ilp in (4, 8):
            for w in (8, 12, 16):
                r, ms = C.c_double(), C.c_double()
                if L.ak_peak_fp64(op, ilp, w, 2048 // ilp, C.byref(r), C.byref(ms)) == 0 and r.value > best:
                    best, how = r.value, f"{name} chains, ilp {ilp}, {w} warps/SMSP"
    return best, how


def make_sim(mesh, rank, world, dev, dist, **kw):
    """Simulation on the C++ host driver with its own NCCL communicator for N > 1 (id broadcast through torch.distributed)."""
    import torch
    from athenak_b200.sim import Simulation
    sim = Simulation(mesh, rank, ng=NG, gamma=GAMMA, cfl_number=0.3, tlim=1.0e9, fused=True, **kw)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(sim.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        sim.comm_init(bytes(idt.cpu().numpy().tobytes()))
    return sim


def parity_leg(rank, world, dev, dist):
    """N > 1 only, before the timed region: two small runs through the SAME driver / exchange path the bench times.
    (1) 16^3 blast, 8 MeshBlocks of 8^3 over the ranks, 4 cycles, against the reference's own restart dump
        (tests/golden/blast16_ppm4_hlld_rk3.npz, made by tests/gen_golden.py from oracle/_ref).
    (2) Orszag-Tang 16^3 on the same 8 blocks, 3 cycles -- there the SendE/RecvE correction changes bits, so a lost or
        misplaced EMF message shows -- against the same run on ONE rank (every rank repeats it locally), ghost faces included.
    Returns the max |difference| over all ranks (must be exactly 0) and whether the peer-memory push carried the exchange."""
    import numpy as np
    import torch
    from athenak_b200.mesh import Mesh
    d = np.load(os.path.join(ROOT, "tests", "golden", "blast16_ppm4_hlld_rk3.npz"))
    worst, push = 0.0, False
    mesh = Mesh((16, 16, 16), (8, 8, 8), nranks=world)
    g0, n = int(mesh.gids_start[rank]), int(mesh.nmb_rank[rank])
    sim = make_sim(mesh, rank, world, dev, dist, recon="ppm4", rsolver="hlld", integrator="rk3")
    sim.set_state(*mesh.pgen_blast(rank, NG, GAMMA))
    sim.initialize()
    worst = max(worst, abs(sim.status()["dt"] - float(d["dt_c0"])))
    sim.execute(4)
    u, x1, x2, x3 = sim.get_state()
    for m in range(n):
        lx, ly, lz = (int(v) * 8 for v in mesh.lloc[g0 + m, :3])
        for got, ref in ((u[m, :, 4:-4, 4:-4, 4:-4], d["u0_c4"][0][:, lz:lz + 8, ly:ly + 8, lx:lx + 8]),
                         (x1[m, 4:-4, 4:-4, 4:-4], d["b1f_c4"][0][lz:lz + 8, ly:ly + 8, lx:lx + 9]),
                         (x2[m, 4:-4, 4:-4, 4:-4], d["b2f_c4"][0][lz:lz + 8, ly:ly + 9, lx:lx + 8]),
                         (x3[m, 4:-4, 4:-4, 4:-4], d["b3f_c4"][0][lz:lz + 9, ly:ly + 8, lx:lx + 8])):
            worst = max(worst, float(np.abs(got - ref).max()))
    worst = max(worst, abs(sim.status()["time"] - float(d["time_c4"])))
    push = sim.status()["peer_push"]
    del sim
    # Orszag-Tang: N ranks against one rank
    from athenak_b200.sim import Simulation
    one = Mesh((16, 16, 16), (8, 8, 8), nranks=1)
    ref = Simulation(one, 0, ng=NG, gamma=GAMMA, cfl_number=0.3, tlim=1.0e9, fused=True, recon="ppm4", rsolver="hlld", integrator="rk3")
    ref.set_state(*one.pgen_orszag_tang(0, NG, GAMMA))
    ref.initialize()
    ref.execute(3)
    r = ref.get_state()
    rt = ref.status()
    del ref
    sim = make_sim(mesh, rank, world, dev, dist, recon="ppm4", rsolver="hlld", integrator="rk3")
    sim.set_state(*mesh.pgen_orszag_tang(rank, NG, GAMMA))
    sim.initialize()
    sim.execute(3)
    got = sim.get_state()
    s = slice(NG, -NG)
    worst = max(worst, float(np.abs(got[0][:, :, s, s, s] - r[0][g0:g0 + n, :, s, s, s]).max()))
    for x, y in zip(got[1:], r[1:]):
        worst = max(worst, float(np.abs(x - y[g0:g0 + n]).max()))       # face fields: ghost zones included
    worst = max(worst, abs(sim.status()["time"] - rt["time"]), abs(sim.status()["dt"] - rt["dt"]))
    del sim
    t = torch.tensor([worst], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), bool(push)


def bench_native(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from athenak_b200 import lib as _lib
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

    def sync_all(sim=None):
        if sim is not None:
            sim.sync()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def maxrank(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- N > 1: bitwise parity of the multi-rank path, before anything is timed -----------------
    parity = None
    if world > 1:
        diff, push = parity_leg(rank, world, dev, dist)
        parity = {"checked": True, "max_abs_diff": diff, "peer_push": push,
                  "what": "16^3 blast on 8 MeshBlocks over the ranks vs the reference's restart dump (tests/golden), "
                          "Orszag-Tang 16^3 (EMF correction changes bits) vs the same run on one rank; C++ driver, same exchange path as the timed run"}
        if diff != 0.0 or os.environ.get("AKB_BENCH_PARITY_ONLY") == "1":
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "multi-rank parity failed" if diff != 0.0 else None, "parity": parity}))
            dist.destroy_process_group()
            raise SystemExit(2 if diff != 0.0 else 0)

    mesh = Mesh(tuple(n * r for r in nroot), (n, n, n), xmin=tuple(-0.5 * r for r in nroot),
                xmax=tuple(0.5 * r for r in nroot), nranks=world)
    assert int(mesh.nmb_rank[rank]) == 1
    mb_size, _, nghbr = mesh.rank_tables(rank)
    hu0, hb1, hb2, hb3 = mesh.pgen_blast(rank, NG, GAMMA, pinned=True)
    # headline path: the C++ host driver (TaskList/Driver/MHD tasks in csrc/host_driver.cpp) calling the C ABI
    HEAD = dict(recon="ppm4", rsolver="hlld", integrator="rk3")
    sim = make_sim(mesh, rank, world, dev, dist, **HEAD)
    sim.set_state(hu0, hb1, hb2, hb3)
    sim.initialize()
    sstream = torch.cuda.ExternalStream(sim.status()["stream"], device=dev)

    K, W = args.steps, args.warmup
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # sampled from the warm-up on, so short timed regions still get samples under load
    sim.execute(W)
    sync_all(sim)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = sim.status()["launches"]
    ev0.record(sstream)
    sim.execute(K)
    ev1.record(sstream)
    sync_all(sim)
    launches = sim.status()["launches"] - l0
    ms = ev0.elapsed_time(ev1)
    halo_mode = ("same-rank gather kernel only" if world == 1 else
                 "peer-memory push kernel over NVLink (cudaIpc-mapped neighbour arrays, flag handshake); EMF correction packed straight into the neighbours' receive segments the same way; dt via ncclAllReduce(min)"
                 if sim.status()["peer_push"] else "grouped ncclSend/ncclRecv, one message per peer")
    clocks = sampler.stop() if rank == 0 else None
    ms = maxrank(ms)
    cells = float(n) ** 3 * world
    value = cells * K / (ms * 1e-3) / 1e6
    dt_now = sim.status()["dt"]

    # ---- end to end: host buffers in, host buffers out, every step --------------------------
    # H2D of the conserved state (u0 + 3 face fields, pinned) -> ghost exchange + c2p -> one full
    # step -> D2H of the updated conserved state, all through the C ABI (ak_sim_*).
    ho = [torch.empty_like(t).pin_memory() for t in (hu0, hb1, hb2, hb3)]
    h2d = sum(t.numel() * 8 for t in (hu0, hb1, hb2, hb3))
    d2h = sum(t.numel() * 8 for t in ho)

    def e2e_step(s_, out):
        s_.set_state(hu0, hb1, hb2, hb3)
        s_.initialize()
        s_.execute(1)
        s_.get_state(out)

    e2e_step(sim, ho)
    sync_all(sim)
    # (a) serial reference point: one engine instance, H2D -> step -> D2H
    ks = 2
    ev0.record(sstream)
    for _ in range(ks):
        e2e_step(sim, ho)
    ev1.record(sstream)
    sync_all(sim)
    ms_serial = maxrank(ev0.elapsed_time(ev1)) / ks
    # (b) pipelined: ak_pipe_* (csrc/host_driver.cpp) over several engine instances -- one upload and one download stream
    # shared by all of them, one worker thread per instance (at N > 1 each instance has its own NCCL communicator and
    # peer-memory mapping): while one instance computes step i, the upload of step i+1 and the download of step i-1 run, so
    # both PCIe directions and the SMs are busy at the same time.  Every step still does its own H2D of the full conserved
    # state from pinned memory, ghost exchange + c2p, one RK3 step, and a D2H of the full result, and the outputs are checked
    # against the serial leg's bit for bit.  Guarded by a watchdog: if the jobs do not finish, the serial figure is reported.
    NI = int(os.environ.get("AKB_E2E_SLOTS", "5"))
    e2e_mode = "serial: H2D -> step -> D2H, one engine instance"
    ms_e2e, ke2e = ms_serial * ks, ks
    hung = False
    try:
        from athenak_b200.sim import HostPipeline
        sims = [sim] + [make_sim(mesh, rank, world, dev, dist, **HEAD) for _ in range(NI - 1)]
        outs = [ho] + [[torch.empty_like(t).pin_memory() for t in ho] for _ in range(NI - 1)]
        pipe = HostPipeline(sims)
        up_s, down_s = (torch.cuda.ExternalStream(h, device=dev) for h in pipe.streams())
        hin = (hu0, hb1, hb2, hb3)
        njobs = max(2 * NI, min(K, 30))

        def run_jobs(nj, result):
            try:
                for j in range(nj):
                    pipe.submit(hin, outs[j % NI], 1)
                    if j >= NI - 1:              # a slot's output buffer is reused only after its previous job is done
                        pipe.wait(j - (NI - 1))
                pipe.wait(-1)
                result.append(True)
            except Exception as ex:      # noqa: BLE001
                result.append(ex)

        def guarded(nj):
            res = []
            th = threading.Thread(target=run_jobs, args=(nj, res), daemon=True)
            th.start()
            th.join(240.0)
            if th.is_alive():
                raise TimeoutError("pipelined e2e jobs did not finish in 240 s")
            if res[0] is not True:
                raise res[0]

        guarded(NI)                   # warm-up: every slot once
        sync_all(sim)
        ev0.record(up_s)              # all streams idle here: the event time-stamps "now" on the device clock
        guarded(njobs)
        ev1.record(down_s)            # behind the last download
        sync_all(sim)
        assert all(bool(torch.equal(x, y)) for o_ in outs[1:] for x, y in zip(ho, o_)), "the e2e instances disagree"
        ms_e2e, ke2e = maxrank(ev0.elapsed_time(ev1)), njobs
        e2e_mode = (f"ak_pipe_*: {NI} engine instances per rank, one shared upload stream + one shared download stream "
                    "(H2D of step i+1, RK3 step i, D2H of step i-1 overlap; each step = full H2D -> ghost exchange + c2p -> step -> full D2H)")
        pipe.close()
        del pipe, sims, outs
    except TimeoutError as ex:
        e2e_mode += f" (pipelined mode abandoned: {ex})"
        hung = True
    e2e_value = cells * ke2e / (ms_e2e * 1e-3) / 1e6
    if not hung:
        del sim
        torch.cuda.empty_cache()

    line = {
        "metric": METRIC, "value": value, "unit": "Mcell-updates/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload(n), "mesh": "x".join(str(n * r) for r in nroot),
                   "l2": "every field array is %.0f MB (> 126 MB L2); no explicit flush" % (8 * (n + 2 * NG) ** 3 / 1e6),
                   "fmad": False, "parity_build": True, "division_guard": "deferred range guard on (IEEE result for every operand)",
                   "driver": "C++ host driver (ak_sim_*), fused task list", "ghost_exchange": halo_mode},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "Mcell-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": ke2e, "ms_per_step": ms_e2e / ke2e, "mode": e2e_mode,
                "serial_ms_per_step": ms_serial, "serial_value": cells / (ms_serial * 1e-3) / 1e6},
        "gpu_launches": launches,
    }
    if parity is not None:
        line["parity"] = parity
    if hung:      # threads stuck inside the library cannot be joined: report what was measured and leave
        if rank == 0:
            print(json.dumps(line), flush=True)
        os._exit(0)

    # ---- BASELINE config 3's per-GPU layout as a secondary block: Orszag-Tang, 64 MeshBlocks of (n/4)^3 per GPU ---------
    try:
        q = n // 4
        m3 = Mesh(tuple(n * r for r in nroot), (q, q, q), xmin=tuple(-0.5 * r for r in nroot), xmax=tuple(0.5 * r for r in nroot),
                  nranks=world)
        s3 = make_sim(m3, rank, world, dev, dist, **HEAD)
        s3.set_state(*m3.pgen_orszag_tang(rank, NG, GAMMA))
        s3.initialize()
        st3 = torch.cuda.ExternalStream(s3.status()["stream"], device=dev)
        k3 = max(3, min(K, 10))
        s3.execute(3)
        sync_all(s3)
        l3 = s3.status()["launches"]
        ev0.record(st3)
        s3.execute(k3)
        ev1.record(st3)
        sync_all(s3)
        ms3 = maxrank(ev0.elapsed_time(ev1))
        line["config3_ot"] = {"workload": f"Orszag-Tang 3D MHD, {'x'.join(str(n * r) for r in nroot)} cells, {int(m3.nmb_rank[rank])} MeshBlocks of {q}^3 per GPU, "
                                          "PPM4+HLLD+RK3, nghost=4, uniform mesh (BASELINE config 3 layout)",
                              "value": cells * k3 / (ms3 * 1e-3) / 1e6, "unit": "Mcell-updates/s", "steps": k3, "ms_per_step": ms3 / k3,
                              "gpu_launches": s3.status()["launches"] - l3, "peer_push": s3.status()["peer_push"]}
        del s3
        torch.cuda.empty_cache()
    except Exception as ex:      # secondary block: never takes the headline down
        line["config3_ot"] = {"error": str(ex)[:200]}

    # ---- the same headline run with the x2/x3 sweeps through the TMA bulk-copy kernel (flux_march_tma_kernel) --------------
    try:
        E = _lib.load()
        prev = E.ak_mhd_set_march_tma(1)
        st_ = make_sim(mesh, rank, world, dev, dist, **HEAD)
        st_.set_state(hu0, hb1, hb2, hb3)
        st_.initialize()
        stt = torch.cuda.ExternalStream(st_.status()["stream"], device=dev)
        kt = max(3, min(K, 10))
        st_.execute(W)
        sync_all(st_)
        ev0.record(stt)
        st_.execute(kt)
        ev1.record(stt)
        sync_all(st_)
        mst = maxrank(ev0.elapsed_time(ev1))
        E.ak_mhd_set_march_tma(prev)
        # the default path over the same cycles, for a like-for-like ratio (the step gets slower as the blast's far-field shell grows)
        sd_ = make_sim(mesh, rank, world, dev, dist, **HEAD)
        sd_.set_state(hu0, hb1, hb2, hb3)
        sd_.initialize()
        sdd = torch.cuda.ExternalStream(sd_.status()["stream"], device=dev)
        sd_.execute(W)
        sync_all(sd_)
        ev0.record(sdd)
        sd_.execute(kt)
        ev1.record(sdd)
        sync_all(sd_)
        msd = maxrank(ev0.elapsed_time(ev1))
        line["tma_march"] = {"what": "headline workload with the x2/x3 flux sweeps fetching their ring rows by cp.async.bulk (TMA, SASS UBLKCP; "
                                     "ak_mhd_set_march_tma) instead of per-thread cp.async; bit-identical results; same cycles for both figures",
                             "value": cells * kt / (mst * 1e-3) / 1e6, "cp_async_value": cells * kt / (msd * 1e-3) / 1e6,
                             "unit": "Mcell-updates/s", "steps": kt, "default": False}
        del st_, sd_
        torch.cuda.empty_cache()
    except Exception as ex:      # secondary block
        line["tma_march"] = {"error": str(ex)[:200]}

    # ---- the reference's own GPU path on this device: its Kokkos-CUDA build of the same deck (oracle/build_ref.sh cuda) ----
    if rank == 0 and world == 1 and os.path.exists(REF_BIN_CUDA) and not args.no_cpu_baseline:
        try:
            r = run_reference_cpu(n, 13, 1, timeout=600, binary=REF_BIN_CUDA)
            pc = r["per_cycle"][3:-1] or r["per_cycle"]          # first cycles warm up, the last one includes the final output
            line["reference_gpu"] = {"value": n ** 3 * len(pc) / sum(pc) / 1e6, "unit": "Mcell-updates/s", "kind": "reference",
                                     "what": f"unmodified reference, Kokkos-CUDA build for sm_100 (oracle/_ref/athena_blast_cuda), {n}^3 single MeshBlock, "
                                             f"{len(pc)} timed cycles after 3 warm-up, same device, per-cycle times from its own elapsed= diagnostics"}
        except Exception as ex:
            line["reference_gpu"] = {"value": None, "what": f"failed: {ex}"[:200]}

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
    flux_share = 3.0 * flux_ms / sum(v["ms"] for v in kern.values()) if kern else float("nan")
    # measured DRAM traffic and FP64 instruction counts of the sweeps: committed ncu summary (256^3 launches only)
    ncu = {}
    try:
        ncu = json.load(open(NCU_FLUX))
    except Exception:
        pass
    traffic = ncu.get("per_launch_mean", {}).get("dram_bytes") if n == 256 else None
    fp64_inst = ncu.get("per_launch_mean", {}).get("fp64_warp_inst") if n == 256 else None
    fp64_peak, fp64_how = measure_fp64_peak(_lib.load())
    fp64 = None
    if fp64_inst and fp64_peak:
        ach = 32.0 * fp64_inst / (flux_ms * 1e-3)
        fp64 = {"what": "FP64 issue roofline of the same kernel: warp-level FP64 instructions (DADD/DMUL/DFMA/DSETP) per launch x 32 lanes / launch time, "
                        "against the lane-instruction rate of a register-only FP64 chain kernel measured in this run (ak_peak_fp64)",
                "achieved": ach, "peak": fp64_peak, "unit": "FP64 lane-instructions/s", "frac": ach / fp64_peak,
                "peak_how": fp64_how, "fp64_warp_inst_per_launch": fp64_inst, "counts_source": os.path.relpath(NCU_FLUX, ROOT)}
    line["roofline"] = {
        "bound": "hbm", "kernel": "flux sweep (flux_x1_kernel / flux_march_kernel<x2> / <x3>; 3 launches per stage, "
                                  f"{100.0 * flux_share:.0f}% of the per-task stage time measured below)",
        "achieved": flux_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": flux_achieved / hbm_peak,
        "traffic": traffic, "traffic_source": f"{os.path.relpath(NCU_FLUX, ROOT)}: ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean of the "
                                              f"three sweep launches, kernel regex '{ncu.get('kernel_regex')}'" if traffic else None,
        "algorithmic_bytes_per_launch": flux_bytes, "avg_launch_ms": flux_ms, "peak_source": peak_src,
        "note": "the sweeps are bound by FP64 issue, not by HBM: see fp64 (measured issue peak) next to the HBM figure",
        "fp64": fp64,
        "path": {"what": "whole stage, 256 B algorithmic per active cell per stage (SURVEY.md 8d) / mean stage time",
                 "achieved": path_achieved, "frac": path_achieved / hbm_peak, "stage_ms": stage_ms},
        "dominant_task": dom, "kernels": kern}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_cores()
        try:
            ncpu, rate = pick_reference_size(threads, budget_s=45.0, ncycles=6)
            r = run_reference_cpu(ncpu, 6, threads)
            v = ncpu ** 3 * len(r["per_cycle"][1:]) / sum(r["per_cycle"][1:]) / 1e6
            line["cpu_baseline"] = {"value": v, "unit": "Mcell-updates/s", "cores": threads, "kind": "reference",
                                    "sample": f"{ncpu}^3 single MeshBlock, {len(r['per_cycle'][1:])} timed cycles after 1 warm-up, reference Kokkos::OpenMP build (oracle/_ref/athena_blast), all host threads"}
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
