"""Turns the round-2 ncu captures (gpurun_out/, made by profiles/r02_profile.sh) into the committed summaries:
  profiles/r02_launches_bench.txt   launch list of `python bench.py --steps 2 --warmup 1` (time, DRAM bytes, share of the listed time)
  profiles/r02_flux_fp64.json       FP64 instruction counts and DRAM traffic of the three sweeps (read by bench.py for roofline.traffic / fp64)
  profiles/r02_flux_counts.txt      the same for the cp.async and the TMA form side by side"""
import collections, csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")

def load(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]
    ik, iv, im, iid, ig = (h.index(x) for x in ("Kernel Name", "Metric Value", "Metric Name", "ID", "Grid Size"))
    d = collections.OrderedDict()
    for r in rows[1:]:
        e = d.setdefault(r[iid], {"kernel": r[ik], "grid": r[ig]})
        try:
            e[r[im]] = float(r[iv].replace(",", ""))
        except ValueError:      # "n/a": a metric this ncu build does not collect
            pass
    return list(d.values())

def short(k):
    k = k.replace("akb::", "").replace("void ", "")
    return k.split("(")[0]

# (1) launch list
L = load(os.path.join(G, "r02_launches_bench.csv"))
tot = sum(e["gpu__time_duration.sum"] for e in L)
agg = collections.OrderedDict()
for e in L:
    a = agg.setdefault(short(e["kernel"]), [0, 0.0, 0.0])
    a[0] += 1; a[1] += e["gpu__time_duration.sum"]; a[2] += e.get("dram__bytes_read.sum", 0) + e.get("dram__bytes_write.sum", 0)
with open(os.path.join(ROOT, "profiles", "r02_launches_bench.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 200 -c 80 "
            "python bench.py --steps 2 --warmup 1 --no-cpu-baseline   (B200, 256^3 blast; launches 200..279 of the process, cold-cache serialised times)\n\n")
    f.write(f"{'kernel':58s} {'launches':>8s} {'mean us':>9s} {'share %':>8s} {'DRAM GB/launch':>15s}\n")
    for k, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k[:58]:58s} {n:8d} {t / n / 1e3:9.1f} {100 * t / tot:8.1f} {b / n / 1e9:15.3f}\n")
    f.write("\nin launch order:\n")
    for e in L:
        f.write(f"{short(e['kernel'])[:58]:58s} {e['grid']:>16s} {e['gpu__time_duration.sum'] / 1e3:9.1f} us  "
                f"{(e.get('dram__bytes_read.sum', 0)) / 1e9:6.3f} + {(e.get('dram__bytes_write.sum', 0)) / 1e9:6.3f} GB\n")

# (2) flux counts
out = {}
txt = []
for tag, fn in (("cp.async", "r02_flux_counts.csv"), ("tma", "r02_flux_counts_tma.csv")):
    K = load(os.path.join(G, fn))
    txt.append(f"== {tag} form: ncu --metrics ... -k regex:flux_ -s 12 -c 3 python profiles/bench_variants.py --only headline --steps 2")
    for e in K:
        name = short(e["kernel"])
        rec = {"ms_ncu": e["gpu__time_duration.sum"] / 1e6, "warp_inst": e["smsp__inst_executed.sum"],
               "fp64_warp_inst": e["smsp__inst_executed_pipe_fp64.sum"],
               "fp64_pipe_pct_ncu": e["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"],
               "warps_active_pct": e["sm__warps_active.avg.pct_of_peak_sustained_active"], "registers": e["launch__registers_per_thread"],
               "dram_read_GB": e["dram__bytes_read.sum"] / 1e9, "dram_write_GB": e["dram__bytes_write.sum"] / 1e9}
        txt.append(f"{name:44s} {rec['ms_ncu']:.3f} ms  regs {rec['registers']:.0f}  warp-inst {rec['warp_inst']:.3e}  FP64 warp-inst {rec['fp64_warp_inst']:.3e} "
                   f"({100 * rec['fp64_warp_inst'] / rec['warp_inst']:.0f}%)  FP64 pipe {rec['fp64_pipe_pct_ncu']:.1f}%  warps active {rec['warps_active_pct']:.1f}%  "
                   f"DRAM {rec['dram_read_GB']:.3f} + {rec['dram_write_GB']:.3f} GB")
        if "sass__inst_executed_local_loads" in e:      # later captures: spill traffic and the two dominant stall reasons
            rec["local_loads"] = e["sass__inst_executed_local_loads"]; rec["local_stores"] = e["sass__inst_executed_local_stores"]
            rec["stall_long_scoreboard"] = e.get("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio")
            rec["stall_wait"] = e.get("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio")
            rec["issue_active_pct"] = e.get("sm__issue_active.avg.pct_of_peak_sustained_elapsed")
            txt.append(f"{'':44s} local loads {rec['local_loads']:.3e}  local stores {rec['local_stores']:.3e}  stalls per issue: long scoreboard "
                       f"{rec['stall_long_scoreboard']:.2f}, wait {rec['stall_wait']:.2f}  issue active {rec['issue_active_pct']:.1f}%")
        if tag == "cp.async":
            out[name] = rec
open(os.path.join(ROOT, "profiles", "r02_flux_counts.txt"), "w").write("\n".join(txt) + "\n")
n = len(out)
js = {"what": "FP64 pipe instruction counts (smsp__inst_executed_pipe_fp64.sum, warp level) and DRAM traffic of the three flux sweeps of ONE stage, "
              "3-D MHD blast 256^3 single MeshBlock, PPM4+HLLD, default build (range guard on), cycle 3",
      "command": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,... "
                 "--clock-control none -k regex:flux_ -s 12 -c 3 python profiles/bench_variants.py --only headline --steps 2  (profiles/r02_profile.sh)",
      "kernels": out,
      "per_launch_mean": {"fp64_warp_inst": sum(v["fp64_warp_inst"] for v in out.values()) / n,
                          "dram_bytes": sum(v["dram_read_GB"] + v["dram_write_GB"] for v in out.values()) / n * 1e9},
      "kernel_regex": "flux_(x1|march)_kernel<2, 2"}
json.dump(js, open(os.path.join(ROOT, "profiles", "r02_flux_fp64.json"), "w"), indent=1)
print(open(os.path.join(ROOT, "profiles", "r02_flux_counts.txt")).read())
