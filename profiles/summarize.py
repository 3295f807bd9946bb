"""Turns the ncu captures brought back in gpurun_out/ into the text summaries committed under profiles/.
   python profiles/summarize.py launches <launches.csv> <out.txt> <title>
   python profiles/summarize.py kernels  <raw.csv from `ncu -i X.ncu-rep --page raw --csv`> <out.txt> <title>"""
import collections
import csv
import sys

mode, src, out, title = sys.argv[1:5]
rows = list(csv.reader(open(src)))
lines = ["# " + title]
if mode == "launches":
    hdr = next(r for r in rows if r and r[0] == "ID")
    data = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr)]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        k = r[kn].split("(")[0].replace("void ", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv]) / 1e6
    tot = sum(v[1] for v in agg.values())
    lines.append("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes")
    lines.append(f"{'kernel':62s} {'launches':>8s} {'total ms':>9s} {'avg ms':>8s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k:62s} {v[0]:8d} {v[1]:9.3f} {v[1]/v[0]:8.3f} {100*v[1]/tot:5.1f}%")
else:
    hdr, units = rows[0], rows[1]
    want = [("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy%"),
            ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
            ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe%"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active%"),
            ("l1tex__t_sector_hit_rate.pct", "l1_hit%"), ("lts__t_sector_hit_rate.pct", "l2_hit%"),
            ("smsp__inst_executed.sum", "warp_insts"),
            ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall_long_sb"),
            ("smsp__pcsamp_warps_issue_stalled_wait", "stall_wait"),
            ("smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "stall_math_throttle"),
            ("smsp__pcsamp_warps_issue_stalled_selected", "selected")]
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        grid = r[hdr.index("Grid Size")] if "Grid Size" in hdr else ""
        key = (name, grid)
        if key in seen:
            continue
        seen.add(key)
        lines.append(f"--- {name}  grid {grid}")
        for w, lab in want:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"    {lab:22s} {r[i][:16]:>16s} {units[i]}")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
