#!/usr/bin/env python
"""Compact per-kernel summary of an .ncu-rep (run where ncu is installed): python tools/ncu_summary.py rep [out.csv]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio" ]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
out = []
for r in data:
    d = {"kernel": r[idx["Kernel Name"]][:70]}
    for c in cols:
        if c in idx:
            d[c] = r[idx[c]] + " " + units[idx[c]]
    st = sorted(((float(r[idx[h]] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stall), reverse=True)[:5]
    d["top_stalls(warps per issue)"] = "; ".join(f"{n}={v:.2f}" for v, n in st)
    out.append(d)
w = csv.writer(open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout)
keys = ["kernel"] + [c for c in cols if c in idx] + ["top_stalls(warps per issue)"]
w.writerow(keys)
for d in out:
    w.writerow([d.get(k, "") for k in keys])
