#!/usr/bin/env python
"""Per-kernel sums of ONE bench step from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: launch_list_summary.py launches.csv out.csv   (takes the LAST complete step: from the last init_build_kernel on)"""
import csv, sys, re
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = []
for r in rows:
    if r is hdr or len(r) != len(hdr) or r[ik] == "Kernel Name":
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    unit = r[iu]
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}.get(unit, 1e-6)
    name = re.sub(r"<.*", "", r[ik].replace("void ", "").replace("ibvh::", "").replace("(anonymous namespace)::", ""))
    name = name.split("(")[0].strip()
    launches.append((name, ms))
starts = [i for i, (n, _) in enumerate(launches) if n.startswith("init_build_kernel")]
step = launches[starts[-2]:starts[-1]] if len(starts) >= 2 else launches[starts[-1]:]
agg = {}
for n, ms in step:
    agg[n] = agg.get(n, 0.0) + ms
tot = sum(agg.values())
w = csv.writer(open(sys.argv[2], "w"))
w.writerow(["kernel", "launches_in_step", "launch_ms_sum_in_step", "share_of_step_pct"])
for n in agg:
    w.writerow([n, sum(1 for m, _ in step if m == n), round(agg[n], 4), round(100 * agg[n] / tot, 1)])
w.writerow(["TOTAL (our kernels)", len(step), round(tot, 4), 100.0])
print(open(sys.argv[2]).read())
