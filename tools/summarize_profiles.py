#!/usr/bin/env python
"""Turn gpurun_out ncu artefacts into the tracked summaries under profiles/.

usage: python tools/summarize_profiles.py ROUND_TAG LAUNCHES.csv REPORT.ncu-rep
  LAUNCHES.csv : ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... <cmd>
  REPORT       : ncu --set full --clock-control none --import-source on -o ... <cmd>
writes profiles/<tag>_launches.csv (per-kernel durations and shares) and profiles/<tag>_kernels.csv
(per captured launch: duration, DRAM bytes, instruction counts, occupancy, stalls).
"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, report = sys.argv[1], sys.argv[2], sys.argv[3]
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)

rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
H, rows = rows[h], rows[h + 1:]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows:
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    agg.setdefault(r[ki].split("(")[0], []).append(v)
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(ROOT, "profiles", tag + "_launches.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "avg_us", "min_us", "max_us", "total_us", "share_pct"])
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        w.writerow([k, len(v), "%.2f" % (sum(v) / len(v)), "%.2f" % min(v), "%.2f" % max(v), "%.1f" % sum(v),
                    "%.1f" % (100 * sum(v) / tot)])

raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
H, U, data = rr[0], rr[1], rr[2:]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
idx = [(w, H.index(w)) for w in want if w in H]
with open(os.path.join(ROOT, "profiles", tag + "_kernels.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["%s [%s]" % (n, U[i]) if U[i] else n for n, i in idx])
    for r in data:
        w.writerow([r[i][:90] for _, i in idx])
# DRAM bytes per launch (read + write), averaged per kernel -> profiles/traffic.json (bench.py's roofline.traffic)
import json
import math
unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
ri, wi, ki = H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum"), H.index("Kernel Name")
gi, si, di = H.index("launch__grid_size"), H.index("launch__shared_mem_per_block_dynamic"), H.index("gpu__time_duration.sum")
# one kernel name can cover several launch shapes (proposal_kernel: top-k / NMS / fused, C2 or prefiltered):
# group by (grid, dynamic smem) and report, per name, the group with the longest launches -- for
# proposal_kernel that is the fused C2 launch bench.py times; every group is listed under "groups"
acc = collections.OrderedDict()
for r in data:
    name = r[ki].split("(")[0].split("<")[0].replace("void ", "").replace("tfrpn::", "").strip()
    us = float(r[di]) * {"us": 1.0, "ns": 1e-3, "ms": 1e3}.get(U[di], 1.0)
    # launches of one shape can still be different modes (top-k of 10 vs of 6000): half-octave duration buckets
    grp = "grid=%s smem=%s%s ~%dus" % (r[gi], r[si], U[si], round(2 ** (round(math.log2(max(us, 0.5)) * 2) / 2)))
    g = acc.setdefault(name, collections.OrderedDict()).setdefault(grp, {"bytes": [], "us": [], "inst": []})
    g["bytes"].append(float(r[ri]) * unit[U[ri]] + float(r[wi]) * unit[U[wi]])
    g["us"].append(us)
    if "smsp__inst_executed.sum" in H:
        g["inst"].append(float(r[H.index("smsp__inst_executed.sum")]))
mean = lambda v: sum(v) / len(v)
per_name, groups = collections.OrderedDict(), collections.OrderedDict()
for name, gs in acc.items():
    best = max(gs.items(), key=lambda kv: mean(kv[1]["us"]))
    per_name[name] = round(mean(best[1]["bytes"]))
    groups[name] = {k: dict({"dram_bytes": round(mean(v["bytes"])), "avg_us": round(mean(v["us"]), 2), "launches": len(v["us"])},
                            **({"warp_instr": round(mean(v["inst"]))} if v["inst"] else {}))
                    for k, v in gs.items()}
with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
    json.dump({"source": "profiles/%s_kernels.csv (ncu --set full --clock-control none, per launch)" % tag,
               "dram_bytes_per_launch": per_name, "groups": groups}, f, indent=1)
print("wrote profiles/%s_launches.csv, profiles/%s_kernels.csv and profiles/traffic.json" % (tag, tag))
