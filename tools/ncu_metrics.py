#!/usr/bin/env python
"""Print the headline metrics of every kernel launch in an ncu report.
usage: python tools/ncu_metrics.py REPORT.ncu-rep [kernel-regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
rx = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
H = r[0]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__waves_per_multiprocessor", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for row in r[2:]:
    name = row[H.index("Kernel Name")]
    if rx and not rx.search(name):
        continue
    print(name[:80])
    for w in want:
        if w in H:
            print("   %-62s %s" % (w, row[H.index(w)]))
