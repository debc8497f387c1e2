#!/usr/bin/env python
"""Top source lines by warp-stall samples from an ncu report (needs -lineinfo + --import-source on).

usage: python tools/ncu_source_summary.py REPORT.ncu-rep KERNEL_REGEX [launch_skip] [topN]
"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      "regex:" + kern, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
H = None
fname = ""
lines = []
for r in rows:
    if r and r[0] == "File Name":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No" and len(r) > 10:
        H = r
    elif H and len(r) == len(H) and r[0].isdigit():
        lines.append((fname, r))
if not H:
    sys.exit("no source table found")
ni = H.index("# Samples")
ii = H.index("Instructions Executed")
stall_cols = [(j, h) for j, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ni] or 0) for _, r in lines) or 1
tot_inst = sum(int(r[ii] or 0) for _, r in lines)
print("kernel %s: %d samples, %d warp instructions" % (kern, tot, tot_inst))
lines.sort(key=lambda fr: -int(fr[1][ni] or 0))
for f, r in lines[:top]:
    st = sorted(((int(r[j] or 0), h[6:]) for j, h in stall_cols), reverse=True)[:3]
    print("%5.1f%% inst=%8s %s:%-4s %-80s %s" % (100 * int(r[ni] or 0) / tot, r[ii], f[:12], r[0], r[1].strip()[:80],
                                               " ".join("%s:%d" % (h, v) for v, h in st if v)))
