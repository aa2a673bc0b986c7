#!/usr/bin/env python
"""Warp-stall sampling totals of one captured launch (ncu --page source), by stall reason.

    python profiles/stall_summary.py gpurun_out/prof.ncu-rep --launch 0
"""
import argparse
import csv
import io
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--launch", type=int, default=0)
args = ap.parse_args()
out = subprocess.run(["ncu", "-i", args.rep, "--page", "source", "--print-source", "sass", "--csv",
                      "--launch-skip", str(args.launch), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
data = []
for r in rows[2:]:
    if r == hdr:      # the page repeats the listing; keep the first copy
        break
    if len(r) == len(hdr):
        data.append(r)
print("#", rows[0][1])
tot = {}
for i, name in enumerate(hdr):
    if name.startswith("stall_") and "Not Issued" not in name:
        tot[name] = sum(int(r[i] or 0) for r in data)
s = sum(tot.values()) or 1
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print(f"{k:28s} {v:9d} {100 * v / s:5.1f}%")
