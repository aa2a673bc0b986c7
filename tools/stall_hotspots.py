#!/usr/bin/env python
"""Hottest SASS instructions of a captured launch by stall samples, with the stall reason and a few lines of context.

    python tools/stall_hotspots.py gpurun_out/x.ncu-rep [--top 25]
"""
import argparse, csv, io, subprocess
ap = argparse.ArgumentParser(); ap.add_argument("rep"); ap.add_argument("--top", type=int, default=25)
ap.add_argument("--reason", default=None, help="only this stall reason, e.g. stall_long_sb")
a = ap.parse_args()
out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
data = []
for r in rows[2:]:
    if r == hdr: break
    if len(r) == len(hdr): data.append(r)
si = hdr.index("# Samples"); src = hdr.index("Source")
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[si] or 0) for r in data)
key = (lambda r: int(r[hdr.index(a.reason)] or 0)) if a.reason else (lambda r: int(r[si] or 0))
order = sorted(range(len(data)), key=lambda i: -key(data[i]))[:a.top]
for i in sorted(order):
    r = data[i]
    reasons = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {100*int(r[si] or 0)/tot:5.2f}%  {r[src].strip():60.60s} " + " ".join(f"{n}:{v}" for v, n in reasons if v))
