#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: count, mean us, share.

    python profiles/launch_summary.py profiles/r1c_launches.csv
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[h + 1:]:
    if len(r) == len(hdr):
        agg[r[k]].append(float(r[v].replace(",", "")))
tot = sum(sum(x) for x in agg.values())
print("# %s: %d launches, %.1f us total (cold-cache, serialised: shares matter, not absolutes)" % (sys.argv[1], sum(map(len, agg.values())), tot / 1e3))
for n, x in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{n[:72]:72s} n={len(x):4d} mean_us={sum(x) / len(x) / 1e3:9.2f} share={sum(x) / tot:.3f}")
