#!/usr/bin/env python
"""Attribute the warp-level instructions an ncu capture counted to source lines of our kernels.

    python profiles/sass_by_line.py gpurun_out/prof.ncu-rep --launch 8 [--so fiss_plus_planner_b200/libfissgpu.so]
                                    [--top 40] [--ranges "P1:185-246,P2:248-279,P3:281-337"]

`ncu --page source` on the command line prints per-SASS-instruction counters but no source
correlation, so this joins them with `nvdisasm -g` (line info from -lineinfo) of the same kernel in
the in-tree .so by instruction order.  Library math (atan2, rsqrt, ...) has no line info of its own
and is attributed to the calling line.  Runs on the CPU box (no GPU needed).
"""
from __future__ import annotations

import argparse
import collections
import csv
import io
import os
import re
import subprocess
import tempfile


def ncu_sass(rep: str, launch: int):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv",
                          "--launch-skip", str(launch), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    kernel = rows[0][1]
    hdr = rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    data = []
    for r in rows[2:]:
        if r == hdr:  # the page repeats the listing; keep the first copy
            break
        if len(r) == len(hdr):
            data.append(r)
    return kernel, col, data


def disasm_lines(so: str, kernel_demangled: str):
    """[(sass text, file, line)] of the kernel whose demangled name matches."""
    tmp = tempfile.mkdtemp(prefix="sassline_")
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
    cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    text = ""
    for c in cubins:
        text += subprocess.run(["nvdisasm", "-g", "-c", c], capture_output=True, text=True).stdout
    # split into functions
    funcs = {}
    cur, cur_name = None, None
    file_, line_ = "?", 0
    for ln in text.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            cur_name = m.group(1)
            cur = funcs.setdefault(cur_name, [])
            file_, line_ = "?", 0
            continue
        if cur is None:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            file_, line_ = m.group(1), int(m.group(2))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            cur.append((int(m.group(1), 16), m.group(2).strip(), file_, line_))
    names = list(funcs)
    dem = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    norm = lambda s: re.sub(r"\(bool\)|fiss::|\s", "", s)
    want = norm(kernel_demangled)
    for n, d in zip(names, dem):
        if norm(d) == want:
            return funcs[n]
    raise SystemExit(f"kernel {kernel_demangled!r} not found in {so}; have {dem}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--launch", type=int, default=0, help="index of the captured launch (0-based)")
    ap.add_argument("--so", default="fiss_plus_planner_b200/libfissgpu.so")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--ranges", default="", help='named line ranges to total, e.g. "P1:185-246,P3:281-337"')
    ap.add_argument("--units", type=float, default=0.0, help="divide totals by this (e.g. candidates per launch)")
    args = ap.parse_args()

    kernel, col, data = ncu_sass(args.rep, args.launch)
    dis = disasm_lines(args.so, kernel)
    if len(dis) != len(data):
        print(f"# WARNING: {len(data)} SASS rows in the report vs {len(dis)} in {args.so} (rebuilt since the capture?)")
    n = min(len(dis), len(data))
    by_line = collections.Counter()
    by_line_stall = collections.Counter()
    by_op = collections.Counter()
    total = 0
    src_cache = {}
    for k in range(n):
        r = data[k]
        ex = int(r[col["Instructions Executed"]])
        smp = int(r[col["# Samples"]]) if "# Samples" in col else 0
        _, sass, f, ln = dis[k]
        key = (os.path.basename(f), ln)
        by_line[key] += ex
        by_line_stall[key] += smp
        by_op[sass.split()[0].split(".")[0] if not sass.startswith("@") else sass.split()[1].split(".")[0]] += ex
        total += ex
    div = args.units or 1.0
    unit = " /unit" if args.units else ""
    print(f"# {kernel}: {total} warp-instructions executed" + (f" = {total / div:.1f}{unit}" if args.units else ""))
    samples = sum(by_line_stall.values()) or 1

    def src(fname, ln):
        for root in ("fiss_plus_planner_b200/csrc", "include"):
            p = os.path.join(root, fname)
            if os.path.exists(p):
                if p not in src_cache:
                    src_cache[p] = open(p).read().splitlines()
                if 0 < ln <= len(src_cache[p]):
                    return src_cache[p][ln - 1].strip()[:90]
        return ""

    print(f"# top {args.top} source lines by instructions executed (share of instr | share of stall samples)")
    for (f, ln), ex in by_line.most_common(args.top):
        print(f"{ex / div:12.1f} {100 * ex / total:5.1f}% {100 * by_line_stall[(f, ln)] / samples:5.1f}%  {f}:{ln}  {src(f, ln)}")
    if args.ranges:
        print("# named ranges (name:[file:]lo-hi; default file fiss_kernels.cuh)")
        for item in args.ranges.split(","):
            parts = item.split(":")
            name, rng = parts[0], parts[-1]
            fname = parts[1] if len(parts) == 3 else "fiss_kernels.cuh"
            lo, hi = [int(v) for v in rng.split("-")]
            ex = sum(v for (f, ln), v in by_line.items() if f == fname and lo <= ln <= hi)
            st = sum(v for (f, ln), v in by_line_stall.items() if f == fname and lo <= ln <= hi)
            print(f"{name:>12}: {ex / div:12.1f}{unit} {100 * ex / total:5.1f}% of instr, {100 * st / samples:5.1f}% of samples")
    print("# by opcode")
    for op, ex in by_op.most_common(25):
        print(f"{op:>10} {ex / div:12.1f} {100 * ex / total:5.1f}%")


if __name__ == "__main__":
    main()
