#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) into the text files committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof_X.ncu-rep profiles/r1_X

writes <out>_metrics.txt (per-kernel headline metrics) and <out>_hot_sass.txt (opcode histogram and
the hottest SASS blocks by executed-instruction count, from `ncu --page source`).
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(rep, out):
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    with open(out + "_metrics.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none, report {rep}\n")
        for r in rows:
            f.write(f"\n== {r[hdr.index('Kernel Name')]}\n")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"{m:70s} {r[i]:>16s} {units[i]}\n")
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    kern, cur, shdr = {}, None, None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = r[1]
            kern[cur] = []
        elif r and r[0] == "Address":
            shdr = r
        elif cur and len(r) > 5:
            kern[cur].append(r)
    with open(out + "_hot_sass.txt", "w") as f:
        for k, rs in kern.items():
            iex, isamp = shdr.index("Instructions Executed"), shdr.index("# Samples")
            tot = sum(int(r[iex]) for r in rs) or 1
            tots = sum(int(r[isamp]) for r in rs) or 1
            f.write(f"\n== {k}\nSASS lines {len(rs)}, warp instructions executed {tot}, stall samples {tots}\n")
            op, ops = collections.Counter(), collections.Counter()
            for r in rs:
                m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[1])
                o = m.group(2) if m else "?"
                op[o] += int(r[iex])
                ops[o] += int(r[isamp])
            f.write("opcode      executed      share   stall-sample share\n")
            for o, c in op.most_common(24):
                f.write(f"{o:10s} {c:12d} {100 * c / tot:6.1f}% {100 * ops[o] / tots:6.1f}%\n")
            f.write("hottest straight-line SASS blocks (first instruction shown)\n")
            blocks, start, prev = [], 0, None
            for i, r in enumerate(rs):
                c = int(r[iex])
                if prev is not None and abs(c - prev) > 0.02 * max(c, prev, 1):
                    blocks.append((start, i - 1, prev))
                    start = i
                prev = c
            blocks.append((start, len(rs) - 1, prev))
            for s, e, c in sorted(blocks, key=lambda b: -(b[1] - b[0] + 1) * b[2])[:12]:
                samp = sum(int(rs[i][isamp]) for i in range(s, e + 1))
                f.write(f"  lines {s:5d}-{e:5d} ({e - s + 1:4d} instr) x {c:10d} = {100 * (e - s + 1) * c / tot:5.1f}% of instr, "
                        f"{100 * samp / tots:5.1f}% of samples | {rs[s][1].strip()[:60]}\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
