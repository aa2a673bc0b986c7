#!/bin/bash
# Instruction / issue counters of the lattice kernel for every build/libfiss_*.so (A/B of kernel variants).
# usage: gpurun -- 'bash tools/ncu_counts.sh TAG'
TAG=${1:-cnt}
mkdir -p gpurun_out
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__warp_issue_stalled_barrier_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_not_selected_per_warp_active.pct,smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct,smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct,smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct,smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct,smsp__warp_issue_stalled_no_instruction_per_warp_active.pct,smsp__warp_issue_stalled_drain_per_warp_active.pct,smsp__warp_issue_stalled_membar_per_warp_active.pct,smsp__warps_active.avg.per_cycle_active
for lib in build/libfiss_*.so; do
  n=$(basename $lib .so)
  FISSGPU_LIB=$PWD/$lib timeout 300 ncu --metrics $M --clock-control none -k regex:fiss_grid_kernel -s 4 -c 1 --csv \
     --log-file gpurun_out/${TAG}_$n.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>gpurun_out/${TAG}_$n.err
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_$n.csv")) if len(r)>10]
h=rows[0]; 
print("== $n", rows[1][h.index("Kernel Name")][:40])
for r in rows[1:]:
    print("   %-75s %s %s" % (r[h.index("Metric Name")], r[h.index("Metric Value")], r[h.index("Metric Unit")]))
PY
done
