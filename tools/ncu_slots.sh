#!/bin/bash
# counters of the lattice kernel at FISS_GRID_SLOTS = 1, 2
mkdir -p gpurun_out
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__warp_issue_stalled_barrier_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct,smsp__warp_issue_stalled_not_selected_per_warp_active.pct,smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct,smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct,smsp__warps_active.avg.per_cycle_active,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__grid_size,launch__shared_mem_per_block_dynamic
for s in 1 2; do
  FISS_GRID_SLOTS=$s timeout 300 ncu --metrics $M --clock-control none -k regex:fiss_grid_kernel -s 4 -c 1 --csv \
     --log-file gpurun_out/ncus_$s.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>gpurun_out/ncus_$s.err
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncus_$s.csv")) if len(r)>10]
h=rows[0]
print("== slots=$s", rows[1][h.index("Kernel Name")][:40])
for r in rows[1:]:
    print("   %-75s %s %s" % (r[h.index("Metric Name")], r[h.index("Metric Value")], r[h.index("Metric Unit")]))
PY
done
FISS_GRID_SLOTS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fiss_grid_kernel -s 3 -c 1 -f -o gpurun_out/slots2_grid python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/slots2_ncu_full.log 2>&1
