#!/bin/bash
# A/B of an environment variable of the library over several values; usage: gpurun -- 'bash tools/ab_vals.sh VAR "v1 v2 ..." [workloads]'
VAR=$1; VALS=$2; WL=${3:-"cfg4"}
mkdir -p gpurun_out
for w in $WL; do
  for v in $VALS; do
    export $VAR=$v
    timeout 300 python bench.py --no-cpu-baseline --no-closed-loop --steps 100 --workload $w > gpurun_out/abv_${w}_$v.json 2>gpurun_out/abv_${w}_$v.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/abv_${w}_$v.json"))
    print("$w $VAR=$v kernel_ms=%.4f frac=%.3f value=%.1fM winner_only=%.1fM e2e=%.1fM p50=%.4f" % (d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["value"]/1e6, d["value_winner_only"]/1e6, d["e2e"]["value"]/1e6, d["plan_cycle_p50_ms"]))
except Exception as e:
    print("$w $VAR=$v FAILED", e)
PY
  done
done
