#!/bin/bash
# A/B of an environment switch of the library on the bench workloads; usage: gpurun -- 'bash tools/ab_env.sh VAR "cfg4 cfg5"'
VAR=$1; WL=${2:-"cfg4"}
mkdir -p gpurun_out
for w in $WL; do
  for on in 0 1; do
    if [ $on == 1 ]; then export $VAR=1; else unset $VAR; fi
    timeout 300 python bench.py --no-cpu-baseline --steps 100 --workload $w > gpurun_out/abenv_${w}_$on.json 2>gpurun_out/abenv_${w}_$on.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/abenv_${w}_$on.json"))
    print("$w $VAR=$on kernel_ms=%.4f frac=%.3f value=%.1fM winner_only=%.1fM e2e=%.1fM" % (d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["value"]/1e6, d["value_winner_only"]/1e6, d["e2e"]["value"]/1e6))
except Exception as e:
    print("$w $VAR=$on FAILED", e)
PY
  done
done
