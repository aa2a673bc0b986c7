#!/bin/bash
# side workloads + the default bench (short); usage: gpurun -- 'bash tools/gpu_side.sh TAG'
TAG=${1:-side}
mkdir -p gpurun_out
for w in cfg4 cfg3 cfg5; do
  timeout 300 python bench.py --no-cpu-baseline --no-closed-loop --steps 100 --workload $w > gpurun_out/${TAG}_$w.json 2>gpurun_out/${TAG}_$w.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_$w.json")); r = d["roofline"]
    print("$w kernel_ms=%.4f (alone %.4f) frac=%.3f value=%.1fM (ms/step %.4f) winner_only=%.1fM e2e=%.1fM p50=%.4f" % (r["kernel_ms"], r["kernel_ms_alone"], r["frac"], d["value"]/1e6, d["ms_per_step"], d["value_winner_only"]/1e6, d["e2e"]["value"]/1e6, d["plan_cycle_p50_ms"]))
except Exception as e:
    print("$w FAILED", e)
PY
done
