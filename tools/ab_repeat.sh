#!/bin/bash
# repeatability of the chained step loop: the bench a few times (driver-style 20 steps and 100 steps), then the probe
mkdir -p gpurun_out
for i in 1 2 3; do
  for st in 20 100; do
    timeout 300 python bench.py --no-cpu-baseline --no-closed-loop --steps $st --warmup 5 > gpurun_out/rep_${i}_$st.json 2>gpurun_out/rep_${i}_$st.err
    python - <<PY
import json
d = json.load(open("gpurun_out/rep_${i}_$st.json")); r = d["roofline"]
print("run $i steps $st: kernel_ms=%.4f (alone %.4f) frac=%.3f value=%.1fM (ms/step %.4f, alone %.4f) winner_only=%.1fM e2e=%.1fM" % (r["kernel_ms"], r["kernel_ms_alone"], r["frac"], d["value"]/1e6, d["ms_per_step"], d["ms_per_step_median"], d["value_winner_only"]/1e6, d["e2e"]["value"]/1e6))
PY
  done
done
python tools/chain_probe.py
PROBE_STREAM=1 python tools/chain_probe.py
