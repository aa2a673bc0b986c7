#!/bin/bash
# parity tests, then the latency numbers of the bench (plan cycle p50 of config 2, closed loop of config 1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --steps 100 > gpurun_out/p50.json 2>gpurun_out/p50.err
python - <<PY
import json
d = json.load(open("gpurun_out/p50.json"))
cl = d["closed_loop"]
print("p50=%.4f closed loop p50: FOP %.4f FOP+ %.4f FISS %.4f FISS+ %.4f  value=%.1fM winner_only=%.1fM e2e=%.1fM" % (d["plan_cycle_p50_ms"], cl["FOP"]["p50_ms"], cl["FOP+"]["p50_ms"], cl["FISS"]["p50_ms"], cl["FISS+"]["p50_ms"], d["value"]/1e6, d["value_winner_only"]/1e6, d["e2e"]["value"]/1e6))
PY
