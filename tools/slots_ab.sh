#!/bin/bash
# parity suite with the lattice kernel forced to 1 / 2 / 3 slots per item, then the bench at each setting
mkdir -p gpurun_out
for s in 2 3 1; do
  FISS_GRID_SLOTS=$s timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/slots${s}_pytest.log
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/slots_default_pytest.log
for s in 1 2 3; do
  FISS_GRID_SLOTS=$s timeout 300 python bench.py --no-cpu-baseline --steps 100 > gpurun_out/slots${s}.json 2>gpurun_out/slots${s}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/slots${s}.json"))
    print("slots=$s kernel_ms=%.4f frac=%.3f value=%.1fM winner_only=%.1fM e2e=%.1fM p50=%.4f" % (d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["value"]/1e6, d["value_winner_only"]/1e6, d["e2e"]["value"]/1e6, d["plan_cycle_p50_ms"]))
except Exception as e:
    print("slots=$s FAILED", e)
PY
done
