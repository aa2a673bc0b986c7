#!/bin/bash
# parity tests + the bench on the headline workload and the two side workloads; usage: gpurun -- 'bash tools/gpu_check.sh TAG'
TAG=${1:-chk}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.log
for w in cfg4 cfg3 cfg5; do
  timeout 300 python bench.py --no-cpu-baseline --steps 100 --workload $w > gpurun_out/${TAG}_$w.json 2>gpurun_out/${TAG}_$w.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_$w.json"))
    print("$w kernel_ms=%.4f frac=%.3f value=%.1fM winner_only=%.1fM e2e=%.1fM (%.4f ms) p50=%s" % (d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["value"]/1e6, d["value_winner_only"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d.get("plan_cycle_p50_ms")))
except Exception as e:
    print("$w FAILED", e)
PY
done
