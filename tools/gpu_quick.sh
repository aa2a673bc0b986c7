#!/bin/bash
# quick GPU check: parity tests + a short bench (no CPU baseline); usage: gpurun -- 'bash tools/gpu_quick.sh TAG'
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --no-cpu-baseline --steps 100 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print("kernel_ms=%.4f frac=%.3f value=%.1fM winner_only=%.1fM e2e=%.1fM p50=%.4f" % (d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["value"]/1e6, d["value_winner_only"]/1e6, d["e2e"]["value"]/1e6, d["plan_cycle_p50_ms"]))
PY
