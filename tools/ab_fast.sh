#!/bin/bash
# fast A/B of kernel build variants (build/libfiss_*.so): short bench, no CPU baseline, no closed loop
# usage: gpurun -- 'bash tools/ab_fast.sh TAG [pytest]'
TAG=${1:-ab}
mkdir -p gpurun_out
if [ "$2" == "pytest" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.log
fi
for lib in build/libfiss_*.so; do
  n=$(basename $lib .so)
  FISSGPU_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu-baseline --no-closed-loop --steps 100 > gpurun_out/${TAG}_$n.json 2>gpurun_out/${TAG}_$n.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_$n.json"))
    print("$n kernel_ms=%.4f frac=%.3f value=%.1fM winner_only=%.1fM e2e=%.1fM p50=%.4f" % (d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["value"]/1e6, d["value_winner_only"]/1e6, d["e2e"]["value"]/1e6, d["plan_cycle_p50_ms"]))
except Exception as e:
    print("$n FAILED", e)
PY
done
