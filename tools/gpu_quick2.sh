#!/bin/bash
# new tests + a bench (graph on / off); usage: gpurun -- 'bash tools/gpu_quick2.sh TAG [pytest-args]'
TAG=${1:-q}
shift
mkdir -p gpurun_out
timeout 900 python -m pytest -m gpu -x -q "$@" 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest.log
for mode in graph nograph; do
  if [ $mode == nograph ]; then export FISS_NO_GRAPH=1; else unset FISS_NO_GRAPH; fi
  timeout 300 python bench.py --no-cpu-baseline --no-closed-loop --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_$mode.json 2> gpurun_out/${TAG}_bench_$mode.err
  tail -3 gpurun_out/${TAG}_bench_$mode.err
  python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_$mode.json"))
print("$mode kernel_ms=%.4f frac=%.3f value=%.1fM (%.4f ms, median %.4f) winner_only=%.1fM e2e=%.1fM (%.4f ms) sync=%.1fM p50=%.4f launches=%d" % (d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["value"]/1e6, d["ms_per_step"], d["ms_per_step_median"], d["value_winner_only"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e"]["value_sync_call"]/1e6, d["plan_cycle_p50_ms"], d["gpu_launches"]))
PY
done
