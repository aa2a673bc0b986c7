#!/bin/bash
# chained (programmatic dependent) launches of consecutive plan steps: parity suite, launch trains, bench with / without.
# usage: gpurun -- 'bash tools/ab_chain.sh TAG'
TAG=${1:-chain}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.log
FISS_CHAIN=0 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/${TAG}_pytest_nochain.log
FISS_CHAIN=0 python tools/chain_probe.py 2>&1 | tee gpurun_out/${TAG}_probe_off.txt
FISS_CHAIN=1 python tools/chain_probe.py 2>&1 | tee gpurun_out/${TAG}_probe_on.txt
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-closed-loop --steps 100 > gpurun_out/${TAG}_$name.json 2>gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_$name.json"))
    r = d["roofline"]
    print("$name kernel_ms=%.4f (alone %.4f) frac=%.3f value=%.1fM (ms/step %.4f, alone %.4f) winner_only=%.1fM e2e=%.1fM p50=%.4f" % (r["kernel_ms"], r["kernel_ms_alone"], r["frac"], d["value"]/1e6, d["ms_per_step"], d["ms_per_step_median"], d["value_winner_only"]/1e6, d["e2e"]["value"]/1e6, d["plan_cycle_p50_ms"]))
except Exception as e:
    print("$name FAILED", e)
PY
}
run chain FISS_X=0
run nochain FISS_CHAIN=0
