#!/bin/bash
# per-stage cycle split of the lattice kernel + a short bench; usage: gpurun -- 'bash tools/gpu_phase.sh TAG'
TAG=${1:-ph}
mkdir -p gpurun_out
FISSGPU_LIB=$PWD/build/other/libfiss_phase.so timeout 300 python tools/phase_timing.py > gpurun_out/${TAG}_phase.txt 2>&1
cat gpurun_out/${TAG}_phase.txt
timeout 300 python bench.py --no-cpu-baseline --no-closed-loop --steps 100 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print("kernel_ms=%.4f frac=%.3f value=%.1fM winner_only=%.1fM e2e=%.1fM p50=%.4f" % (d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["value"]/1e6, d["value_winner_only"]/1e6, d["e2e"]["value"]/1e6, d["plan_cycle_p50_ms"]))
PY
