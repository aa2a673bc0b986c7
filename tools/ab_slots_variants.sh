#!/bin/bash
# A/B of build variants x slots per item; usage: gpurun -- 'bash tools/ab_slots_variants.sh TAG "1 2"'
TAG=${1:-abs}
SLOTS=${2:-"1 2"}
mkdir -p gpurun_out
for lib in build/libfiss_*.so; do
  n=$(basename $lib .so)
  for s in $SLOTS; do
    FISS_GRID_SLOTS=$s FISSGPU_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu-baseline --steps 100 > gpurun_out/${TAG}_${n}_s$s.json 2>gpurun_out/${TAG}_${n}_s$s.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${n}_s$s.json"))
    print("$n slots=$s kernel_ms=%.4f frac=%.3f value=%.1fM winner_only=%.1fM e2e=%.1fM" % (d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["value"]/1e6, d["value_winner_only"]/1e6, d["e2e"]["value"]/1e6))
except Exception as e:
    print("$n slots=$s FAILED", e)
PY
  done
done
