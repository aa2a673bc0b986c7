#!/bin/bash
# e2e of fiss_plan_grid_host with the batch pipelined in pieces of FISS_SPLIT_PART problems (bench workload, B = 512)
mkdir -p gpurun_out
for part in 512 256 170 128; do
  FISS_SPLIT_PART=$part timeout 300 python bench.py --no-cpu-baseline --steps 200 > gpurun_out/split_$part.json 2>gpurun_out/split_$part.err
  python - <<PY
import json
d = json.load(open("gpurun_out/split_$part.json"))
print("part=$part value=%.1fM winner_only=%.1fM e2e=%.1fM (%.4f ms) pageable=%.1fM p50=%.4f" % (d["value"]/1e6, d["value_winner_only"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e"]["value_pageable_buffers"]/1e6, d["plan_cycle_p50_ms"]))
PY
done
