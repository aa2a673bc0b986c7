#!/bin/bash
# the default bench, both arms, as the driver runs them (plus the driver-style 20-step line)
TAG=${1:-fin}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_s20.json 2>> gpurun_out/${TAG}_bench.err
python - <<PY
import json
for f in ("bench", "bench_s20"):
    d = json.load(open("gpurun_out/${TAG}_%s.json" % f)); r = d["roofline"]
    print(f, "value %.1fM ms/step %.4f (alone %.4f) winner_only %.1fM e2e %.1fM (%.4f ms) sync %.1fM | kernel_ms %.4f frac %.3f alone %.4f frac_alone %.3f | p50 %.4f launches %d" % (d["value"]/1e6, d["ms_per_step"], d["ms_per_step_median"], d["value_winner_only"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e"]["value_sync_call"]/1e6, r["kernel_ms"], r["frac"], r["kernel_ms_alone"], r["frac_alone"], d["plan_cycle_p50_ms"], d["gpu_launches"]))
    if d.get("cpu_baseline"): print("   cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:40])
    if d.get("closed_loop"): print("   closed loop", {k: v["p50_ms"] for k, v in d["closed_loop"].items() if isinstance(v, dict) and "p50_ms" in v})
d = json.load(open("gpurun_out/${TAG}_bench_ref.json")); print("reference arm", d["value"], d["steps"], d["cpu_baseline"]["kind"])
PY
