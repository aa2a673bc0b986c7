#!/bin/bash
# driver-style weak-scaling line at N GPUs (20 steps, 5 warm-up); usage: gpurun --gpus N -- 'bash tools/gpu_scale_quick.sh N TAG'
N=${1:-8}; TAG=${2:-sc}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_n$N.json").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value %.1fM" % (d["value"] / 1e6), "ms/step %.4f (alone median %.4f)" % (d["ms_per_step"], d["ms_per_step_median"]), "winner_only %.1fM" % (d["value_winner_only"] / 1e6), "e2e %.1fM (%.4f ms)" % (d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"]), "kernel_ms %.4f alone %.4f" % (d["roofline"]["kernel_ms"], d["roofline"]["kernel_ms_alone"]))
PY
tail -2 gpurun_out/${TAG}_n$N.err
