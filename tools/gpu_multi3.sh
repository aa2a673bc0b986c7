#!/bin/bash
# short N-GPU visit: the NCCL tests and the weak-scaling bench at N (driver-style 20 steps)
N=${1:-2}; TAG=${2:-mg}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rA 2>&1 | tail -12 | tee gpurun_out/${TAG}_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_weak_n$N.json 2> gpurun_out/${TAG}_weak_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_weak_n$N.json").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value %.1fM" % (d["value"] / 1e6), "ms/step %.4f" % d["ms_per_step"], "winner_only %.1fM" % (d["value_winner_only"] / 1e6), "e2e %.1fM (%.4f ms)" % (d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"]), "kernel_ms", d["roofline"]["kernel_ms"])
PY
tail -3 gpurun_out/${TAG}_weak_n$N.err
