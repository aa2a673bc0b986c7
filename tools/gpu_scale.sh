#!/bin/bash
# scaling sweep on one box: bench at N = 1, 2, 4, ... up to $1 GPUs (as the driver launches it), reference arm at N.
NMAX=${1:-8}; TAG=${2:-sc}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.txt
for N in 1 2 4 8; do
  [ $N -gt $NMAX ] && break
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 200 --warmup 3 > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 200 --warmup 3 > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
  fi
  echo "N=$N rc=$? lines=$(wc -l < gpurun_out/${TAG}_n$N.json)"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port 29700 bench.py --impl reference --gpus $NMAX --steps 5 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "ref rc=$? lines=$(wc -l < gpurun_out/${TAG}_ref.json)"
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
python - <<PY
import json, glob
base = None
for f in sorted(glob.glob("gpurun_out/${TAG}_n*.json"), key=lambda s: int(s.split("_n")[-1][:-5])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        base = base or d["value"]
        print(f, "n_gpus", d["n_gpus"], "value %.1fM (x%.2f)" % (d["value"] / 1e6, d["value"] / base), "e2e %.1fM" % (d["e2e"]["value"] / 1e6), "ms/step %.4f" % d["ms_per_step"])
    except Exception as e:
        print(f, "FAILED", e)
PY
