#!/bin/bash
# N-GPU visit: all gpu tests (incl. the NCCL one), bench at N=1 and N=$1 through torchrun, reference arm under torchrun.
N=${1:-2}; TAG=${2:-mg}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"
timeout 600 python bench.py --gpus 1 --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench N=1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/${TAG}_ref_n$N.json 2> gpurun_out/${TAG}_ref_n$N.err; echo "ref N=$N rc=$?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_n1.json", "gpurun_out/${TAG}_bench_n$N.json", "gpurun_out/${TAG}_ref_n$N.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("impl", "ours"), "n_gpus", d["n_gpus"], "value %.1fM" % (d["value"] / 1e6), "e2e %.1fM" % (d["e2e"]["value"] / 1e6), "ms/step", d["ms_per_step"])
    except Exception as e:
        print(f, "FAILED", e)
PY
