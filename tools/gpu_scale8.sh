#!/bin/bash
# 8-GPU visit: NCCL tests (4 ranks), weak scaling of the default bench at N = 1, 2, 4, 8 (driver's flags), split lattice at 8
TAG=${1:-s8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.txt; nvidia-smi topo -m >> gpurun_out/${TAG}_smi.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rA 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_multi.log
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 "${@:4}" > gpurun_out/${TAG}_$3.json 2> gpurun_out/${TAG}_$3.err; echo "$3 rc=$?"; }
timeout 300 python bench.py --steps 20 --warmup 5 --no-closed-loop > gpurun_out/${TAG}_weak_n1.json 2> gpurun_out/${TAG}_weak_n1.err
run 2 29511 weak_n2 --steps 20 --warmup 5
run 4 29512 weak_n4 --steps 20 --warmup 5
run 8 29513 weak_n8 --steps 20 --warmup 5
run 8 29514 weak_n8_b --steps 20 --warmup 5
run 8 29515 split_n8 --workload cfg5 --split-lattice --steps 50 --warmup 5
run 8 29516 split64_n8 --workload cfg5 --split-lattice --batch-per-gpu 64 --steps 50 --warmup 5
timeout 300 python bench.py --workload cfg5 --split-lattice --batch-per-gpu 64 --steps 50 --warmup 5 > gpurun_out/${TAG}_split64_n1.json 2> gpurun_out/${TAG}_split64_n1.err
timeout 300 python bench.py --workload cfg5 --split-lattice --steps 50 --warmup 5 > gpurun_out/${TAG}_split_n1.json 2> gpurun_out/${TAG}_split_n1.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "n_gpus", d["n_gpus"], "value %.1fM" % (d["value"] / 1e6), "ms/step %.4f (median %.4f)" % (d["ms_per_step"], d.get("ms_per_step_median", 0)), "e2e %.1fM (%.4f ms)" % (d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"]), "sync %.1fM" % (d["e2e"].get("value_sync_call", 0) / 1e6))
    except Exception as e:
        print(f, "FAILED", e)
PY
