#!/bin/bash
# N-GPU visit: the NCCL tests, the weak-scaling bench at N, the split-lattice (strong-scaling) bench at 1 and N
N=${1:-2}; TAG=${2:-mg}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rA 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_multi.log
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 "${@:4}" > gpurun_out/${TAG}_$3.json 2> gpurun_out/${TAG}_$3.err; echo "$3 rc=$?"; }
timeout 300 python bench.py --no-cpu-baseline --no-closed-loop --steps 20 --warmup 5 > gpurun_out/${TAG}_weak_n1.json 2> gpurun_out/${TAG}_weak_n1.err
run $N 29517 weak_n$N --steps 20 --warmup 5
timeout 300 python bench.py --workload cfg5 --split-lattice --steps 50 --warmup 5 > gpurun_out/${TAG}_split_n1.json 2> gpurun_out/${TAG}_split_n1.err
run $N 29518 split_n$N --workload cfg5 --split-lattice --steps 50 --warmup 5
run $N 29519 split1_n$N --workload cfg5 --split-lattice --batch-per-gpu 1 --steps 50 --warmup 5
timeout 300 python bench.py --workload cfg5 --split-lattice --batch-per-gpu 1 --steps 50 --warmup 5 > gpurun_out/${TAG}_split1_n1.json 2> gpurun_out/${TAG}_split1_n1.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "n_gpus", d["n_gpus"], "value %.1fM" % (d["value"] / 1e6), "ms/step %.4f (median %.4f)" % (d["ms_per_step"], d.get("ms_per_step_median", 0)), "e2e %.1fM (%.4f ms)" % (d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"]), d.get("winner_ids", ""))
    except Exception as e:
        print(f, "FAILED", e)
PY
for f in gpurun_out/${TAG}_*.err; do tail -n 3 $f; done | tail -n 20
