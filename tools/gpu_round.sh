#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list, one full capture of the lattice kernel.
# usage (from repo root): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG'
TAG=${1:-r1x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fiss_grid_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_grid \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log; tail -2 gpurun_out/${TAG}_smoke.log; cat gpurun_out/${TAG}_bench.json
