#!/bin/bash
# one full ncu capture of the materialising lattice kernel (+ launch list); usage: gpurun -- 'bash tools/gpu_ncu.sh TAG'
TAG=${1:-ncu}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fiss_grid_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_grid \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-closed-loop > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out/${TAG}_grid.ncu-rep
