#!/bin/bash
# Profile evidence for one round: launch list of a bench run + full ncu captures of the three kernels of a plan step
# usage: gpurun -- 'bash tools/gpu_profile.sh TAG'
TAG=${1:-prof}
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-closed-loop"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fiss_grid_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_grid $B > gpurun_out/${TAG}_ncu_grid.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fiss_record_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_rec $B > gpurun_out/${TAG}_ncu_rec.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fiss_grid_kernel -s 40 -c 1 -f -o gpurun_out/${TAG}_grid_b1 python tools/p50_probe.py > gpurun_out/${TAG}_ncu_b1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fiss_record_kernel -s 40 -c 1 -f -o gpurun_out/${TAG}_rec_b1 python tools/p50_probe.py > gpurun_out/${TAG}_ncu_rec_b1.log 2>&1
ls -la gpurun_out/${TAG}_*.ncu-rep; wc -l gpurun_out/${TAG}_launches.csv
# the winner-only lattice kernel at B = 512 (the e2e path): the 15th lattice launch of the bench is one
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fiss_grid_kernel -s 28 -c 1 -f -o gpurun_out/${TAG}_grid0 $B > gpurun_out/${TAG}_ncu_grid0.log 2>&1
ls -la gpurun_out/${TAG}_grid0.ncu-rep
