#!/bin/bash
# per-warp stage timelines of the lattice kernel (build/other/libfiss_trace.so): the static deal, then the defaults
# usage: gpurun -- 'bash tools/gpu_trace.sh TAG'
TAG=${1:-tr}
mkdir -p gpurun_out
lib=$PWD/build/other/libfiss_trace.so
echo "#### static deal (FISS_CHAIN=0 FISS_DYN=0 FISS_DYN_MAT=0)" > gpurun_out/${TAG}_trace.txt
FISSGPU_LIB=$lib FISS_CHAIN=0 FISS_DYN=0 FISS_DYN_MAT=0 timeout 300 python tools/warp_trace.py gpurun_out/${TAG}_static.npz >> gpurun_out/${TAG}_trace.txt 2>&1
echo "#### defaults (work drawn through the device counter; single launches: nothing to chain to)" >> gpurun_out/${TAG}_trace.txt
FISSGPU_LIB=$lib timeout 300 python tools/warp_trace.py gpurun_out/${TAG}_default.npz >> gpurun_out/${TAG}_trace.txt 2>&1
cat gpurun_out/${TAG}_trace.txt
