#!/bin/bash
# per-warp stage timelines of the lattice kernel for every build/other/libfiss_trace*.so; usage: gpurun -- 'bash tools/gpu_trace.sh TAG'
TAG=${1:-tr}
mkdir -p gpurun_out
for lib in build/other/libfiss_trace*.so; do
  n=$(basename $lib .so)
  FISSGPU_LIB=$PWD/$lib timeout 300 python tools/warp_trace.py gpurun_out/${TAG}_$n.npz > gpurun_out/${TAG}_$n.txt 2>&1
  echo "#### $n"; cat gpurun_out/${TAG}_$n.txt
done
