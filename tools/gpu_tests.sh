#!/bin/bash
# parity tests + smoke; usage: gpurun -- 'bash tools/gpu_tests.sh TAG'
TAG=${1:-t}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.log
