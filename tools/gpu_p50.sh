#!/bin/bash
TAG=${1:-p50}
mkdir -p gpurun_out
( python tools/p50_probe.py; FISS_GRAPH_CAPTURE=1 python tools/p50_probe.py; FISS_NO_GRAPH=1 python tools/p50_probe.py ) > gpurun_out/${TAG}_p50.txt 2>&1
cat gpurun_out/${TAG}_p50.txt
