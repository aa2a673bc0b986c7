#!/bin/bash
# compute-sanitizer passes over the parity tests that exercise every kernel (GPU box).  usage: gpurun -- 'bash tools/sanitize.sh TAG'
TAG=${1:-san}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
SEL="tests/test_gpu_planners.py tests/test_gpu_parity.py::test_grid_matches_generic_kernel tests/test_gpu_parity.py::test_dense_materialisation_matches_records tests/test_gpu_spline_fit.py::test_batched_lanes_are_independent tests/test_gpu_robustness.py tests/test_gpu_waymo.py tests/test_gpu_graph_stream.py tests/test_gpu_sat_known_answers.py::test_many_obstacles_one_toucher tests/test_gpu_multi.py::test_allreduce_pick_single_rank_is_identity_plus_offset tests/test_gpu_r2_goldens.py::test_heading_wraps_like_arctan2_on_a_westbound_road tests/test_gpu_chained.py::test_shared_volume_winners_per_step tests/test_gpu_chained.py::test_lattice_launches_alone_chain_too"
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1200 $CS --tool $tool --error-exitcode 7 --print-limit 20 python -m pytest $SEL -x -q > gpurun_out/${TAG}_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/${TAG}_$tool.log | tr '\n' ' ')"
done
