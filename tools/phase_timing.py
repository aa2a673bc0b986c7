#!/usr/bin/env python
"""Per-stage cycle split of the lattice kernel (debug build with -DFISS_PHASE_TIMING; GPU box):

    python -m fiss_plus_planner_b200.build --out=build/other/libfiss_phase.so -DFISS_PHASE_TIMING
    FISSGPU_LIB=$PWD/build/other/libfiss_phase.so python tools/phase_timing.py

Thread 0 of every CTA accumulates clock64() deltas between the stage barriers; the sums over CTAs are printed as
shares of the CTA-resident time, for the materialising and the winner-only variants of the cfg4 workload."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiss_plus_planner_b200 import _shim, synthetic as syn  # noqa: E402
from fiss_plus_planner_b200.engine import FissEngine, fop_grid, make_params  # noqa: E402
from fiss_plus_planner_b200.planners.common.cost.cost_function import CostFunction  # noqa: E402
from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle  # noqa: E402
from fiss_plus_planner_b200.planners.frenet_optimal_planner import FrenetOptimalPlannerSettings  # noqa: E402

NAMES = ["prologue (TMA staging)", "stage C + item set-up", "A rows", "A' boxes", "A' box test", "A' per-row test + list",
         "B collision", "B materialisation"]
sc = syn.make_scene("cfg4_batch4096_32obs", batch=512, num_obstacles=32)
veh = Vehicle(syn.vehicle_params())
st = FrenetOptimalPlannerSettings(9, 6, 5)
st.min_t, st.max_t, st.highest_speed = 4.0, 5.0, sc.max_target_speed
eng = FissEngine(0)
eng.set_spline(sc.spline.device_table())
eng.set_obstacles(sc.obs.xyth, sc.obs.lw, sc.obs.valid, sc.obs.final_time_step)
grid = fop_grid(st, veh.w)
prm = make_params(st, veh, CostFunction("WX1").as_device_weights())
dev = torch.device("cuda", 0)
B, Cn, ns = 512, grid.num_candidates, grid.n_stride
ego_t = torch.tensor(sc.ego, dtype=torch.float64, device=dev)
cost_t = torch.empty(B * Cn, dtype=torch.float64, device=dev)
flags_t = torch.empty(B * Cn, dtype=torch.int32, device=dev)
mat_t = torch.empty((5, B * Cn, ns), dtype=torch.float64, device=dev)
lib = _shim.load()
buf = (C.c_longlong * 8)()
s = torch.cuda.current_stream().cuda_stream
for label, mat in (("materialising", mat_t), ("winner-only", None)):
    for _ in range(3):
        eng.eval_grid_dev(ego_t, grid, prm, cost_t, flags_t, mat, ns, stream=s)
    torch.cuda.synchronize()
    lib.fiss_debug_phase_cycles(buf)
    reps = 20
    for _ in range(reps):
        eng.eval_grid_dev(ego_t, grid, prm, cost_t, flags_t, mat, ns, stream=s)
    torch.cuda.synchronize()
    lib.fiss_debug_phase_cycles(buf)
    cyc = np.array(list(buf), dtype=np.float64) / reps
    tot = cyc.sum()
    print(f"== {label}: {tot / 444:.0f} cycles per CTA ({tot / 444 / 1.965e3:.1f} us at 1.965 GHz), 2560 items over 444 CTAs")
    for n, c in zip(NAMES, cyc):
        print(f"   {n:28s} {100 * c / tot:5.1f} %   {c / 2560:8.0f} cycles / item")
