#!/usr/bin/env python
"""Trains of back-to-back launches of the cfg4 step (no events between the launches): lattice kernel alone and the whole plan step,
materialising and winner-only.  Run with FISS_CHAIN=0 / 1 (and FISS_NO_GRAPH=1 for direct launches of the plan step)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiss_plus_planner_b200 import synthetic as syn  # noqa: E402
from fiss_plus_planner_b200.engine import FissEngine, fop_grid, make_params  # noqa: E402
from fiss_plus_planner_b200.planners.common.cost.cost_function import CostFunction  # noqa: E402
from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle  # noqa: E402
from fiss_plus_planner_b200.planners.frenet_optimal_planner import FrenetOptimalPlannerSettings  # noqa: E402

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
f64 = torch.float64
ego_t = torch.tensor(sc.ego, dtype=f64, device=dev)
cost_t = torch.empty(B * Cn, dtype=f64, device=dev)
flags_t = torch.empty(B * Cn, dtype=torch.int32, device=dev)
mat_t = torch.empty((5, B * Cn, ns), dtype=f64, device=dev)
bidx_t = torch.empty(B, dtype=torch.int32, device=dev)
bcost_t = torch.empty(B, dtype=f64, device=dev)
rec_t = torch.empty((B, 16, ns), dtype=f64, device=dev)
meta_t = torch.empty((B, 2), dtype=torch.int32, device=dev)
stream = torch.cuda.Stream() if os.environ.get("PROBE_STREAM") else torch.cuda.current_stream()
torch.cuda.set_stream(stream)
s = stream.cuda_stream
K = 100


def train(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(K):
        fn()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / K


print("env: FISS_CHAIN=%s FISS_NO_GRAPH=%s" % (os.environ.get("FISS_CHAIN"), os.environ.get("FISS_NO_GRAPH")))
for label, mat in (("materialising", mat_t), ("winner-only", None)):
    k = train(lambda: eng.eval_grid_dev(ego_t, grid, prm, cost_t, flags_t, mat, ns, stream=s))
    p = train(lambda: eng.plan_grid_dev(ego_t, grid, prm, cost_t, flags_t, mat, bidx_t, bcost_t, meta_t, rec_t, ns, stream=s))
    print(f"  {label:14s} lattice kernel train {k:.4f} ms/launch   plan step train {p:.4f} ms/step  ({B * Cn / p / 1e6:.0f} M cand/s)")
ref = bidx_t.clone()
eng.plan_grid_dev(ego_t, grid, prm, cost_t, flags_t, mat_t, bidx_t, bcost_t, meta_t, rec_t, ns, stream=s)
torch.cuda.synchronize()
print("  winners stable:", bool((ref == bidx_t).all()))
