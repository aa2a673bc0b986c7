#!/usr/bin/env python
"""Where a plan cycle of ONE ego state (BASELINE config 2) spends its time (GPU box):
wall p50 of the Python call, of the bare C call, and the GPU-side span between events around the call's stream work."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiss_plus_planner_b200 import _shim, synthetic as syn  # noqa: E402
from fiss_plus_planner_b200.engine import FissEngine, fop_grid, make_params  # noqa: E402
from fiss_plus_planner_b200.planners.common.cost.cost_function import CostFunction  # noqa: E402
from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle  # noqa: E402
from fiss_plus_planner_b200.planners.frenet_optimal_planner import FrenetOptimalPlannerSettings  # noqa: E402

sc = syn.make_scene("cfg2_single_ego_8obs", batch=1)
veh = Vehicle(syn.vehicle_params())
st = FrenetOptimalPlannerSettings(9, 6, 5)
st.min_t, st.max_t, st.highest_speed = 4.0, 5.0, sc.max_target_speed
eng = FissEngine(0)
eng.set_spline(sc.spline.device_table())
eng.set_obstacles(sc.obs.xyth, sc.obs.lw, sc.obs.valid, sc.obs.final_time_step)
grid = fop_grid(st, veh.w)
prm = make_params(st, veh, CostFunction("WX1").as_device_weights())
stream = torch.cuda.current_stream()
s = stream.cuda_stream
ego = np.ascontiguousarray(sc.ego[:1])
out = eng.alloc_plan_outputs(1, grid, want_records=True, want_volume=True, pinned=False)
lib = _shim.load()


def p50(fn, n=400, warm=30):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return 1e6 * float(np.median(ts)), 1e6 * float(np.percentile(ts, 90))


print("graphs:", "off" if os.environ.get("FISS_NO_GRAPH") else "on", " pdl edge:", "off" if os.environ.get("FISS_GRAPH_NO_PDL") else "on")
print("python plan_grid (allocating outputs)   p50 %.1f us  p90 %.1f" % p50(lambda: eng.plan_grid(ego, grid, prm, True, True, stream=s)))
print("python plan_grid (out= reused)          p50 %.1f us  p90 %.1f" % p50(lambda: eng.plan_grid(ego, grid, prm, True, True, stream=s, out=out)))
args = (eng._h, C.c_void_p(s), _shim.ptr(ego), 1, C.byref(grid.c_struct), C.byref(prm), _shim.ptr(out["best_idx"]),
        _shim.ptr(out["best_cost"]), _shim.ptr(out["meta"]), _shim.ptr(out["records"]), grid.n_stride, _shim.ptr(out["cost"]),
        _shim.ptr(out["flags"]))
print("bare ctypes fiss_plan_grid_host         p50 %.1f us  p90 %.1f" % p50(lambda: lib.fiss_plan_grid_host(*args)))
args2 = args[:9] + (None, grid.n_stride, None, None)
print("  ... winners only (no records/volume)  p50 %.1f us  p90 %.1f" % p50(lambda: lib.fiss_plan_grid_host(*args2)))
# GPU-side span of one call's stream work
spans = []
for _ in range(230):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    lib.fiss_plan_grid_host(*args)
    e1.record(stream)
    torch.cuda.synchronize()
    spans.append(e0.elapsed_time(e1) * 1e3)
print("GPU span between events around the call p50 %.1f us" % float(np.median(spans[30:])))
# device-resident kernels only
dev = torch.device("cuda", 0)
Cn, ns = grid.num_candidates, grid.n_stride
ego_t = torch.tensor(ego, dtype=torch.float64, device=dev)
cost_t = torch.empty(Cn, dtype=torch.float64, device=dev)
flags_t = torch.empty(Cn, dtype=torch.int32, device=dev)
idx_t = torch.empty(1, dtype=torch.int32, device=dev)
best_t = torch.empty(1, dtype=torch.float64, device=dev)
meta_t = torch.empty((1, 2), dtype=torch.int32, device=dev)
rec_t = torch.empty((1, 16, ns), dtype=torch.float64, device=dev)
end_t = torch.tensor(grid.table(), dtype=torch.float64, device=dev)
for label, fn in (("lattice kernel alone", lambda: eng.eval_grid_dev(ego_t, grid, prm, cost_t, flags_t, None, ns, stream=s)),
                  ("record kernel alone (pick fused)", lambda: eng.pick_winners_dev(ego_t, end_t, prm, cost_t, flags_t, idx_t, best_t, rec_t, meta_t, ns, stream=s)),
                  ("plan_grid_dev (both, one call)", lambda: eng.plan_grid_dev(ego_t, grid, prm, cost_t, flags_t, None, idx_t, best_t, meta_t, rec_t, ns, stream=s))):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    spans = []
    for _ in range(200):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        spans.append(e0.elapsed_time(e1) * 1e3)
    print("GPU span: %-34s p50 %.1f us" % (label, float(np.median(spans))))
