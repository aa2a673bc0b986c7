#!/usr/bin/env python
"""Per-warp stage timeline of the lattice kernel (debug build with -DFISS_TRACE; GPU box):

    python -m fiss_plus_planner_b200.build --out=build/other/libfiss_trace.so -DFISS_TRACE
    FISSGPU_LIB=$PWD/build/other/libfiss_trace.so python tools/warp_trace.py [out.npz]

Lane 0 of every warp stamps clock64() at the stage boundaries of its first items; this prints, for the materialising and
the winner-only variants of the cfg4 workload, how long a warp works and waits in each stage, how long an item takes, and how
the CTAs that share an SM are phased against each other."""
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

CTAS, WARPS, ITEMS, STAMPS = 512, 16, 8, 16
SEG = [("A rows: work", 0, 1), ("A rows: barrier wait", 1, 2), ("A' box test: work", 2, 3), ("A' box test: wait", 3, 4),
       ("A' per-row test: work", 4, 5), ("A' per-row test: wait", 5, 6), ("B collision: work", 6, 7),
       ("B materialisation: work", 7, 8), ("B: barrier wait", 8, 9), ("C: work", 9, 10)]

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
lib.fiss_debug_trace.argtypes = [C.c_void_p, C.c_int64]
N = CTAS * WARPS * ITEMS * STAMPS
buf = np.zeros(N, dtype=np.int64)
s = torch.cuda.current_stream().cuda_stream
save = {}
for label, mat in (("materialising", mat_t), ("winner-only", None)):
    for _ in range(5):
        eng.eval_grid_dev(ego_t, grid, prm, cost_t, flags_t, mat, ns, stream=s)
    torch.cuda.synchronize()
    lib.fiss_debug_trace(buf.ctypes.data, N)  # clear
    eng.eval_grid_dev(ego_t, grid, prm, cost_t, flags_t, mat, ns, stream=s)
    torch.cuda.synchronize()
    lib.fiss_debug_trace(buf.ctypes.data, N)
    t = buf.reshape(CTAS, WARPS, ITEMS, STAMPS).astype(np.float64)
    save[label] = t.copy()
    live_cta = t[:, 0, 0, 14] > 0
    n_cta = int(live_cta.sum())
    n_warp = int((t[0, :, 0, 14] > 0).sum())
    t = t[:n_cta, :n_warp]
    smid = t[:, 0, 0, 15].astype(int)
    entry = t[:, :, 0, 14]
    items_per_cta = (t[:, 0, :, 9] > 0).sum(axis=1)
    print(f"== {label}: {n_cta} CTAs x {n_warp} warps on {len(set(smid))} SMs; items per CTA {np.bincount(items_per_cta)[1:]} (1, 2, ...)")
    # kernel span per SM clock: entry of the first warp to the last stamp of the CTA
    last = t[:, :, :, :11].max(axis=(1, 2, 3))
    span = last - entry.min(axis=1)
    print(f"   CTA lifetime: mean {span.mean():.0f} cycles, max {span.max():.0f} ({span.max() / 1.965e3:.1f} us at 1.965 GHz); "
          f"entry -> first item start (prologue): {(t[:, :, 0, 0] - entry).mean():.0f} cycles")
    print(f"   CTA lifetime percentiles 0/10/50/90/100: {np.percentile(span, [0, 10, 50, 90, 100]).astype(int)}")
    for it in range(2):
        ok = t[:, 0, it, 9] > 0
        if not ok.any():
            continue
        ti = t[ok][:, :, it, :]
        dur = ti[:, :, 9].max(axis=1) - ti[:, :, 0].min(axis=1)
        print(f"   item {it}: {int(ok.sum())} CTAs, item time (first warp in -> barrier after B) mean {dur.mean():.0f} cycles")
        for name, a, b in SEG:
            if (ti[:, :, a] > 0).all() and (ti[:, :, b] > 0).all():
                d = ti[:, :, b] - ti[:, :, a]
                print(f"      {name:26s} mean {d.mean():7.0f}  max-over-warps {d.max(axis=1).mean():7.0f}  min-over-warps {d.min(axis=1).mean():7.0f}")
    # phasing of the CTAs that share an SM: overlap of their stage-B intervals (stamp 6 .. 9) in item 1
    by_sm = {}
    for c in range(n_cta):
        by_sm.setdefault(smid[c], []).append(c)
    offs, both = [], []
    for sm, cs in by_sm.items():
        if len(cs) < 2:
            continue
        for it in (0, 1):
            a0, a1 = t[cs[0], :, it, 6].min(), t[cs[0], :, it, 9].max()
            b0, b1 = t[cs[1], :, it, 6].min(), t[cs[1], :, it, 9].max()
            if min(a0, b0) <= 0:
                continue
            ov = max(0.0, min(a1, b1) - max(a0, b0))
            offs.append(abs(a0 - b0))
            both.append(ov / max(1.0, min(a1 - a0, b1 - b0)))
    if offs:
        print(f"   co-resident CTAs (first two per SM): start of stage B differs by mean {np.mean(offs):.0f} cycles; their stage-B intervals "
              f"overlap by {100 * np.mean(both):.0f} % of the shorter one")
if len(sys.argv) > 1:
    np.savez_compressed(sys.argv[1], **{k.replace("-", "_"): v for k, v in save.items()})
