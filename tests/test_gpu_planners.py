"""Drop-in planner classes (FOP, FOP+, FISS, FISS+) on the GPU vs closed-loop goldens produced by the
reference's own plan() (tests/golden/loop_*.npz): winner index, end state, cost, Stats counters and
the winner's arrays, cycle by cycle, with the ego advanced from the planner's own output."""
import os
import types

import numpy as np
import pytest

from conftest import golden_files, load_golden

pytestmark = pytest.mark.gpu

LOOPS = golden_files("loop_")
RTOL, RTOL_TIGHT = 1e-4, 1e-9
ATOL = {"yaw": 1e-9, "c": 1e-7, "c_d": 1e-5, "c_dd": 1e-3}   # finite-difference chains amplify by 1/dt per level


def _planner(g, method):
    from fiss_plus_planner_b200 import synthetic as syn
    from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
    from fiss_plus_planner_b200.planners.fiss_planner import FissPlanner, FissPlannerSettings
    from fiss_plus_planner_b200.planners.fiss_plus_planner import FissPlusPlanner, FissPlusPlannerSettings
    from fiss_plus_planner_b200.planners.fop_plus_planner import FopPlusPlanner
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import (FrenetOptimalPlanner,
                                                                        FrenetOptimalPlannerSettings, ObstacleTable)
    cls, scls = {"FOP": (FrenetOptimalPlanner, FrenetOptimalPlannerSettings),
                 "FOP+": (FopPlusPlanner, FrenetOptimalPlannerSettings),
                 "FISS": (FissPlanner, FissPlannerSettings),
                 "FISS+": (FissPlusPlanner, FissPlusPlannerSettings)}[method]
    st = scls(*[int(v) for v in g["num_samples"]])
    st.min_t, st.max_t = float(g["min_t"]), float(g["max_t"])
    if method == "FISS+":
        st.time_limit = 1e9       # same as the golden run: refinement never cut by wall-clock
    veh = Vehicle(syn.vehicle_params(l=float(g["ego_l"]), w=float(g["ego_w"]), v_max=float(g["max_speed"]),
                                     a_max=float(g["max_accel"])))
    pl = cls(st, veh)
    pl.generate_frenet_frame(g["centerline"])
    obs = ObstacleTable(g["obs_xyth"], g["obs_lw"], g["obs_valid"], int(g["final_time_step"]))
    return pl, obs


@pytest.mark.parametrize("path", LOOPS, ids=[os.path.basename(p)[:-4] for p in LOOPS])
def test_closed_loop_vs_reference_golden(path):
    from fiss_plus_planner_b200.planners.common.scenario.frenet import FrenetState
    g = load_golden(path)
    method = os.path.basename(path).split("_")[1].replace("plus", "+")
    pl, obs = _planner(g, method)
    e = g["ego"][0]
    fs = FrenetState(0.0, e[0], e[1], e[2], 0.0, e[3], e[4], e[5], 0.0)
    for i in range(len(g["cost"])):
        np.testing.assert_allclose(fs.as_ego6(), g["ego"][i], rtol=1e-10, atol=1e-12)
        best = pl.plan(fs, float(g["max_target_speed"]), obs, i)
        assert best is not None
        np.testing.assert_allclose(best.cost_final, g["cost"][i], rtol=RTOL_TIGHT)
        np.testing.assert_array_equal(np.asarray(best.idx), g["idx"][i])
        st = pl.stats
        assert (st.num_iter, st.num_trajs_generated, st.num_trajs_validated, st.num_collison_checks) == \
            tuple(int(v) for v in g["stats"][i])
        if not np.isnan(g["end"][i]).any():
            es = best.end_state
            np.testing.assert_allclose([es.d, es.s_d, es.t], g["end"][i], rtol=1e-9, atol=1e-12)
        assert len(best.t) == g["n"][i] and len(best.x) == g["n_cart"][i]
        for f in ("t", "s", "s_d", "s_dd", "s_ddd", "d", "d_d", "d_dd", "d_ddd", "x", "y", "yaw", "ds", "c", "c_d", "c_dd"):
            want = g["best_" + f][i]
            want = want[~np.isnan(want)]
            got = np.asarray(getattr(best, f), dtype=np.float64)
            assert got.shape == want.shape, (f, got.shape, want.shape)
            np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL.get(f, 1e-12), err_msg=f"{f} cycle {i}")
        assert len(pl.all_trajs) == i + 1
        fs = best.frenet_state_at_time_step(1)      # planning.py:137-138
        st0 = best.state_at_time_step(1)
        assert np.isfinite([st0.x, st0.y, st0.yaw, st0.v, st0.a]).all()


def test_all_trajs_bundle_is_lazy_and_complete():
    g = load_golden([p for p in LOOPS if "loop_FOP_cfg1" in p][0])
    pl, obs = _planner(g, "FOP")
    from fiss_plus_planner_b200.planners.common.scenario.frenet import FrenetState
    e = g["ego"][0]
    best = pl.plan(FrenetState(0.0, e[0], e[1], e[2], 0.0, e[3], e[4], e[5], 0.0), float(g["max_target_speed"]), obs, 0)
    bundle = pl.all_trajs[0]
    assert len(bundle) == 125 and bundle._items is None          # nothing materialised yet
    tr = bundle[best.lattice_index]                              # what the GIF renderer does (planning.py:352-355)
    np.testing.assert_array_equal(tr.x, best.x)
    np.testing.assert_array_equal(tr.y, best.y)
    assert tr.cost_final == best.cost_final
    assert all(len(t.x) == len(t.y) for t in bundle)


def test_fop_returns_stale_best_when_nothing_survives():
    """frenet_optimal_planner.py:263-270: best_traj is not reset; FOP+ returns None (fop_plus_planner.py:40)."""
    from fiss_plus_planner_b200.planners.common.scenario.frenet import FrenetState
    g = load_golden([p for p in LOOPS if "loop_FOP_cfg1" in p][0])
    pl, obs = _planner(g, "FOP")
    e = g["ego"][0]
    fs = FrenetState(0.0, e[0], e[1], e[2], 0.0, e[3], e[4], e[5], 0.0)
    first = pl.plan(fs, float(g["max_target_speed"]), obs, 0)
    pl.vehicle.max_accel = 1e-3                                   # nothing can pass the accel mask now
    again = pl.plan(fs, float(g["max_target_speed"]), obs, 1)
    assert again is first
    plp, obs = _planner(g, "FOP+")
    plp.vehicle.max_accel = 1e-3
    assert plp.plan(fs, float(g["max_target_speed"]), obs, 0) is None
    assert plp.stats.num_iter == 125


def test_batch_front_end_single_rank_matches_engine():
    """ShardedBatchPlanner / SplitLatticePlanner at world size 1 are the plain engine calls (the N > 1
    reductions are covered on CPU by tests/test_multi_rank_gloo.py)."""
    from fiss_plus_planner_b200 import synthetic as syn
    from fiss_plus_planner_b200.batch import ShardedBatchPlanner, SplitLatticePlanner
    from fiss_plus_planner_b200.engine import FissEngine, fop_grid, make_params
    from fiss_plus_planner_b200.planners.common.cost.cost_function import CostFunction
    from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import FrenetOptimalPlannerSettings
    sc = syn.make_scene("cfg4_batch4096_32obs", batch=16)
    veh = Vehicle(syn.vehicle_params())
    st = FrenetOptimalPlannerSettings(*sc.num_samples)
    st.min_t, st.max_t, st.highest_speed = sc.min_t, sc.max_t, sc.max_target_speed
    eng = FissEngine(0)
    eng.set_spline(sc.spline.device_table())
    eng.set_obstacles(sc.obs.xyth, sc.obs.lw, sc.obs.valid, sc.obs.final_time_step)
    grid = fop_grid(st, veh.w)
    prm = make_params(st, veh, CostFunction("WX1").as_device_weights())
    ref = eng.plan_grid(sc.ego, grid, prm, want_records=True)
    loc = ShardedBatchPlanner(eng, grid, prm).plan_local(sc.ego)
    assert loc["problems"] == (0, 16)
    np.testing.assert_array_equal(loc["best_idx"], ref["best_idx"])
    out = SplitLatticePlanner(eng, grid, prm).plan(sc.ego)
    np.testing.assert_array_equal(out["best_idx"], ref["best_idx"])
    np.testing.assert_array_equal(out["records"], ref["records"])
    # the lattice split by hand into two slabs of lateral rows, picked with the reference's tie rule
    from fiss_plus_planner_b200.engine import LatticeGrid
    lo = LatticeGrid(grid.d[:5], grid.v, grid.T, grid.tick, "dtv")
    hi = LatticeGrid(grid.d[5:], grid.v, grid.T, grid.tick, "dtv")
    a, b = eng.plan_grid(sc.ego, lo, prm), eng.plan_grid(sc.ego, hi, prm)
    off = 5 * grid.strides[0]
    for p in range(16):
        cands = [(a["best_cost"][p], int(a["best_idx"][p])) if a["best_idx"][p] >= 0 else None,
                 (b["best_cost"][p], int(b["best_idx"][p]) + off) if b["best_idx"][p] >= 0 else None]
        cands = [c for c in cands if c is not None]
        want = max((c for c in cands if c[0] == min(x[0] for x in cands)), key=lambda c: c[1])[1] if cands else -1
        assert want == int(ref["best_idx"][p])
