"""the reference line fitted and resampled on the device (Thomas recurrence) against the host
path, which assembles and solves the reference's own dense system (cubic_spline.py:19-43,118-142): coefficient
tables, the 0.1 m polyline of generate_frenet_frame, and a whole plan() on the device-fitted line."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, load_golden

pytestmark = pytest.mark.gpu


def _lanes():
    from fiss_plus_planner_b200 import synthetic as syn
    lanes = {"synthetic81": syn.reference_line(), "short13": syn.reference_line(13), "two_points": syn.reference_line(2)}
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "driver_FOP_*.npz"))):
        lanes[os.path.basename(p)[len("driver_FOP_"):-4]] = load_golden(p)["centerline"][:, :2]   # real CommonRoad routes
    return lanes


@pytest.mark.parametrize("name", list(_lanes()))
def test_device_fit_matches_host_fit(name):
    from fiss_plus_planner_b200.engine import FissEngine
    from fiss_plus_planner_b200.planners.common.geometry.cubic_spline import CubicSpline2D
    pts = _lanes()[name]
    eng = FissEngine(0)
    table = eng.fit_splines(pts, install=0)[0]
    host = CubicSpline2D(pts[:, 0], pts[:, 1])
    want = host.device_table()
    assert table.shape == want.shape
    np.testing.assert_allclose(table[0], want[0], rtol=1e-14)                       # knots (cumsum of hypot)
    scale = np.abs(want).max(axis=1, keepdims=True) + 1e-300
    np.testing.assert_allclose(table / scale, want / scale, rtol=0, atol=1e-11)     # coefficient rows, row-relative
    if len(pts) > 2:
        ref = eng.frame_samples(host.s[-1], 0.1)
        s = np.arange(0, host.s[-1], 0.1)
        assert len(ref) == len(s)
        xy = np.array([host.calc_position(v) for v in s])
        np.testing.assert_allclose(ref[:, :2], xy, rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(ref[:, 2], [host.calc_yaw(v) for v in s], rtol=0, atol=1e-10)
        np.testing.assert_allclose(ref[:, 3], [host.calc_curvature(v) for v in s], rtol=1e-6, atol=1e-10)
        dev = CubicSpline2D.from_device_table(table)                                 # host view of the device fit
        np.testing.assert_allclose(dev.calc_position(0.37 * host.s[-1]), host.calc_position(0.37 * host.s[-1]), rtol=1e-12)
        assert dev.calc_position(host.s[-1] + 1.0) == (None, None)


def test_batched_lanes_are_independent():
    from fiss_plus_planner_b200.engine import FissEngine
    rng = np.random.default_rng(3)
    k = 40
    base = np.column_stack((np.arange(k) * 4.0, np.zeros(k)))
    lanes = np.stack([base + np.column_stack((rng.uniform(-0.5, 0.5, k), 3.0 * np.sin(np.arange(k) / (3.0 + i)) + 3.5 * i))
                      for i in range(16)])
    eng = FissEngine(0)
    tabs = eng.fit_splines(lanes, install=5)
    for i in (0, 5, 15):
        np.testing.assert_array_equal(tabs[i], eng.fit_splines(lanes[i], install=-1)[0])   # same arithmetic alone or in a batch
    # the installed lane is what the samples come from
    ref = eng.frame_samples(tabs[5][0, -1], 0.1)
    np.testing.assert_allclose(ref[0, :2], lanes[5][0], rtol=1e-13)


def test_plan_on_device_fitted_line_matches_reference_golden():
    """The whole hot path on a device-fitted reference line vs the reference's outputs: inside the 1e-4 contract
    (the coefficients differ from the LAPACK ones at ~1e-15, so bit-exactness of masks is not claimed here)."""
    from fiss_plus_planner_b200 import synthetic as syn
    from fiss_plus_planner_b200.engine import decode_flags
    from fiss_plus_planner_b200.planners.common.scenario.frenet import FrenetState
    from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import (FrenetOptimalPlanner, FrenetOptimalPlannerSettings,
                                                                        ObstacleTable)
    g = load_golden(os.path.join(GOLDEN_DIR, "dense_cfg2_m8.npz"))
    st = FrenetOptimalPlannerSettings(*[int(v) for v in g["num_samples"]])
    st.min_t, st.max_t = float(g["min_t"]), float(g["max_t"])
    veh = Vehicle(syn.vehicle_params(l=float(g["ego_l"]), w=float(g["ego_w"]), v_max=float(g["max_speed"]), a_max=float(g["max_accel"])))
    pl = FrenetOptimalPlanner(st, veh)
    spline, ref = pl.generate_frenet_frame(g["centerline"], fit="device")
    assert ref.shape[1] == 4 and np.isfinite(ref).all()
    e = g["ego"]
    best = pl.plan(FrenetState(0.0, e[0], e[1], e[2], 0.0, e[3], e[4], e[5], 0.0), float(g["max_target_speed"]),
                   ObstacleTable(g["obs_xyth"], g["obs_lw"], g["obs_valid"], int(g["final_time_step"])), int(g["time_step_now"]))
    bundle = pl.all_trajs[-1]
    np.testing.assert_allclose(bundle.cost, g["cost"], rtol=1e-9)
    ok, coll, n_cart = decode_flags(bundle.flags)
    np.testing.assert_array_equal(n_cart, g["n_cart"])
    np.testing.assert_array_equal(ok, g["constraint_ok"])
    assert best.lattice_index == int(g["best"])
    k = list(g["keep"]).index(int(g["best"]))
    for f, tol in (("x", 1e-9), ("y", 1e-9), ("yaw", 1e-9)):
        want = g["traj_" + f][k]
        want = want[~np.isnan(want)]
        np.testing.assert_allclose(getattr(best, f), want, rtol=1e-9, atol=tol)
