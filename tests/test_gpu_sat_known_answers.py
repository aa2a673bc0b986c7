"""Known-answer tests of the CUDA collision predicate (``rect_sat`` + its circle pre-filters) through the C-ABI, on the
exact-geometry cases of tests/sat_cases.py (touching edges / vertices, containment, one-ulp near misses), for BOTH
kernels: the list kernel (``fiss_eval_end_states_host``) and the lattice kernel (``fiss_plan_grid_host``, 1x1x1 lattice).

Scene: a straight reference line along the x axis (every spline coefficient exact), one candidate that keeps d = 0.5 and
drives at a constant 10 m/s from s = 16, so that at checked step 0 the ego rectangle is EXACTLY the one of sat_cases
(centre (16, 0.5), heading atan2(0, 1) = 0); ``final_time_step = 1`` limits has_collision (:173-176) to that step."""
import numpy as np
import pytest

import sat_cases as sc

pytestmark = pytest.mark.gpu

CASES = sc.cases()


@pytest.fixture(scope="module")
def scene():
    from fiss_plus_planner_b200 import synthetic as syn
    from fiss_plus_planner_b200.engine import FissEngine, LatticeGrid, make_params
    from fiss_plus_planner_b200.planners.common.cost.cost_function import CostFunction
    from fiss_plus_planner_b200.planners.common.geometry.cubic_spline import CubicSpline2D
    from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import FrenetOptimalPlannerSettings
    xs = np.arange(0.0, 205.0, 5.0)
    spline = CubicSpline2D(xs, np.zeros_like(xs))
    tab = spline.device_table()
    # the line is exact: x(s) = s, y = 0
    np.testing.assert_array_equal(tab[2, :-1], 1.0)
    assert not tab[3:5].any() and not tab[5:].any()
    eng = FissEngine(0)
    eng.set_spline(tab)
    veh = Vehicle(syn.vehicle_params(l=sc.EGO_L, w=sc.EGO_W))
    st = FrenetOptimalPlannerSettings(1, 1, 1)
    st.highest_speed = 10.0
    prm = make_params(st, veh, CostFunction("WX1").as_device_weights(), time_step_now=0, collide_all=True)
    ego = np.array([sc.EGO_X, 10.0, 0.0, sc.EGO_Y, 0.0, 0.0])
    grid = LatticeGrid([sc.EGO_Y], [10.0], [1.0], 0.1, "dtv")
    return eng, prm, ego, grid


def _set_obstacle(eng, cx, cy, th, length, width):
    xyth = np.zeros((1, 2, 3))
    xyth[0, :] = (cx, cy, th)
    eng.set_obstacles(xyth, np.array([[length, width]]), np.ones((1, 2), np.uint8), final_time_step=1)


def test_ego_pose_is_exact(scene):
    eng, prm, ego, grid = scene
    _set_obstacle(eng, 1000.0, 1000.0, 0.0, 1.0, 1.0)
    out = eng.eval_end_states(ego, grid.table(), prm, want_records=True)
    rec = out["records"][0]
    np.testing.assert_array_equal(rec[9, :3], [16.0, 17.0, 18.0])     # x
    np.testing.assert_array_equal(rec[10, :3], [0.5, 0.5, 0.5])      # y
    np.testing.assert_array_equal(rec[11, :3], [0.0, 0.0, 0.0])      # yaw
    assert not (out["flags"][0] & 8)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_rect_sat_known_answer(scene, case):
    from fiss_plus_planner_b200 import _shim
    eng, prm, ego, grid = scene
    name, cx, cy, th, length, width, expected = case
    _set_obstacle(eng, cx, cy, th, length, width)
    listed = eng.eval_end_states(ego, grid.table(), prm)
    assert bool(listed["flags"][0] & _shim.FLAG_COLLISION) is expected, "list kernel"
    lattice = eng.plan_grid(ego[None], grid, prm, want_records=False, want_volume=True)
    assert bool(lattice["flags"][0, 0] & _shim.FLAG_COLLISION) is expected, "lattice kernel"
    assert int(lattice["best_idx"][0]) == (-1 if expected else 0)
    # the oracle on the same scene gives the same answer (vertex-projection SAT after shapely's affine maps)
    from oracle import sat_geometry as sat
    ring_e = sat.place(sat.ego_ring(sc.EGO_L, sc.EGO_W), sc.EGO_X, sc.EGO_Y, 0.0)
    assert sat.sat_closed(ring_e, sat.place(sat.obstacle_ring(length, width), cx, cy, th)) is expected


def test_many_obstacles_one_toucher(scene):
    """64 obstacles (two mask words in the lattice kernel), exactly one of which touches the ego at a corner."""
    from fiss_plus_planner_b200 import _shim
    eng, prm, ego, grid = scene
    m = 64
    rng = np.random.default_rng(3)
    xyth = np.zeros((m, 2, 3))
    xyth[:, :, 0] = rng.uniform(40, 150, (m, 1))
    xyth[:, :, 1] = rng.uniform(-5, 5, (m, 1))
    lw = np.tile([[2.0, 1.0]], (m, 1))
    for toucher, want in ((None, False), (45, True)):
        if toucher is not None:
            xyth[toucher, :] = (19.0, 2.0, 0.0)     # corner_touch
        eng.set_obstacles(xyth, lw, np.ones((m, 2), np.uint8), final_time_step=1)
        assert bool(eng.eval_end_states(ego, grid.table(), prm)["flags"][0] & _shim.FLAG_COLLISION) is want
        out = eng.plan_grid(ego[None], grid, prm, want_records=False, want_volume=True)
        assert bool(out["flags"][0, 0] & _shim.FLAG_COLLISION) is want
