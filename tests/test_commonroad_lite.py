"""Host side of SURVEY 8(f-2): the minimal CommonRoad reader, the route -> centre line assembly, the
Cartesian -> Frenet start state and the closed-loop driver's control flow (CPU).

The driver goldens (tests/golden/driver_*.npz) were produced by the reference's OWN planning.py and
global_planner.py run unmodified (tests/golden/make_golden_driver.py); the CPU leg drives our
``frenet_optimal_planning`` with the oracle's planners in place of the CUDA ones (the GPU leg,
tests/test_gpu_driver.py, uses the real ones)."""
import os
import types

import numpy as np
import pytest

from conftest import load_golden
from driver_fixtures import METHODS, SCENARIOS, driver_golden, unpack_scenario

from fiss_plus_planner_b200.planners.commonroad_interface import commonroad_lite as crl
from fiss_plus_planner_b200.planners.commonroad_interface.global_planner import GlobalPlanner, RoutePlanner


def _strip(n, x0, y0, x1, y1, width, npts=6):
    """Straight lanelet from (x0, y0) to (x1, y1)."""
    t = np.linspace(0.0, 1.0, npts)[:, None]
    c = np.array([x0, y0]) + t * (np.array([x1, y1]) - np.array([x0, y0]))
    d = np.array([x1 - x0, y1 - y0], dtype=float)
    nrm = np.array([-d[1], d[0]]) / np.hypot(*d) * width / 2
    return c + nrm, c - nrm


def _toy_scenario():
    """1 -> 2 -> 3 on the right lane, 11 -> 12 -> 13 on the left lane (same direction), obstacle on lane 2."""
    lanelets = []
    for k in range(3):
        l, r = _strip(k, 30.0 * k, 0.0, 30.0 * (k + 1), 0.0, 3.5)
        lanelets.append(crl.Lanelet(k + 1, l, r, [k] if k else [], [k + 2] if k < 2 else [], adj_left=11 + k,
                                    adj_left_same_direction=True))
        l, r = _strip(k, 30.0 * k, 3.5, 30.0 * (k + 1), 3.5, 3.5)
        lanelets.append(crl.Lanelet(11 + k, l, r, [10 + k] if k else [], [12 + k] if k < 2 else [], adj_right=k + 1,
                                    adj_right_same_direction=True))
    net = crl.LaneletNetwork(lanelets)
    shape = crl.Rectangle(4.5, 1.8)
    states = [crl.CustomState(position=np.array([40.0 + 0.5 * t, 0.1]), orientation=0.01 * t, time_step=t, velocity=5.0,
                              acceleration=0.0) for t in range(1, 21)]
    ob = crl.DynamicObstacle(42, "car", shape, crl.CustomState(position=np.array([40.0, 0.1]), orientation=0.0, time_step=0,
                                                              velocity=5.0, acceleration=0.0),
                             crl.TrajectoryPrediction(crl.Trajectory(1, states), shape))
    sc = crl.Scenario(0.1, "ZAM_Toy-1_1_T-1", net, [], [ob])
    goal = crl.GoalRegion([crl.CustomState(time_step=crl.Interval(10, 30), position=[net.find_lanelet_by_id(13)])], {0: [13]}, net)
    pp = crl.PlanningProblem(7, crl.CustomState(position=np.array([3.0, 0.2]), orientation=0.05, time_step=0, velocity=8.0,
                                                acceleration=0.0, yaw_rate=0.0, slip_angle=0.0), goal)
    return sc, pp


def test_xml_round_trip_and_obstacle_semantics(tmp_path):
    sc, pp = _toy_scenario()
    path = os.path.join(str(tmp_path), "toy.xml")
    crl.write_commonroad_xml(path, sc, pp)
    sc2, pps = crl.CommonRoadFileReader(path).open()
    pp2 = list(pps.planning_problem_dict.values())[0]
    assert sc2.dt == 0.1 and sc2.scenario_id == "ZAM_Toy-1_1_T-1"
    for a, b in zip(sc.lanelet_network.lanelets, sc2.lanelet_network.lanelets):
        np.testing.assert_array_equal(a.left_vertices, b.left_vertices)
        np.testing.assert_array_equal(a.center_vertices, b.center_vertices)
        assert (a.successor, a.predecessor, a.adj_left, a.adj_right) == (b.successor, b.predecessor, b.adj_left, b.adj_right)
    ob = sc2.dynamic_obstacles[0]
    assert ob.prediction.final_time_step == 20 and ob.obstacle_shape.length == 4.5
    np.testing.assert_array_equal(ob.state_at_time(0).position, [40.0, 0.1])     # the initial state is step 0
    np.testing.assert_array_equal(ob.state_at_time(7).position, [43.5, 0.1])
    assert ob.state_at_time(21) is None and ob.state_at_time(-1) is None           # outside the prediction: None
    xyth, valid = ob.dense_table(25)
    assert valid.tolist() == [1] * 21 + [0] * 4 and xyth[20, 2] == 0.2
    assert pp2.initial_state.acceleration == 0.0 and pp2.initial_state.velocity == 8.0
    assert pp2.goal.lanelets_of_goal_position == {0: [13]}
    assert not pp2.goal.state_list[0].has_value("velocity")


def test_route_search_and_centerline_assembly():
    sc, pp = _toy_scenario()
    rp = RoutePlanner(sc, pp).plan_routes()
    assert rp.ids_start == [1] and rp.ids_goal == [13]
    # searched backwards from the goal: the lane change happens as late as possible
    assert rp.retrieve_best_route_by_orientation().list_ids_lanelets == [1, 2, 3, 13]
    plan = GlobalPlanner().plan_global_route(sc, pp)
    cl = plan.concat_centerline
    assert cl.shape[1] == 4
    # shared end points of consecutive lanelets appear once, in route order
    assert len(np.unique(cl[:, :2], axis=0)) == len(cl)
    np.testing.assert_allclose(cl[:16, 0], np.linspace(0, 90, 16))
    np.testing.assert_allclose(cl[:, 3], 3.5)
    np.testing.assert_allclose(cl[:15, 2], 0.0, atol=1e-15)
    assert cl[-1, 2] == cl[-2, 2]


def test_goal_region_is_reached():
    sc, pp = _toy_scenario()
    inside = crl.CustomState(time_step=12, position=np.array([75.0, 3.4]), orientation=0.0, velocity=3.0)
    assert pp.goal.is_reached(inside)
    assert not pp.goal.is_reached(crl.CustomState(time_step=9, position=np.array([75.0, 3.4])))     # too early
    assert not pp.goal.is_reached(crl.CustomState(time_step=12, position=np.array([75.0, 0.0])))    # other lane
    assert pp.goal.is_reached(crl.CustomState(time_step=30, position=np.array([60.0, 3.5])))        # on the boundary


@pytest.mark.parametrize("name", SCENARIOS)
def test_reference_centerline_and_start_state(name, tmp_path):
    if not os.path.exists(driver_golden("FISS", name)):
        pytest.skip("no golden")
    _centerline_and_start_state(name, tmp_path)


def _centerline_and_start_state(name, tmp_path):
    """Reader + route stand-in + assembly reproduce the centre line the reference's global_planner.py built, and
    FrenetState.from_state reproduces the first plan() input of the reference's driver."""
    from fiss_plus_planner_b200.planners.common.geometry.cubic_spline import CubicSpline2D
    from fiss_plus_planner_b200.planners.common.scenario.frenet import FrenetState, State
    g = load_golden(driver_golden("FISS", name))
    sc, pps = crl.CommonRoadFileReader(unpack_scenario(name, tmp_path)).open()
    pp = list(pps.planning_problem_dict.values())[0]
    cl = GlobalPlanner().plan_global_route(sc, pp).concat_centerline
    np.testing.assert_array_equal(cl, g["centerline"])
    sp = CubicSpline2D(cl[:, 0], cl[:, 1])
    s = np.arange(0, sp.s[-1], 0.1)
    ref = np.column_stack(([sp.calc_position(v) for v in s], [sp.calc_yaw(v) for v in s], [sp.calc_curvature(v) for v in s]))
    init = pp.initial_state
    fs = FrenetState()
    fs.from_state(State(t=0.0, x=init.position[0], y=init.position[1], yaw=init.orientation, v=init.velocity,
                        a=init.acceleration), ref)
    np.testing.assert_allclose(fs.as_ego6(), g["cycle_ego"][0], rtol=1e-12, atol=1e-12)


class _OraclePlanner(object):
    """The oracle's planners behind the product planner interface (test double for the CPU leg)."""

    def __init__(self, method):
        self.method = method

    def __call__(self, settings, vehicle, scenario=None, **kw):
        from oracle import fop_oracle as fo
        cls = {"FOP": fo.FopOracle, "FOP+": fo.FopPlusOracle, "FISS": fo.FissOracle, "FISS+": fo.FissPlusOracle}[self.method]
        st = fo.Settings(settings.num_width, settings.num_speed, settings.num_t)
        self.o = cls(st, vehicle.l, vehicle.w, vehicle.max_speed, vehicle.max_accel)
        self.all_trajs = []
        self.stats = None
        self._obs_key = None
        self.settings = settings
        return self

    def generate_frenet_frame(self, pts):
        from fiss_plus_planner_b200.planners.common.geometry.cubic_spline import CubicSpline2D
        self.o.generate_frenet_frame(pts)
        sp = CubicSpline2D(pts[:, 0], pts[:, 1])
        s = np.arange(0, sp.s[-1], 0.1)
        return sp, np.column_stack(([sp.calc_position(v) for v in s], [sp.calc_yaw(v) for v in s],
                                    [sp.calc_curvature(v) for v in s]))

    def plan(self, fs, max_speed, obstacles, now=0):
        from oracle import fop_oracle as fo
        from fiss_plus_planner_b200.planners.common.scenario.frenet import FrenetState, State
        from fiss_plus_planner_b200.planners.frenet_optimal_planner import Stats, marshal_obstacles
        if self._obs_key != id(obstacles):
            t = marshal_obstacles(obstacles)          # the product's own marshalling of CommonRoad obstacles
            self._obs = fo.ObstacleTable(t.xyth, t.lw, t.valid.astype(bool), t.final_time_step)
            self._obs_key = id(obstacles)
        tr = self.o.plan((fs.s, fs.s_d, fs.s_dd, fs.d, fs.d_d, fs.d_dd), max_speed, self._obs, now)
        self.stats = Stats()
        (self.stats.num_iter, self.stats.num_trajs_generated, self.stats.num_trajs_validated,
         self.stats.num_collison_checks) = self.o.stats.as_tuple()
        if tr is None:
            return None
        out = types.SimpleNamespace(tr=tr, idx=tr.idx, cost_final=tr.cost_final)
        out.state_at_time_step = lambda k: State(tr.t[k], tr.x[k], tr.y[k], tr.yaw[k], tr.s_d[k], tr.s_dd[k])
        out.frenet_state_at_time_step = lambda k: FrenetState(tr.t[k], tr.s[k], tr.s_d[k], tr.s_dd[k], tr.s_ddd[k],
                                                              tr.d[k], tr.d_d[k], tr.d_dd[k], tr.d_ddd[k])
        return out


@pytest.mark.parametrize("name", [n for n in SCENARIOS if "Flensburg" in n])   # (the Lohmar scenes run on the GPU leg)
@pytest.mark.parametrize("method", ["FISS", "FISS+"])       # the cheap searches; FOP/FOP+ run on the GPU leg
def test_driver_control_flow_vs_reference_golden(method, name, tmp_path, monkeypatch):
    from fiss_plus_planner_b200.planners.benchmark import planning
    from fiss_plus_planner_b200.planners.commonroad_interface.vehicle_parameters import VehicleParameterMapping
    g = load_golden(driver_golden(method, name))
    sc, pps = crl.CommonRoadFileReader(unpack_scenario(name, tmp_path)).open()
    pp = list(pps.planning_problem_dict.values())[0]
    monkeypatch.setitem(planning._PLANNERS, method, (_OraclePlanner(method), planning._PLANNERS[method][1]))
    log = []

    def hook(pl):
        pl.o.settings.time_limit = 1e9
        orig = pl.plan

        def spy(fs, v, obs, now=0):
            best = orig(fs, v, obs, now)
            log.append((fs.as_ego6(), None if best is None else best.cost_final, pl.o.stats.as_tuple()))
            return best
        pl.plan = spy
    reached, traj, avg_t, times, stats, _ = planning.frenet_optimal_planning(
        sc, pp, VehicleParameterMapping["VW_VANAGON"].value, method, tuple(int(v) for v in g["num_samples"]),
        verbose=False, planner_hook=hook)
    assert reached == bool(g["goal_reached"]) and len(log) == int(g["cycles"]) == len(times)
    np.testing.assert_allclose([l[0] for l in log], g["cycle_ego"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose([l[1] for l in log], g["cycle_cost"], rtol=1e-9)
    np.testing.assert_array_equal([l[2] for l in log], g["cycle_stats"])
    got = np.array([[s.time_step, s.position[0], s.position[1], s.orientation, s.velocity, s.velocity_y] for s in traj.state_list])
    np.testing.assert_allclose(got, g["states"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose([stats.num_iter, stats.num_trajs_generated, stats.num_trajs_validated,
                                stats.num_collison_checks], g["avg_stats"], rtol=1e-12)
