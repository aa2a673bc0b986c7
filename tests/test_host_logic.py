"""CPU-only host logic: spline fit vs the oracle's restatement (bit-identical), obstacle
marshalling from CommonRoad-style objects, FrenetState.from_state, Stats arithmetic."""
import types

import numpy as np

from fiss_plus_planner_b200 import synthetic as syn
from fiss_plus_planner_b200.planners.common.geometry.cubic_spline import CubicSpline2D
from fiss_plus_planner_b200.planners.common.scenario.frenet import FrenetState, FrenetTrajectory, State
from oracle import fop_oracle as fo


def test_spline_table_bit_identical_to_oracle():
    for k in (3, 13, 81):
        line = syn.reference_line(k)
        ours = CubicSpline2D(line[:, 0], line[:, 1])
        ref = fo.Spline2D(line[:, 0], line[:, 1])
        np.testing.assert_array_equal(ours.device_table(), ref.table())
        for s in np.linspace(0.0, ours.s[-1] * 0.999, 57):
            assert ours.calc_position(s) == ref.position(s)
            assert ours.calc_yaw(s) == ref.yaw(s)
        assert ours.calc_position(-0.1) == (None, None) and ours.calc_position(ours.s[-1] + 0.1) == (None, None)


def test_marshal_obstacles_from_duck_typed_objects():
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import marshal_obstacles

    class Ob:
        def __init__(self, states, l, w, final):
            self.states = states
            self.obstacle_shape = types.SimpleNamespace(length=l, width=w)
            self.prediction = types.SimpleNamespace(final_time_step=final)

        def state_at_time(self, t):
            s = self.states.get(t)
            return None if s is None else types.SimpleNamespace(position=np.array(s[:2]), orientation=s[2])

    obs = [Ob({t: (1.0 * t, 2.0, 0.1) for t in range(0, 6)}, 4.5, 1.8, 5),
           Ob({t: (3.0, 1.0 * t, -0.2) for t in range(2, 9)}, 5.0, 2.0, 8)]
    tab = marshal_obstacles(obs)
    assert tab.final_time_step == 5                      # obstacles[0] only (frenet_optimal_planner.py:173)
    assert tab.xyth.shape == (2, 9, 3)
    np.testing.assert_array_equal(tab.valid[0], [1, 1, 1, 1, 1, 1, 0, 0, 0])
    np.testing.assert_array_equal(tab.valid[1], [0, 0, 1, 1, 1, 1, 1, 1, 1])
    np.testing.assert_array_equal(tab.lw, [[4.5, 1.8], [5.0, 2.0]])
    np.testing.assert_array_equal(tab.xyth[1, 4], [3.0, 4.0, -0.2])
    assert len(marshal_obstacles([])) == 0


def test_from_state_roundtrip_on_reference_line():
    line = syn.reference_line(21)
    sp = CubicSpline2D(line[:, 0], line[:, 1])
    s = np.arange(0, sp.s[-1], 0.1)
    poly = np.column_stack(([sp.calc_position(v) for v in s], [sp.calc_yaw(v) for v in s]))
    s_true, d_true = 37.3, 0.6
    px, py = sp.calc_position(s_true)
    yaw = sp.calc_yaw(s_true)
    st = State(t=0.0, x=px - d_true * np.sin(yaw), y=py + d_true * np.cos(yaw), yaw=yaw, v=8.0, a=0.0)
    fs = FrenetState()
    fs.from_state(st, poly)
    assert abs(fs.s - s_true) < 0.15 and abs(abs(fs.d) - d_true) < 1e-2
    assert fs.d < 0                                     # CommonRoad sign convention (frenet.py:84-85)
    assert abs(fs.s_d - 8.0) < 1e-2 and abs(fs.d_d) < 0.1


def test_stats_and_trajectory_helpers():
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import Stats
    a, b = Stats(), Stats()
    a.num_iter, b.num_iter, b.num_trajs_generated = 2, 3, 10
    c = a + b
    assert c is a and a.num_iter == 5 and a.num_trajs_generated == 10
    assert a.average(5).num_trajs_generated == 2.0
    rec = np.full((16, 6), np.nan)
    rec[:9, :5] = np.arange(45).reshape(9, 5)
    rec[9:12, :4] = 1.0
    rec[12:14, :3] = 2.0
    rec[14, :2] = 3.0
    rec[15, :1] = 4.0
    tr = FrenetTrajectory().fill_from_device_record(rec, 5, 4, 1.25)
    assert len(tr.t) == 5 and len(tr.x) == 4 and len(tr.ds) == 3 and len(tr.c_d) == 2 and len(tr.c_dd) == 1
    assert tr.cost_final == 1.25 and tr < FrenetTrajectory().fill_from_device_record(rec, 5, 4, 2.0)
    fwd = tr.forward_t_steps(2)
    assert len(fwd.t) == 3 and len(fwd.x) == 2 and tr.forward_t_steps(9) is None
    st = tr.state_at_time_step(1)
    assert (st.x, st.v) == (1.0, rec[2, 1])
    short = FrenetTrajectory().fill_from_device_record(rec, 5, 1, 0.0)
    assert len(short.x) == 1 and len(short.yaw) == 0


def test_shapes_that_are_not_origin_rectangles_are_refused():
    import types
    import pytest
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import marshal_obstacles
    pred = types.SimpleNamespace(final_time_step=3)
    st = types.SimpleNamespace(position=np.zeros(2), orientation=0.0)
    mk = lambda shape: [types.SimpleNamespace(prediction=pred, obstacle_shape=shape, state_at_time=lambda t: st)]  # noqa: E731
    ok = marshal_obstacles(mk(types.SimpleNamespace(length=4.0, width=2.0)))
    assert ok.lw.tolist() == [[4.0, 2.0]]
    with pytest.raises(NotImplementedError):
        marshal_obstacles(mk(types.SimpleNamespace(radius=1.0)))                                    # a circle
    with pytest.raises(NotImplementedError):
        marshal_obstacles(mk(types.SimpleNamespace(length=4.0, width=2.0, center=np.array([1.0, 0.0]))))
    with pytest.raises(NotImplementedError):
        marshal_obstacles(mk(types.SimpleNamespace(length=4.0, width=2.0, orientation=0.3)))
