"""Scene ownership (B200): several planners on ONE engine, obstacle lists edited in place, lazy candidate bundles that
outlive their reference line -- every plan() must see ITS tables (the reference re-reads the obstacle objects and its own
spline every cycle, frenet_optimal_planner.py:185-189,110-119).  Also: entry points leave the caller's CUDA device alone."""
import numpy as np
import pytest

from conftest import golden_files, load_golden

pytestmark = pytest.mark.gpu


def _mk(g, engine=None, line_shift=0.0):
    from fiss_plus_planner_b200 import synthetic as syn
    from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import FrenetOptimalPlanner, FrenetOptimalPlannerSettings
    st = FrenetOptimalPlannerSettings(*[int(v) for v in g["num_samples"]])
    st.min_t, st.max_t = float(g["min_t"]), float(g["max_t"])
    veh = Vehicle(syn.vehicle_params(l=float(g["ego_l"]), w=float(g["ego_w"]), v_max=float(g["max_speed"]), a_max=float(g["max_accel"])))
    pl = FrenetOptimalPlanner(st, veh, engine=engine)
    line = g["centerline"].copy()
    line[:, 1] += line_shift
    pl.generate_frenet_frame(line)
    return pl


def _obstacle_objects(g, keep=None):
    """commonroad_lite DynamicObstacle objects with the golden's predictions."""
    from fiss_plus_planner_b200.planners.commonroad_interface.commonroad_lite import (CustomState, DynamicObstacle, Rectangle,
                                                                                       Trajectory, TrajectoryPrediction)
    out = []
    for j in range(len(g["obs_lw"])) if keep is None else keep:
        valid = np.flatnonzero(g["obs_valid"][j])
        last = valid[-1]
        states = [CustomState(position=g["obs_xyth"][j, t, :2].copy(), orientation=float(g["obs_xyth"][j, t, 2]), time_step=int(t))
                  for t in range(1, last + 1)]
        init = CustomState(position=g["obs_xyth"][j, 0, :2].copy(), orientation=float(g["obs_xyth"][j, 0, 2]), time_step=0)
        shape = Rectangle(float(g["obs_lw"][j, 0]), float(g["obs_lw"][j, 1]))
        out.append(DynamicObstacle(j, "car", shape, init, TrajectoryPrediction(Trajectory(1, states), shape)))
    return out


def _state(e):
    from fiss_plus_planner_b200.planners.common.scenario.frenet import FrenetState
    return FrenetState(0.0, e[0], e[1], e[2], 0.0, e[3], e[4], e[5], 0.0)


def test_two_planners_one_engine_and_in_place_obstacle_edits():
    from fiss_plus_planner_b200.engine import FissEngine
    g = load_golden([p for p in golden_files("dense_") if "cfg1_blocked" in p][0])
    speed, ego = float(g["max_target_speed"]), g["ego"]
    everything = [j for j in range(len(g["obs_lw"])) if g["obs_valid"][j].all()]
    assert 0 in everything                                  # obstacle 0 is the blocker of this scene
    # references: private engines
    blocked = _mk(g).plan(_state(ego), speed, _obstacle_objects(g, everything), 0)
    free = _mk(g).plan(_state(ego), speed, _obstacle_objects(g, everything[1:]), 0)
    shifted = _mk(g, line_shift=1.5).plan(_state(ego), speed, _obstacle_objects(g, everything[1:]), 0)
    assert blocked.lattice_index != free.lattice_index
    # one engine, two planners with different reference lines and obstacle lists, interleaved
    eng = FissEngine(0)
    a, b = _mk(g, engine=eng), _mk(g, engine=eng, line_shift=1.5)
    obs_a, obs_b = _obstacle_objects(g, everything), _obstacle_objects(g, everything[1:])
    for _ in range(3):
        ta = a.plan(_state(ego), speed, obs_a, 0)
        tb = b.plan(_state(ego), speed, obs_b, 0)
        assert ta.lattice_index == blocked.lattice_index and ta.cost_final == blocked.cost_final
        assert tb.lattice_index == shifted.lattice_index and tb.cost_final == shifted.cost_final
        np.testing.assert_array_equal(tb.x, shifted.x)
    # the same list object edited in place: element removed, then replaced by an equal-length list
    launches = eng.launch_count
    del obs_a[0]                                            # the blocker leaves
    ta = a.plan(_state(ego), speed, obs_a, 0)
    assert ta.lattice_index == free.lattice_index
    obs_a[:] = _obstacle_objects(g, everything)[:len(obs_a)]      # same length, different elements: the blocker is back
    ta = a.plan(_state(ego), speed, obs_a, 0)
    assert ta.lattice_index == blocked.lattice_index
    # an unchanged list is NOT marshalled again (no prep-kernel launch: lattice + record kernel only)
    before = eng.launch_count
    a.plan(_state(ego), speed, obs_a, 0)
    assert eng.launch_count - before == 2
    # a prediction object mutated in place is the one case that needs telling
    obs_a[0].initial_state.position = obs_a[0].initial_state.position + 500.0
    for s in obs_a[0].prediction.trajectory.state_list:
        s.position = s.position + 500.0
    assert a.plan(_state(ego), speed, obs_a, 0).lattice_index == blocked.lattice_index     # stale by design ...
    a.invalidate_obstacles()
    assert a.plan(_state(ego), speed, obs_a, 0).lattice_index == free.lattice_index        # ... until invalidated
    assert eng.launch_count > launches


def test_candidate_bundle_keeps_its_reference_line():
    g = load_golden([p for p in golden_files("dense_") if "cfg2_m8.npz" in p][0])
    speed, ego = float(g["max_target_speed"]), g["ego"]
    pl = _mk(g)
    pl.plan(_state(ego), speed, [], 0)
    bundle = pl.all_trajs[-1]
    pl.generate_frenet_frame(g["centerline"] + np.array([0.0, 2.0]))        # the planner moves to another road
    pl.plan(_state(ego), speed, [], 0)
    x_late = np.array(bundle[7].x)                                          # materialised only now
    ref = _mk(g)
    ref.plan(_state(ego), speed, [], 0)
    np.testing.assert_array_equal(x_late, ref.all_trajs[-1][7].x)           # computed against the bundle's OWN line
    # ... and the planner gets its own line back on its next cycle
    again = pl.plan(_state(ego), speed, [], 0)
    moved = _mk(g, line_shift=2.0).plan(_state(ego), speed, [], 0)
    np.testing.assert_array_equal(again.y, moved.y)
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import CandidateBundle
    assert list(CandidateBundle(pl.engine, ego, np.zeros((0, 4)), None, [], [])) == []


def test_entry_points_restore_the_callers_device():
    import torch
    from fiss_plus_planner_b200.engine import FissEngine
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    g = load_golden([p for p in golden_files("dense_") if "cfg2_m8.npz" in p][0])
    torch.cuda.set_device(0)
    eng = FissEngine(1)
    pl = _mk(g, engine=eng)
    pl.plan(_state(g["ego"]), float(g["max_target_speed"]), [], 0)
    assert torch.cuda.current_device() == 0
    x = torch.zeros(4, device="cuda")
    assert x.device.index == 0
    eng.close()
