"""frenet_optimal_planning() closed loop on the B200 engine vs the reference's OWN driver
(tests/golden/driver_*.npz: planning.py + global_planner.py + the four planners run unmodified on
data/demo scenarios -- BASELINE config #1 and a scene where collisions decide).  Reader -> route -> spline
frame -> Cartesian->Frenet -> one CUDA plan() per cycle; compared cycle by cycle."""
import numpy as np
import pytest

from conftest import load_golden
from driver_fixtures import METHODS, SCENARIOS, driver_golden, unpack_scenario

pytestmark = pytest.mark.gpu


import os  # noqa: E402

CASES = [(m, n) for n in SCENARIOS for m in METHODS if os.path.exists(driver_golden(m, n))]


@pytest.mark.parametrize("method,name", CASES, ids=["%s-%s" % c for c in CASES])
def test_closed_loop_driver_vs_reference(method, name, tmp_path):
    from fiss_plus_planner_b200.planners.benchmark import planning
    from fiss_plus_planner_b200.planners.commonroad_interface.commonroad_lite import CommonRoadFileReader
    from fiss_plus_planner_b200.planners.commonroad_interface.vehicle_parameters import VehicleParameterMapping
    g = load_golden(driver_golden(method, name))
    sc, pps = CommonRoadFileReader(unpack_scenario(name, tmp_path)).open()
    pp = list(pps.planning_problem_dict.values())[0]
    log = []

    def hook(pl):
        if method == "FISS+":
            pl.settings.time_limit = 1e9      # as in the golden run: refinement never cut by wall-clock
        orig = pl.plan

        def spy(fs, v, obs, now=0):
            best = orig(fs, v, obs, now)
            st = pl.stats
            log.append(dict(ego=fs.as_ego6(), cost=np.nan if best is None else best.cost_final,
                            idx=[-9] * 3 if best is None else list(np.asarray(best.idx)),
                            n=-1 if best is None else len(best.t), n_cart=-1 if best is None else len(best.x),
                            stats=(st.num_iter, st.num_trajs_generated, st.num_trajs_validated, st.num_collison_checks)))
            return best
        pl.plan = spy
    reached, traj, avg_t, times, stats, all_trajs = planning.frenet_optimal_planning(
        sc, pp, VehicleParameterMapping["VW_VANAGON"].value, method, tuple(int(v) for v in g["num_samples"]),
        verbose=False, planner_hook=hook)
    assert reached == bool(g["goal_reached"])
    assert len(log) == int(g["cycles"]) == len(times)
    np.testing.assert_allclose([l["ego"] for l in log], g["cycle_ego"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose([l["cost"] for l in log], g["cycle_cost"], rtol=1e-8)
    np.testing.assert_array_equal([l["idx"] for l in log], g["cycle_idx"])
    np.testing.assert_array_equal([l["stats"] for l in log], g["cycle_stats"])
    np.testing.assert_array_equal([l["n"] for l in log], g["cycle_n"])
    np.testing.assert_array_equal([l["n_cart"] for l in log], g["cycle_n_cart"])
    got = np.array([[s.time_step, s.position[0], s.position[1], s.orientation, s.velocity, s.velocity_y]
                    for s in traj.state_list])
    np.testing.assert_allclose(got, g["states"], rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose([stats.num_iter, stats.num_trajs_generated, stats.num_trajs_validated,
                                stats.num_collison_checks], g["avg_stats"], rtol=1e-12)
    assert len(all_trajs) >= 1


def test_planning_entry_point(tmp_path):
    """planning(cfg, output_dir, input_dir, file): the scripts/demo_cr.py entry, cfg keys as cfgs/demo_config.yaml."""
    import os
    from fiss_plus_planner_b200.planners.benchmark import planning
    name = SCENARIOS[0]
    path = unpack_scenario(name, tmp_path)
    cfg = {"PLANNER": "FISS+", "N_W_SAMPLE": 5, "N_S_SAMPLE": 5, "N_T_SAMPLE": 5, "SAVE_GIF": False, "SAVE_TRAJECTORY": True}
    res = planning.planning(cfg, str(tmp_path), os.path.dirname(path), os.path.basename(path))
    assert res is not None and res[0] is True
    assert os.path.exists(os.path.join(str(tmp_path), name + "_FISS+.csv"))
