"""Round-2 pins of the oracle (CPU): BASELINE-size lattices (config 4's 9x6x5 x 32 obstacles, a 337-candidate slice of
config 5's 33x17x9 x 100 steps), the optional curvature mask (frenet_optimal_planner.py:145-146) and the Waymo wire
format (waymo_interface.py:24-76) against goldens produced by EXECUTING the reference (tests/golden/make_golden_r2.py)."""
import os
import warnings

import numpy as np
import pytest

from conftest import golden_files, load_golden
from oracle import fop_oracle as fo

warnings.filterwarnings("ignore")

R2_DENSE = golden_files("r2_dense_")
R2_WAYMO = golden_files("r2_waymo_")


def _planner(g):
    st = fo.Settings(*[int(v) for v in g["num_samples"]])
    st.min_t, st.max_t, st.highest_speed = float(g["min_t"]), float(g["max_t"]), float(g["max_target_speed"])
    pl = fo.FopOracle(st, float(g["ego_l"]), float(g["ego_w"]), float(g["max_speed"]), float(g["max_accel"]))
    pl.generate_frenet_frame(g["centerline"])
    return pl


def _check_dense(g, pl, obs):
    lattice = pl.lattice()
    sel = g["sel"].astype(int)
    st = pl.settings
    ego = tuple(g["ego"])
    # calc_frenet_paths for EVERY candidate: cost and step count
    every = [fo.generate(fo.Traj(), ego, d, v, T, st.tick_t, st.highest_speed) for d, v, T in lattice]
    np.testing.assert_array_equal(np.array([tr.cost_final for tr in every]), g["cost"])
    np.testing.assert_array_equal(np.array([len(tr.t) for tr in every]), g["n"])
    # conversion + masks on the slice
    r = fo.dense_lattice_eval(ego, [lattice[i] for i in sel], pl.spline, obs, tick=st.tick_t, target_speed=st.highest_speed,
                              max_speed=pl.max_speed, max_accel=pl.max_accel, ego_l=pl.ego_l, ego_w=pl.ego_w,
                              now=int(g["time_step_now"]))
    np.testing.assert_array_equal(r["n_cart"], g["n_cart"])
    np.testing.assert_array_equal(r["constraint_ok"], g["constraint_ok"])
    np.testing.assert_array_equal(r["collision"], g["collision"])
    assert (int(sel[r["best"]]) if r["best"] >= 0 else -1) == int(g["best_in_slice"])
    lim = float(g["max_curvature"])
    curv = np.array([fo.curvature_ok(tr, lim) for tr in r["trajs"]])
    np.testing.assert_array_equal(curv, g["curvature_ok"])
    both = np.array([fo.passes_constraints(tr, pl.max_speed, pl.max_accel, lim) for tr in r["trajs"]])
    np.testing.assert_array_equal(both, g["constraint_ok"] & g["curvature_ok"])
    assert 0 < curv.sum() < len(curv), "the limit must split the lattice"
    # the GPU comparison of this mask is meaningful only if no candidate sits on the limit
    peak = np.array([np.max(np.abs(tr.c)) if len(tr.c) else 0.0 for tr in r["trajs"]])
    assert np.min(np.abs(peak - lim)) > 1e-5
    for row, q in enumerate(g["keep"].astype(int)):
        tr = r["trajs"][q]
        for f in ("x", "y", "yaw", "s_d", "c", "s", "d"):
            want = g["traj_" + f][row]
            np.testing.assert_array_equal(np.asarray(getattr(tr, f), dtype=np.float64), want[~np.isnan(want)], err_msg=f)


def test_golden_files_present():
    assert len(R2_DENSE) == 3 and len(R2_WAYMO) == 2


@pytest.mark.parametrize("path", R2_DENSE, ids=[os.path.basename(p)[:-4] for p in R2_DENSE])
def test_baseline_size_lattices_match_reference(path):
    g = load_golden(path)
    pl = _planner(g)
    obs = fo.ObstacleTable(g["obs_xyth"], g["obs_lw"], g["obs_valid"], int(g["final_time_step"]))
    assert len(g["obs_lw"]) == 32
    _check_dense(g, pl, obs)


@pytest.mark.parametrize("path", R2_WAYMO, ids=[os.path.basename(p)[:-4] for p in R2_WAYMO])
def test_waymo_conversion_matches_reference(path):
    g = load_golden(path)
    tab = fo.waymo_obstacle_table(g["waymo_trajs"], g["waymo_mask"])
    np.testing.assert_array_equal(tab.valid, g["ref_valid"])
    np.testing.assert_array_equal(tab.xyth, g["ref_xyth"])
    np.testing.assert_array_equal(tab.lw, g["ref_lw"])
    assert tab.final_time_step == int(g["ref_final_time_step"]) == 36
    assert 0 not in g["kept_ids"] and 5 not in g["kept_ids"] and len(g["kept_ids"]) == len(g["waymo_mask"]) - 2
    _check_dense(g, _planner(g), tab)


@pytest.mark.parametrize("path", R2_WAYMO, ids=[os.path.basename(p)[:-4] for p in R2_WAYMO])
def test_waymo_goldens_discriminate_wrong_semantics(path):
    """The cases are built so that each rule of :45-58 decides the mask: keeping the dropped agents' step-0 state, or
    taking final_time_step from the tensor length instead of the first KEPT agent, changes the collision mask."""
    g = load_golden(path)
    pl = _planner(g)
    tr, mk = g["waymo_trajs"], g["waymo_mask"]
    n, t, _ = tr.shape
    run_valid = np.ones((n, t), bool)
    for i in range(n):
        for u in range(1, t):
            if not mk[i, u]:
                run_valid[i, u:] = False
                break
    xyth, lw = tr[:, :, [0, 1, 6]].astype(np.float64), tr[:, 0, 3:5].astype(np.float64)

    def collisions(obs):
        lattice = pl.lattice()
        ring = fo.sat.ego_ring(pl.ego_l, pl.ego_w)
        out = []
        for d, v, T in lattice[::3]:
            trj = fo.to_global(fo.generate(fo.Traj(), tuple(g["ego"]), d, v, T, 0.1, pl.settings.highest_speed), pl.spline, 0.1)
            out.append(fo.has_collision(trj, obs, ring, 0))
        return np.array(out)

    right = collisions(fo.waymo_obstacle_table(tr, mk))
    np.testing.assert_array_equal(right, g["collision"][::3])
    step0_kept = collisions(fo.ObstacleTable(xyth, lw, run_valid, 36))          # dropped agents keep their initial state
    assert step0_kept.all() and not right.all()                                  # agent 5 sits on the ego at step 0
    keep = np.flatnonzero(mk[:, 1])
    long_horizon = collisions(fo.ObstacleTable(xyth[keep], lw[keep], run_valid[keep], t - 1))
    assert long_horizon.sum() > right.sum() or "m8" in path


def test_product_host_rules_equal_reference_conversion():
    """The product's own host-side restatement (WaymoObstacles: keep rules, dense table, lazy obstacle objects)."""
    from fiss_plus_planner_b200.planners.waymo_interface.waymo_interface import convert_waymo_obstacle_to_cr
    for path in R2_WAYMO:
        g = load_golden(path)
        for trajs in (g["waymo_trajs"], g["waymo_trajs"].astype(np.float64)):
            obs = convert_waymo_obstacle_to_cr(trajs, g["waymo_mask"])
            assert len(obs) == len(g["kept_ids"])
            xyth, lw, valid, final = obs.dense_table()
            np.testing.assert_array_equal(valid.astype(bool), g["ref_valid"])
            np.testing.assert_array_equal(xyth, g["ref_xyth"])
            np.testing.assert_array_equal(lw, g["ref_lw"])
            assert final == int(g["ref_final_time_step"])
            o = obs[2]
            assert o.obstacle_id == int(g["kept_ids"][2]) and o.prediction.final_time_step == 11
            assert o.state_at_time(12) is None and o.state_at_time(0) is not None
            np.testing.assert_array_equal(o.state_at_time(11).position, g["ref_xyth"][2, 11, :2])
        empty = convert_waymo_obstacle_to_cr(g["waymo_trajs"][:, :1], g["waymo_mask"][:, :1])
        assert len(empty) == 0 and list(empty) == []
