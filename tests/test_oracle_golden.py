"""Pin oracle/fop_oracle.py to the reference: compare with tests/golden/*.npz, which were produced
by executing the reference's own modules (tests/golden/make_golden.py).  CPU only."""
import os
import warnings

import numpy as np
import pytest

from conftest import golden_files, load_golden
from oracle import fop_oracle as fo

warnings.filterwarnings("ignore")

DENSE = golden_files("dense_")
LOOPS = golden_files("loop_")


def _oracle_planner(g, cls):
    st = fo.Settings(*[int(v) for v in g["num_samples"]])
    st.min_t, st.max_t = float(g["min_t"]), float(g["max_t"])
    pl = cls(st, float(g["ego_l"]), float(g["ego_w"]), float(g["max_speed"]), float(g["max_accel"]))
    pl.generate_frenet_frame(g["centerline"])
    obs = fo.ObstacleTable(g["obs_xyth"], g["obs_lw"], g["obs_valid"], int(g["final_time_step"]))
    return pl, obs


def test_golden_files_present():
    assert len(DENSE) >= 8 and len(LOOPS) >= 8


@pytest.mark.parametrize("path", DENSE, ids=[os.path.basename(p)[:-4] for p in DENSE])
def test_dense_lattice_matches_reference(path):
    g = load_golden(path)
    pl, obs = _oracle_planner(g, fo.FopOracle)
    pl.settings.highest_speed = float(g["max_target_speed"])
    r = fo.dense_lattice_eval(tuple(g["ego"]), pl.lattice(), pl.spline, obs, tick=0.1,
                              target_speed=float(g["max_target_speed"]), max_speed=float(g["max_speed"]),
                              max_accel=float(g["max_accel"]), ego_l=float(g["ego_l"]), ego_w=float(g["ego_w"]),
                              now=int(g["time_step_now"]))
    # the oracle keeps the reference's operation order: FP64 results are bit-identical
    np.testing.assert_array_equal(r["cost"], g["cost"])
    np.testing.assert_array_equal(r["n"], g["n"])
    np.testing.assert_array_equal(r["n_cart"], g["n_cart"])
    np.testing.assert_array_equal(r["constraint_ok"], g["constraint_ok"])
    np.testing.assert_array_equal(r["collision"], g["collision"])
    assert r["best"] == int(g["best"])
    np.testing.assert_array_equal(np.array(pl.spline.s, dtype=np.float64), g["knots"])
    for row, seq in enumerate(g["keep"]):
        tr = r["trajs"][int(seq)]
        for f in ("x", "y", "yaw", "s_d", "c", "s", "d"):
            want = g["traj_" + f][row]
            want = want[~np.isnan(want)]
            np.testing.assert_array_equal(np.asarray(getattr(tr, f), dtype=np.float64), want, err_msg=f"{f} cand {seq}")


@pytest.mark.parametrize("path", LOOPS, ids=[os.path.basename(p)[:-4] for p in LOOPS])
def test_closed_loop_matches_reference(path):
    g = load_golden(path)
    tag = os.path.basename(path).split("_")[1]
    method = tag.replace("plus", "+")
    pl, obs = _oracle_planner(g, fo.PLANNERS[method])
    ego = tuple(g["ego"][0])
    for i in range(len(g["cost"])):
        np.testing.assert_array_equal(np.array(ego, dtype=np.float64), g["ego"][i])
        best = pl.plan(ego, float(g["max_target_speed"]), obs, i)
        assert best is not None
        assert best.cost_final == g["cost"][i]
        np.testing.assert_array_equal(np.asarray(best.idx), g["idx"][i])
        assert pl.stats.as_tuple() == tuple(int(v) for v in g["stats"][i])
        assert len(best.t) == g["n"][i] and len(best.x) == g["n_cart"][i]
        for f in ("t", "s", "s_d", "s_dd", "s_ddd", "d", "d_d", "d_dd", "d_ddd", "x", "y", "yaw", "ds", "c", "c_d", "c_dd"):
            want = g["best_" + f][i]
            want = want[~np.isnan(want)]
            np.testing.assert_array_equal(np.asarray(getattr(best, f), dtype=np.float64), want, err_msg=f)
        # frenet_state_at_time_step(1) -- planning.py:137-138
        ego = (best.s[1], best.s_d[1], best.s_dd[1], best.d[1], best.d_d[1], best.d_dd[1])
