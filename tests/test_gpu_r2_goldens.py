"""Round-2 parity at the headline sizes (B200, through the C-ABI) against goldens produced by EXECUTING the reference
(tests/golden/make_golden_r2.py): BASELINE config 4's per-problem shape (9x6x5, n = 40..50, 32 obstacles), a
337-candidate slice of config 5 (33x17x9, n = 80..100: costs for all 5049 candidates), and the optional curvature mask
(north_star's third mask; the reference carries the check commented out, frenet_optimal_planner.py:145-146) -- both
kernels."""
import os

import numpy as np
import pytest

from conftest import golden_files, load_golden

pytestmark = pytest.mark.gpu

RTOL_TIGHT = 1e-9
ATOL_YAW = 1e-9
ATOL_KAPPA = 1e-7
R2 = golden_files("r2_dense_") + golden_files("r2_waymo_")


def _setup(g, check_curvature):
    from fiss_plus_planner_b200 import synthetic as syn
    from fiss_plus_planner_b200.engine import FissEngine, fop_grid, make_params
    from fiss_plus_planner_b200.planners.common.cost.cost_function import CostFunction
    from fiss_plus_planner_b200.planners.common.geometry.cubic_spline import CubicSpline2D
    from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import FrenetOptimalPlannerSettings
    veh = Vehicle(syn.vehicle_params(l=float(g["ego_l"]), w=float(g["ego_w"]), v_max=float(g["max_speed"]),
                                     a_max=float(g["max_accel"])))
    st = FrenetOptimalPlannerSettings(*[int(v) for v in g["num_samples"]])
    st.min_t, st.max_t, st.highest_speed = float(g["min_t"]), float(g["max_t"]), float(g["max_target_speed"])
    eng = FissEngine(0)
    eng.set_spline(CubicSpline2D(g["centerline"][:, 0], g["centerline"][:, 1]).device_table())
    if "ref_xyth" in g:
        eng.set_obstacles(g["ref_xyth"], g["ref_lw"], g["ref_valid"], int(g["ref_final_time_step"]))
    else:
        eng.set_obstacles(g["obs_xyth"], g["obs_lw"], g["obs_valid"], int(g["final_time_step"]))
    grid = fop_grid(st, veh.w)
    prm = make_params(st, veh, CostFunction("WX1").as_device_weights(), time_step_now=int(g["time_step_now"]),
                      check_curvature=check_curvature, collide_all=True)
    prm.max_curvature = float(g["max_curvature"])      # Vehicle.max_curvature of the reference (vehicle.py:44)
    return eng, grid, prm


@pytest.mark.parametrize("path", R2, ids=[os.path.basename(p)[:-4] for p in R2])
@pytest.mark.parametrize("kernel", ["lattice", "list"])
def test_masks_costs_winner_vs_reference(path, kernel):
    from fiss_plus_planner_b200 import _shim
    from fiss_plus_planner_b200.engine import decode_flags
    g = load_golden(path)
    sel = g["sel"].astype(int)
    for check_curvature in (False, True):
        eng, grid, prm = _setup(g, check_curvature)
        np.testing.assert_array_equal(grid.table()[:, 3].astype(int), g["n"])
        if kernel == "lattice":
            out = eng.plan_grid(g["ego"][None], grid, prm, want_records=False, want_volume=True)
        else:
            out = eng.plan_lattice(g["ego"][None], grid.table(), prm, want_records=False, want_volume=True)
        flags, cost = out["flags"][0], out["cost"][0]
        np.testing.assert_allclose(cost, g["cost"], rtol=RTOL_TIGHT)                 # all candidates (5049 in config 5)
        _, coll, n_cart = decode_flags(flags)
        np.testing.assert_array_equal(n_cart[sel], g["n_cart"])
        speed_accel_ok = (flags & (_shim.FLAG_SPEED | _shim.FLAG_ACCEL)) == 0
        np.testing.assert_array_equal(speed_accel_ok[sel], g["constraint_ok"])       # bit-exact masks
        np.testing.assert_array_equal(coll[sel], g["collision"])
        curv_bit = (flags & _shim.FLAG_CURVATURE) != 0
        if check_curvature:
            np.testing.assert_array_equal(~curv_bit[sel], g["curvature_ok"])
            feasible = g["constraint_ok"] & g["curvature_ok"] & ~g["collision"]
        else:
            assert not curv_bit.any()                                                # reference behaviour: bit never set
            feasible = g["constraint_ok"] & ~g["collision"]
        np.testing.assert_array_equal(((flags & _shim.FLAG_INFEASIBLE_MASK) == 0)[sel], feasible)
        # the argmin rule (:263-268) over the slice, from the device's own cost / flags ...
        best, lo = -1, np.inf
        for q, i in enumerate(sel):
            if (flags[i] & _shim.FLAG_INFEASIBLE_MASK) == 0 and lo >= cost[i]:
                lo, best = cost[i], int(i)
        if not check_curvature:
            assert best == int(g["best_in_slice"])
        # ... and the device's own pick over the WHOLE lattice obeys the same rule
        feas_all = (flags & _shim.FLAG_INFEASIBLE_MASK) == 0
        if feas_all.any():
            want = int(np.flatnonzero(feas_all & (cost == cost[feas_all].min()))[-1])
            assert int(out["best_idx"][0]) == want


@pytest.mark.parametrize("path", R2[:3], ids=[os.path.basename(p)[:-4] for p in R2[:3]])
def test_trajectories_vs_reference(path):
    g = load_golden(path)
    eng, grid, prm = _setup(g, False)
    sel, keep = g["sel"].astype(int), g["keep"].astype(int)
    ids = sel[keep]
    end = grid.table()[ids]
    rec = eng.eval_end_states(g["ego"], end, prm, want_records=True)["records"]
    ns = grid.n_stride
    # the lattice kernel's materialised rows of the same candidates
    import torch
    dev = torch.device("cuda", 0)
    C = grid.num_candidates
    ego_t = torch.tensor(g["ego"][None], dtype=torch.float64, device=dev)
    cost_t = torch.empty(C, dtype=torch.float64, device=dev)
    flags_t = torch.empty(C, dtype=torch.int32, device=dev)
    mat_t = torch.empty((5, C, ns), dtype=torch.float64, device=dev)
    eng.eval_grid_dev(ego_t, grid, prm, cost_t, flags_t, mat_t, ns, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    mat = mat_t.cpu().numpy()
    for row, c in enumerate(ids):
        for f, r_row, m_row, tol in (("x", 9, 0, dict(rtol=RTOL_TIGHT)), ("y", 10, 1, dict(rtol=RTOL_TIGHT)),
                                     ("yaw", 11, 2, dict(rtol=0, atol=ATOL_YAW)), ("s_d", 2, 3, dict(rtol=RTOL_TIGHT)),
                                     ("c", 13, 4, dict(rtol=1e-4, atol=ATOL_KAPPA))):
            want = g["traj_" + f][row]
            want = want[~np.isnan(want)]
            np.testing.assert_allclose(rec[row, r_row, :len(want)], want, err_msg=f"record {f} cand {c}", **tol)
            np.testing.assert_allclose(mat[m_row, c, :len(want)], want, err_msg=f"materialised {f} cand {c}", **tol)
            assert np.isnan(mat[m_row, c, len(want):]).all()


def test_heading_wraps_like_arctan2_on_a_westbound_road():
    """The lattice kernel forms a candidate's heading as (reference-line heading) + atan(lateral / longitudinal progress)
    and wraps the sum into (-pi, pi].  On a road that runs along -x the headings straddle +-pi: yaw must jump exactly where
    numpy's arctan2 does, and kappa -- differenced WITHOUT unwrapping by the reference (frenet_optimal_planner.py:132) --
    must show the same 2 pi / ds spikes.  Oracle comparison on every candidate of a 5x4x3 lattice, both kernels."""
    import torch
    from fiss_plus_planner_b200 import synthetic as syn
    from fiss_plus_planner_b200.engine import FissEngine, fop_grid, make_params
    from fiss_plus_planner_b200.planners.common.cost.cost_function import CostFunction
    from fiss_plus_planner_b200.planners.common.geometry.cubic_spline import CubicSpline2D
    from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import FrenetOptimalPlannerSettings
    from oracle import fop_oracle as fo
    k = np.arange(61)
    line = np.column_stack((-5.0 * k + 465.7, 3.0 * np.sin(5.0 * k / 30.0) - 304.8))      # heading ~ pi, wiggling across it
    veh = Vehicle(syn.vehicle_params())
    st = FrenetOptimalPlannerSettings(5, 4, 3)
    st.min_t, st.max_t, st.highest_speed = 4.0, 5.0, 13.4112
    eng = FissEngine(0)
    eng.set_spline(CubicSpline2D(line[:, 0], line[:, 1]).device_table())
    eng.set_obstacles(None, np.zeros((0, 2)), None, 0)
    grid = fop_grid(st, veh.w)
    prm = make_params(st, veh, CostFunction("WX1").as_device_weights())
    egos = np.array([[20.0, 9.0, 0.2, 0.3, -0.2, 0.1], [95.0, 4.0, -0.3, -0.5, 0.25, 0.0], [150.0, 12.0, 0.0, 0.0, 0.0, 0.0]])
    dev = torch.device("cuda", 0)
    b, c, ns = len(egos), grid.num_candidates, grid.n_stride
    ego_t = torch.tensor(egos, dtype=torch.float64, device=dev)
    cost_t = torch.empty(b * c, dtype=torch.float64, device=dev)
    flags_t = torch.empty(b * c, dtype=torch.int32, device=dev)
    mat_t = torch.empty((5, b * c, ns), dtype=torch.float64, device=dev)
    eng.eval_grid_dev(ego_t, grid, prm, cost_t, flags_t, mat_t, ns, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    mat = mat_t.cpu().numpy().reshape(5, b, c, ns)
    ost = fo.Settings(5, 4, 3)
    ost.min_t, ost.max_t, ost.highest_speed = 4.0, 5.0, 13.4112
    opl = fo.FopOracle(ost, veh.l, veh.w, veh.max_speed, veh.max_accel)
    sp = opl.generate_frenet_frame(line)
    wraps = 0
    for bi, ego in enumerate(egos):
        rec = eng.eval_end_states(ego, grid.table(), prm, want_records=True)["records"]          # list kernel
        for ci, (d, v, T) in enumerate(opl.lattice()):
            tr = fo.to_global(fo.generate(fo.Traj(), tuple(ego), d, v, T, 0.1, 13.4112), sp, 0.1)
            n1 = len(tr.x)
            assert n1 >= 2
            for got_yaw, got_c in ((mat[2, bi, ci], mat[4, bi, ci]), (rec[ci, 11], rec[ci, 13])):
                np.testing.assert_allclose(got_yaw[:n1], tr.yaw, rtol=0, atol=ATOL_YAW)
                np.testing.assert_allclose(got_c[:n1 - 1], tr.c, rtol=1e-4, atol=ATOL_KAPPA)
            wraps += int(np.count_nonzero(np.abs(np.diff(tr.yaw)) > 6.0))
    assert wraps > 20, "the scene must actually cross +-pi"
