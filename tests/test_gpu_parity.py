"""Parity tests proper (B200): the CUDA path, called through the C-ABI, against the golden vectors
produced by the reference and against the oracle on the same seeded inputs.

Contract (BASELINE.json north_star): <= 1e-4 relative on trajectory (x, y, yaw, v, kappa) and cost;
feasibility masks and the winner bit-exact.  The tolerances below are written out per quantity;
they are far inside the contract except for kappa, which the reference itself computes as
dyaw/ds with ds down to ~1e-3 m -- hence the small absolute term.
"""
import os

import numpy as np
import pytest

from conftest import golden_files, load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-4                 # the contract
RTOL_TIGHT = 1e-9           # what FP64 on both sides actually delivers for x, y, v, cost
ATOL_YAW = 1e-9
ATOL_KAPPA = 1e-7

DENSE = golden_files("dense_")


def _setup(g, collide_all=True):
    from fiss_plus_planner_b200 import synthetic as syn
    from fiss_plus_planner_b200.engine import FissEngine, fop_lattice, make_params
    from fiss_plus_planner_b200.planners.common.cost.cost_function import CostFunction
    from fiss_plus_planner_b200.planners.common.geometry.cubic_spline import CubicSpline2D
    from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import FrenetOptimalPlannerSettings

    veh = Vehicle(syn.vehicle_params(l=float(g["ego_l"]), w=float(g["ego_w"]), v_max=float(g["max_speed"]),
                                     a_max=float(g["max_accel"])))
    st = FrenetOptimalPlannerSettings(*[int(v) for v in g["num_samples"]])
    st.min_t, st.max_t, st.highest_speed = float(g["min_t"]), float(g["max_t"]), float(g["max_target_speed"])
    eng = FissEngine(0)
    spline = CubicSpline2D(g["centerline"][:, 0], g["centerline"][:, 1])
    eng.set_spline(spline.device_table())
    eng.set_obstacles(g["obs_xyth"], g["obs_lw"], g["obs_valid"], int(g["final_time_step"]))
    end = fop_lattice(st, veh.w)
    prm = make_params(st, veh, CostFunction("WX1").as_device_weights(), time_step_now=int(g["time_step_now"]),
                      collide_all=collide_all)
    return eng, end, prm


@pytest.mark.parametrize("path", DENSE, ids=[os.path.basename(p)[:-4] for p in DENSE])
def test_dense_lattice_vs_reference_golden(path):
    from fiss_plus_planner_b200.engine import decode_flags
    g = load_golden(path)
    eng, end, prm = _setup(g, collide_all=True)
    np.testing.assert_array_equal(end[:, 3].astype(int), g["n"])           # step counts n (np.arange rule)
    out = eng.plan_lattice(g["ego"][None], end, prm, want_records=True, want_volume=True)
    ok, coll, n_cart = decode_flags(out["flags"][0])
    np.testing.assert_allclose(out["cost"][0], g["cost"], rtol=RTOL_TIGHT)
    np.testing.assert_array_equal(n_cart, g["n_cart"])
    np.testing.assert_array_equal(ok, g["constraint_ok"])                 # bit-exact masks
    np.testing.assert_array_equal(coll, g["collision"])
    assert int(out["best_idx"][0]) == int(g["best"])                      # winner incl. the tie rule
    if int(g["best"]) >= 0:
        assert out["best_cost"][0] == out["cost"][0][int(g["best"])]
        assert tuple(out["meta"][0]) == (int(g["n"][int(g["best"])]), int(g["n_cart"][int(g["best"])]))

    # plan() semantics: collision only evaluated on constraint survivors, same winner
    eng2, end2, prm2 = _setup(g, collide_all=False)
    out2 = eng2.plan_lattice(g["ego"][None], end2, prm2, want_records=False, want_volume=True)
    ok2, coll2, _ = decode_flags(out2["flags"][0])
    np.testing.assert_array_equal(ok2 & ~coll2, g["constraint_ok"] & ~g["collision"])
    assert int(out2["best_idx"][0]) == int(g["best"])

    # per-step arrays of the kept candidates (full records)
    keep = g["keep"].astype(int)
    rec = eng.eval_end_states(g["ego"], end[keep], prm, want_records=True)
    np.testing.assert_allclose(rec["cost"], g["cost"][keep], rtol=RTOL_TIGHT)
    rows = {"s": 1, "s_d": 2, "d": 5, "x": 9, "y": 10, "yaw": 11, "c": 13}
    for r, seq in enumerate(keep):
        n, nc = int(g["n"][seq]), int(g["n_cart"][seq])
        for f, row in rows.items():
            want = g["traj_" + f][r]
            want = want[~np.isnan(want)]
            ln = {"s": n, "s_d": n, "d": n, "x": nc, "y": nc, "yaw": nc if nc >= 2 else 0, "c": max(nc - 1, 0) if nc >= 2 else 0}[f]
            assert len(want) == ln, (f, seq, len(want), ln)
            got = rec["records"][r, row, :ln]
            atol = {"yaw": ATOL_YAW, "c": ATOL_KAPPA}.get(f, 0.0)
            np.testing.assert_allclose(got, want, rtol=RTOL, atol=atol, err_msg=f"{f} cand {seq}")
            if f in ("s", "s_d", "d", "x", "y"):
                np.testing.assert_allclose(got, want, rtol=RTOL_TIGHT, atol=1e-12, err_msg=f"tight {f} cand {seq}")
            # beyond the valid length the device writes NaN
            assert np.all(np.isnan(rec["records"][r, row, ln:]))


def test_dense_materialisation_matches_records():
    """The (x, y, yaw, v, kappa) materialisation of the hot kernel equals the full records."""
    import torch
    g = load_golden([p for p in DENSE if p.endswith("dense_cfg2_m8.npz")][0])
    eng, end, prm = _setup(g, collide_all=True)
    dev = torch.device("cuda:0")
    c = len(end)
    n_stride = int(end[:, 3].max())
    ego_t = torch.tensor(g["ego"][None], dtype=torch.float64, device=dev)
    end_t = torch.tensor(end, dtype=torch.float64, device=dev)
    cost_t = torch.empty(c, dtype=torch.float64, device=dev)
    flags_t = torch.empty(c, dtype=torch.int32, device=dev)
    mat_t = torch.empty((5, c, n_stride), dtype=torch.float64, device=dev)
    eng.eval_candidates_dev(ego_t, end_t, prm, cost_t, flags_t, mat_t, n_stride,
                            stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    rec = eng.eval_end_states(g["ego"], end, prm, want_records=True)
    mat = mat_t.cpu().numpy()
    np.testing.assert_array_equal(cost_t.cpu().numpy(), rec["cost"])
    np.testing.assert_array_equal(flags_t.cpu().numpy().astype(np.uint32), rec["flags"])
    for mrow, rrow in ((0, 9), (1, 10), (2, 11), (3, 2), (4, 13)):
        np.testing.assert_array_equal(mat[mrow], rec["records"][:, rrow, :])


def test_oracle_random_batch():
    """A batch of random ego states against the oracle (no golden): cost, masks, winners."""
    from fiss_plus_planner_b200 import synthetic as syn
    from fiss_plus_planner_b200.engine import FissEngine, decode_flags, fop_lattice, make_params
    from fiss_plus_planner_b200.planners.common.cost.cost_function import CostFunction
    from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import FrenetOptimalPlannerSettings
    from oracle import fop_oracle as fo

    sc = syn.make_scene("cfg4_batch4096_32obs", batch=6)
    veh = Vehicle(syn.vehicle_params(a_max=2.0))
    st = FrenetOptimalPlannerSettings(*sc.num_samples)
    st.min_t, st.max_t, st.highest_speed = sc.min_t, sc.max_t, sc.max_target_speed
    eng = FissEngine(0)
    eng.set_spline(sc.spline.device_table())
    eng.set_obstacles(sc.obs.xyth, sc.obs.lw, sc.obs.valid, sc.obs.final_time_step)
    end = fop_lattice(st, veh.w)
    prm = make_params(st, veh, CostFunction("WX1").as_device_weights(), collide_all=True)
    out = eng.plan_lattice(sc.ego, end, prm, want_records=False, want_volume=True)

    ost = fo.Settings(*sc.num_samples)
    ost.min_t, ost.max_t, ost.highest_speed = sc.min_t, sc.max_t, sc.max_target_speed
    opl = fo.FopOracle(ost, veh.l, veh.w, veh.max_speed, veh.max_accel)
    sp = opl.generate_frenet_frame(sc.centerline)
    obs = fo.ObstacleTable(sc.obs.xyth, sc.obs.lw, sc.obs.valid, sc.obs.final_time_step)
    for b in range(len(sc.ego)):
        ref = fo.dense_lattice_eval(tuple(sc.ego[b]), opl.lattice(), sp, obs, tick=0.1,
                                    target_speed=sc.max_target_speed, max_speed=veh.max_speed,
                                    max_accel=veh.max_accel, ego_l=veh.l, ego_w=veh.w)
        ok, coll, n_cart = decode_flags(out["flags"][b])
        np.testing.assert_allclose(out["cost"][b], ref["cost"], rtol=RTOL_TIGHT)
        np.testing.assert_array_equal(ok, ref["constraint_ok"])
        np.testing.assert_array_equal(coll, ref["collision"])
        np.testing.assert_array_equal(n_cart, ref["n_cart"])
        assert int(out["best_idx"][b]) == ref["best"]


# ------------------------------------------------------------------------------------------------
# The lattice ("grid") kernel: same contract, same goldens, plus bit-level agreement with the generic
# one-warp-per-candidate kernel on masks / n' / winners.
def _grid_setup(g, collide_all=True):
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
    spline = CubicSpline2D(g["centerline"][:, 0], g["centerline"][:, 1])
    eng.set_spline(spline.device_table())
    eng.set_obstacles(g["obs_xyth"], g["obs_lw"], g["obs_valid"], int(g["final_time_step"]))
    grid = fop_grid(st, veh.w)
    prm = make_params(st, veh, CostFunction("WX1").as_device_weights(), time_step_now=int(g["time_step_now"]),
                      collide_all=collide_all)
    return eng, grid, prm


@pytest.mark.parametrize("path", DENSE, ids=[os.path.basename(p)[:-4] for p in DENSE])
def test_grid_kernel_vs_reference_golden(path):
    import torch
    from fiss_plus_planner_b200.engine import decode_flags
    g = load_golden(path)
    eng, grid, prm = _grid_setup(g, collide_all=True)
    np.testing.assert_array_equal(grid.table()[:, 3].astype(int), g["n"])
    out = eng.plan_grid(g["ego"][None], grid, prm, want_records=True, want_volume=True)
    ok, coll, n_cart = decode_flags(out["flags"][0])
    np.testing.assert_allclose(out["cost"][0], g["cost"], rtol=RTOL_TIGHT)
    np.testing.assert_array_equal(n_cart, g["n_cart"])
    np.testing.assert_array_equal(ok, g["constraint_ok"])                 # bit-exact masks
    np.testing.assert_array_equal(coll, g["collision"])
    assert int(out["best_idx"][0]) == int(g["best"])                      # winner incl. the tie rule
    if int(g["best"]) >= 0:
        assert tuple(out["meta"][0]) == (int(g["n"][int(g["best"])]), int(g["n_cart"][int(g["best"])]))

    # plan() semantics (collision only on constraint survivors): same feasible set, same winner
    eng2, grid2, prm2 = _grid_setup(g, collide_all=False)
    out2 = eng2.plan_grid(g["ego"][None], grid2, prm2, want_records=False, want_volume=True)
    ok2, coll2, _ = decode_flags(out2["flags"][0])
    np.testing.assert_array_equal(ok2 & ~coll2, g["constraint_ok"] & ~g["collision"])
    assert int(out2["best_idx"][0]) == int(g["best"])

    # the five materialised rows (x, y, yaw, v, kappa) of the kept candidates against the reference arrays
    dev = torch.device("cuda:0")
    c, n_stride = grid.num_candidates, grid.n_stride
    ego_t = torch.tensor(g["ego"][None], dtype=torch.float64, device=dev)
    cost_t = torch.empty(c, dtype=torch.float64, device=dev)
    flags_t = torch.empty(c, dtype=torch.int32, device=dev)
    mat_t = torch.full((5, c, n_stride), 7.0, dtype=torch.float64, device=dev)
    eng.eval_grid_dev(ego_t, grid, prm, cost_t, flags_t, mat_t, n_stride, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    mat = mat_t.cpu().numpy()
    np.testing.assert_array_equal(flags_t.cpu().numpy().astype(np.uint32), out["flags"][0])
    np.testing.assert_array_equal(cost_t.cpu().numpy(), out["cost"][0])
    keep = g["keep"].astype(int)
    rows = {"x": 0, "y": 1, "yaw": 2, "s_d": 3, "c": 4}
    for r, seq in enumerate(keep):
        n, nc = int(g["n"][seq]), int(g["n_cart"][seq])
        for f, row in rows.items():
            want = g["traj_" + f][r]
            want = want[~np.isnan(want)]
            ln = {"s_d": n, "x": nc, "y": nc, "yaw": nc if nc >= 2 else 0, "c": max(nc - 1, 0) if nc >= 2 else 0}[f]
            assert len(want) == ln, (f, seq, len(want), ln)
            got = mat[row, seq, :ln]
            atol = {"yaw": ATOL_YAW, "c": ATOL_KAPPA}.get(f, 0.0)
            np.testing.assert_allclose(got, want, rtol=RTOL, atol=atol, err_msg=f"{f} cand {seq}")
            if f in ("s_d", "x", "y"):
                np.testing.assert_allclose(got, want, rtol=RTOL_TIGHT, atol=1e-12, err_msg=f"tight {f} cand {seq}")
            assert np.all(np.isnan(mat[row, seq, ln:]))                    # NaN beyond each row's length


GRID_CASES = [
    # name,                   lattice,      (min_t, max_t), M,  B,  now, a_max
    ("cfg4_batch4096_32obs", (9, 6, 5), (4.0, 5.0), 32, 7, 0, 11.5),       # B small: lateral axis is chunked
    ("cfg4_batch4096_32obs", (9, 6, 5), (4.0, 5.0), 32, 300, 0, 2.0),      # B large: one item per (b, T)
    ("cfg3_64obs", (9, 6, 5), (4.0, 5.0), 64, 5, 0, 11.5),                 # two mask words
    ("cfg2_single_ego_8obs", (9, 6, 5), (4.0, 5.0), 8, 1, 0, 11.5),        # Mp < 32: several steps per ballot
    ("cfg2_single_ego_8obs", (9, 6, 5), (4.0, 5.0), 5, 3, 0, 11.5),        # M not a power of two
    ("cfg2_single_ego_8obs", (9, 6, 5), (4.0, 5.0), 0, 3, 0, 11.5),        # no obstacles
    ("cfg1_demo_substitute", (5, 5, 5), (8.0, 10.0), 27, 4, 37, 11.5),     # n = 80..100, time_step_now > 0
    ("cfg1_demo_substitute", (5, 5, 5), (8.0, 10.0), 27, 2, 80, 11.5),     # horizon cut by final_time_step
    ("cfg5_fine_lattice", (33, 17, 9), (8.0, 10.0), 32, 2, 0, 11.5),       # 5049 candidates, n <= 100
    ("cfg3_64obs", (4, 3, 2), (4.0, 5.0), 100, 2, 0, 11.5),                # Mp = 128: four mask words
]


@pytest.mark.parametrize("case", GRID_CASES, ids=[f"{c[0]}-{'x'.join(map(str, c[1]))}-M{c[3]}-B{c[4]}-t{c[5]}" for c in GRID_CASES])
def test_grid_matches_generic_kernel(case):
    """Lattice kernel vs the generic list kernel on the same inputs: flags (both masks, n') and winners
    identical, cost to rounding, and the materialised rows to rounding with the same NaN pattern."""
    import torch
    from fiss_plus_planner_b200 import synthetic as syn
    from fiss_plus_planner_b200.engine import FissEngine, fop_grid, make_params
    from fiss_plus_planner_b200.planners.common.cost.cost_function import CostFunction
    from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import FrenetOptimalPlannerSettings

    name, lattice, (min_t, max_t), m, b, now, a_max = case
    sc = syn.make_scene(name, batch=b, num_obstacles=max(m, 1))
    veh = Vehicle(syn.vehicle_params(a_max=a_max))
    st = FrenetOptimalPlannerSettings(*lattice)
    st.min_t, st.max_t, st.highest_speed = min_t, max_t, sc.max_target_speed
    eng = FissEngine(0)
    eng.set_spline(sc.spline.device_table())
    if m > 0:
        eng.set_obstacles(sc.obs.xyth, sc.obs.lw, sc.obs.valid, sc.obs.final_time_step)
    grid = fop_grid(st, veh.w)
    end = grid.table()
    for collide_all in (True, False):
        prm = make_params(st, veh, CostFunction("WX1").as_device_weights(), time_step_now=now, collide_all=collide_all)
        a = eng.plan_grid(sc.ego, grid, prm, want_records=True, want_volume=True)
        r = eng.plan_lattice(sc.ego, end, prm, want_records=True, want_volume=True)
        np.testing.assert_array_equal(a["flags"], r["flags"])
        np.testing.assert_allclose(a["cost"], r["cost"], rtol=1e-13)
        np.testing.assert_array_equal(a["best_idx"], r["best_idx"])
        np.testing.assert_array_equal(a["meta"], r["meta"])
        np.testing.assert_array_equal(np.isnan(a["records"]), np.isnan(r["records"]))

    # materialised rows, in FrenetOptimalPlanner numbering (v fastest: grouped stores) and in FissPlanner
    # numbering (T fastest: one candidate per group)
    from fiss_plus_planner_b200.engine import LatticeGrid
    dev = torch.device("cuda:0")
    sptr = torch.cuda.current_stream().cuda_stream
    ego_t = torch.tensor(sc.ego, dtype=torch.float64, device=dev)
    for order in ("dtv", "dvt"):
        gr = grid if order == "dtv" else LatticeGrid(grid.d, grid.v, grid.T, st.tick_t, order)
        c, n_stride = gr.num_candidates, gr.n_stride
        end_t = torch.tensor(gr.table(), dtype=torch.float64, device=dev)
        outs = []
        for which in ("grid", "generic"):
            cost_t = torch.empty(b * c, dtype=torch.float64, device=dev)
            flags_t = torch.empty(b * c, dtype=torch.int32, device=dev)
            mat_t = torch.full((5, b * c, n_stride), 7.0, dtype=torch.float64, device=dev)
            if which == "grid":
                eng.eval_grid_dev(ego_t, gr, prm, cost_t, flags_t, mat_t, n_stride, stream=sptr)
            else:
                eng.eval_candidates_dev(ego_t, end_t, prm, cost_t, flags_t, mat_t, n_stride, stream=sptr)
            torch.cuda.synchronize()
            outs.append((cost_t.cpu().numpy(), flags_t.cpu().numpy(), mat_t.cpu().numpy()))
        (gc, gf, gm), (rc, rf, rm) = outs
        np.testing.assert_array_equal(gf, rf)
        np.testing.assert_allclose(gc, rc, rtol=1e-13)
        np.testing.assert_array_equal(np.isnan(gm), np.isnan(rm))
        # x, y agree to a few ulp (FMA contraction differs between the two kernels); yaw and kappa are finite
        # differences of those over segments down to ~1e-3 m, hence the absolute terms
        for row, (rtol, atol) in enumerate(((1e-13, 0), (1e-13, 0), (1e-9, ATOL_YAW), (1e-10, 1e-13), (1e-6, ATOL_KAPPA))):
            np.testing.assert_allclose(gm[row], rm[row], rtol=rtol, atol=atol, err_msg=f"mat row {row} order {order}")


def test_grid_rejects_bad_grids():
    from fiss_plus_planner_b200 import synthetic as syn
    from fiss_plus_planner_b200._shim import FissError
    from fiss_plus_planner_b200.engine import FissEngine, LatticeGrid, make_params
    from fiss_plus_planner_b200.planners.common.cost.cost_function import CostFunction
    from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import FrenetOptimalPlannerSettings
    sc = syn.make_scene("cfg2_single_ego_8obs", batch=1)
    st = FrenetOptimalPlannerSettings(3, 3, 3)
    veh = Vehicle(syn.vehicle_params())
    prm = make_params(st, veh, CostFunction("WX1").as_device_weights())
    eng = FissEngine(0)
    grid = LatticeGrid([0.0, 0.5], [1.0, 2.0], [4.0], 0.1)
    with pytest.raises(FissError, match="fiss_set_spline"):
        eng.plan_grid(sc.ego, grid, prm)
    eng.set_spline(sc.spline.device_table())
    eng.plan_grid(sc.ego, grid, prm)
    grid.c_struct.stride_v = 7
    with pytest.raises(FissError, match="dense numbering"):
        eng.plan_grid(sc.ego, grid, prm)
    with pytest.raises(FissError, match="1..64"):
        eng.plan_grid(sc.ego, LatticeGrid(np.linspace(-1, 1, 65), [1.0], [4.0], 0.1), prm)


@pytest.mark.parametrize("slots", [1, 3])
def test_grid_kernel_slots_per_item(slots):
    """The lattice kernel packs several (ego, horizon) pairs into one work item when the batch is large (two by
    default).  FISS_GRID_SLOTS forces the count for every launch (the library reads it once per process, hence the
    subprocess): the golden and the lattice-vs-list-kernel parity tests must hold for one and for three pairs too."""
    import subprocess
    import sys
    env = dict(os.environ, FISS_GRID_SLOTS=str(slots))
    here = os.path.abspath(__file__)
    res = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider",
                          f"{here}::test_grid_matches_generic_kernel", f"{here}::test_grid_kernel_vs_reference_golden",
                          f"{here}::test_dense_materialisation_matches_records"],
                         env=env, capture_output=True, text=True, timeout=600, cwd=os.path.dirname(os.path.dirname(here)))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
