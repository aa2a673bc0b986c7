"""BASELINE.json's full sizes through size-independent properties (the oracle needs ~0.1 s per candidate, so at
B = 4096 x 270 or 33 x 17 x 9 x 100 steps it can only spot-check): batch invariance, permutation invariance, the
argmin rule recomputed from the volume, mirror symmetry of the lattice, agreement of the two independent kernels
(lattice vs list), materialised rows vs the winners' records, and oracle spot checks."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(scene_name, batch, lattice=None, a_max=None):
    from fiss_plus_planner_b200 import synthetic as syn
    from fiss_plus_planner_b200.engine import FissEngine, fop_grid, make_params
    from fiss_plus_planner_b200.planners.common.cost.cost_function import CostFunction
    from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import FrenetOptimalPlannerSettings
    sc = syn.make_scene(scene_name, batch=batch)
    veh = Vehicle(syn.vehicle_params(**({} if a_max is None else {"a_max": a_max})))
    st = FrenetOptimalPlannerSettings(*(lattice or sc.num_samples))
    st.min_t, st.max_t, st.highest_speed = sc.min_t, sc.max_t, sc.max_target_speed
    eng = FissEngine(0)
    eng.set_spline(sc.spline.device_table())
    eng.set_obstacles(sc.obs.xyth, sc.obs.lw, sc.obs.valid, sc.obs.final_time_step)
    grid = fop_grid(st, veh.w)
    prm = make_params(st, veh, CostFunction("WX1").as_device_weights())
    return sc, veh, st, eng, grid, prm


def _argmin_rule(cost, flags):
    """frenet_optimal_planner.py:263-268: `min_cost >= cost` scan over the survivors => LAST minimum wins."""
    from fiss_plus_planner_b200 import _shim
    feas = (flags & _shim.FLAG_INFEASIBLE_MASK) == 0
    best = np.full(len(cost), -1, np.int64)
    for b in range(len(cost)):
        if feas[b].any():
            c = np.where(feas[b], cost[b], np.inf)
            best[b] = np.flatnonzero(c == c.min()).max()
    return best


def _oracle(sc, veh, st):
    from oracle import fop_oracle as fo
    ost = fo.Settings(st.num_width, st.num_speed, st.num_t)
    ost.min_t, ost.max_t, ost.highest_speed = sc.min_t, sc.max_t, sc.max_target_speed
    opl = fo.FopOracle(ost, veh.l, veh.w, veh.max_speed, veh.max_accel)
    sp = opl.generate_frenet_frame(sc.centerline)
    obs = fo.ObstacleTable(sc.obs.xyth, sc.obs.lw, sc.obs.valid, sc.obs.final_time_step)
    return fo, opl, sp, obs


def test_config4_batch_4096():
    """4096 ego states x 9x6x5 x <=50 steps x 32 obstacles (BASELINE configs[3]) on one GPU."""
    from fiss_plus_planner_b200.engine import decode_flags
    sc, veh, st, eng, grid, prm = _setup("cfg4_batch4096_32obs", 4096, a_max=2.0)   # a_max 2: the accel mask is mixed
    out = eng.plan_grid(sc.ego, grid, prm, want_records=True, want_volume=True)
    B, C = out["cost"].shape
    assert (B, C) == (4096, 270)
    assert np.isfinite(out["cost"]).all()
    # the winner is the reference's argmin rule applied to the volume, its cost is the volume's entry
    best = _argmin_rule(out["cost"], out["flags"])
    np.testing.assert_array_equal(out["best_idx"], best)
    has = best >= 0
    np.testing.assert_array_equal(out["best_cost"][has], out["cost"][np.arange(B)[has], best[has]])
    ok, coll, n_cart = decode_flags(out["flags"])
    assert 0.05 < coll.mean() < 0.95 and 0.05 < ok.mean() < 0.999 and has.mean() > 0.5      # all masks exercised
    # batch invariance: a problem planned alone gives bit-identical results
    rng = np.random.default_rng(5)
    for b in rng.choice(B, 12, replace=False):
        one = eng.plan_grid(sc.ego[b:b + 1], grid, prm, want_records=True, want_volume=True)
        np.testing.assert_array_equal(one["cost"][0], out["cost"][b])
        np.testing.assert_array_equal(one["flags"][0], out["flags"][b])
        assert one["best_idx"][0] == out["best_idx"][b]
        np.testing.assert_array_equal(one["records"][0], out["records"][b])
    # permutation invariance
    perm = rng.permutation(B)
    outp = eng.plan_grid(np.ascontiguousarray(sc.ego[perm]), grid, prm, want_records=False, want_volume=True)
    np.testing.assert_array_equal(outp["cost"], out["cost"][perm])
    np.testing.assert_array_equal(outp["flags"], out["flags"][perm])
    np.testing.assert_array_equal(outp["best_idx"], out["best_idx"][perm])
    # oracle spot check: 3 whole problems
    fo, opl, sp, obs = _oracle(sc, veh, st)
    for b in rng.choice(B, 3, replace=False):
        ref = fo.dense_lattice_eval(tuple(sc.ego[b]), opl.lattice(), sp, obs, tick=0.1, target_speed=sc.max_target_speed,
                                    max_speed=veh.max_speed, max_accel=veh.max_accel, ego_l=veh.l, ego_w=veh.w)
        np.testing.assert_allclose(out["cost"][b], ref["cost"], rtol=1e-9)
        np.testing.assert_array_equal(ok[b], ref["constraint_ok"])
        np.testing.assert_array_equal(n_cart[b], ref["n_cart"])
        # plan() only collision-checks the survivors of the constraint mask (frenet_optimal_planner.py:257-259)
        np.testing.assert_array_equal(coll[b][ok[b]], np.asarray(ref["collision"])[ok[b]])
        assert int(out["best_idx"][b]) == ref["best"]


def test_config4_materialisation_matches_records_and_list_kernel():
    """Full-materialisation mode at bench size (512 x 270): every winner's five rows equal its record (list kernel),
    padding is NaN, and the lattice kernel's flags equal the list kernel's on the whole batch."""
    import torch
    from fiss_plus_planner_b200.engine import decode_flags
    sc, veh, st, eng, grid, prm = _setup("cfg4_batch4096_32obs", 512)
    dev = torch.device("cuda", 0)
    end = grid.table()
    B, C, ns = 512, len(end), grid.n_stride
    f64 = torch.float64
    ego_t = torch.tensor(sc.ego, dtype=f64, device=dev)
    end_t = torch.tensor(end, dtype=f64, device=dev)
    cost_t = torch.empty(B * C, dtype=f64, device=dev)
    flags_t = torch.empty(B * C, dtype=torch.int32, device=dev)
    mat_t = torch.full((5, B * C, ns), 123.0, dtype=f64, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    eng.eval_grid_dev(ego_t, grid, prm, cost_t, flags_t, mat_t, ns, stream=s)
    cost2_t = torch.empty_like(cost_t)
    flags2_t = torch.empty_like(flags_t)
    eng.eval_candidates_dev(ego_t, end_t, prm, cost2_t, flags2_t, None, ns, stream=s)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(flags_t.cpu().numpy(), flags2_t.cpu().numpy())
    np.testing.assert_allclose(cost_t.cpu().numpy(), cost2_t.cpu().numpy(), rtol=1e-12)
    out = eng.plan_grid(sc.ego, grid, prm, want_records=True, want_volume=False)
    mat = mat_t.cpu().numpy().reshape(5, B, C, ns)
    _, _, n_cart = decode_flags(flags_t.cpu().numpy().astype(np.uint32).reshape(B, C))
    n = end[:, 3].astype(int)
    for b in range(0, B, 7):
        c = int(out["best_idx"][b])
        if c < 0:
            continue
        rec = out["records"][b]
        for mrow, rrow, tol in ((0, 9, 1e-9), (1, 10, 1e-9), (2, 11, 1e-9), (3, 2, 1e-12), (4, 13, 1e-7)):
            got, want = mat[mrow, b, c], rec[rrow]
            np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
            np.testing.assert_allclose(got, want, rtol=1e-4 if mrow == 4 else 1e-9, atol=tol, equal_nan=True)
        assert np.isnan(mat[0, b, c, n_cart[b, c]:]).all() and np.isnan(mat[3, b, c, n[c]:]).all()
    assert not (mat == 123.0).any()          # every element of every row was written


def test_config5_fine_lattice():
    """33 x 17 x 9 lattice x <=100 steps x 32 obstacles (BASELINE configs[4]), a few ego states."""
    import torch
    from fiss_plus_planner_b200.engine import decode_flags
    sc, veh, st, eng, grid, prm = _setup("cfg5_fine_lattice", 6)
    out = eng.plan_grid(sc.ego, grid, prm, want_records=True, want_volume=True)
    B, C = out["cost"].shape
    assert C == 33 * 17 * 9 == 5049 and grid.n_stride == 100
    np.testing.assert_array_equal(out["best_idx"], _argmin_rule(out["cost"], out["flags"]))
    # lattice kernel vs list kernel over all 30 294 candidates
    dev = torch.device("cuda", 0)
    end = grid.table()
    f64 = torch.float64
    ego_t = torch.tensor(sc.ego, dtype=f64, device=dev)
    end_t = torch.tensor(end, dtype=f64, device=dev)
    cost2_t = torch.empty(B * C, dtype=f64, device=dev)
    flags2_t = torch.empty(B * C, dtype=torch.int32, device=dev)
    eng.eval_candidates_dev(ego_t, end_t, prm, cost2_t, flags2_t, None, 100, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out["flags"].ravel(), flags2_t.cpu().numpy().astype(np.uint32))
    np.testing.assert_allclose(out["cost"].ravel(), cost2_t.cpu().numpy(), rtol=1e-12)
    # oracle spot check on 40 random candidates of problem 0 (100 steps x 32 obstacles each)
    fo, opl, sp, obs = _oracle(sc, veh, st)
    lat = opl.lattice()
    ok, coll, n_cart = decode_flags(out["flags"][0])
    rng = np.random.default_rng(11)
    pick = rng.choice(C, 40, replace=False)
    ref = fo.dense_lattice_eval(tuple(sc.ego[0]), [lat[c] for c in pick], sp, obs, tick=0.1, target_speed=sc.max_target_speed,
                                max_speed=veh.max_speed, max_accel=veh.max_accel, ego_l=veh.l, ego_w=veh.w)
    np.testing.assert_allclose(out["cost"][0][pick], ref["cost"], rtol=1e-9)
    np.testing.assert_array_equal(ok[pick], ref["constraint_ok"])
    np.testing.assert_array_equal(n_cart[pick], ref["n_cart"])
    sel = np.asarray(ref["constraint_ok"])
    np.testing.assert_array_equal(coll[pick][sel], np.asarray(ref["collision"])[sel])


def test_mirror_symmetry_ties():
    """d0 = d_d0 = d_dd0 = 0 and no obstacles: the +d and -d halves of the lattice cost exactly the same, and the
    winner is the LARGER id of the tied pair (the reference's `>=` scan)."""
    sc, veh, st, eng, grid, prm = _setup("cfg2_single_ego_8obs", 64)
    eng.set_obstacles(None, np.zeros((0, 2)), None, 0)
    ego = sc.ego.copy()
    ego[:, 3:] = 0.0
    out = eng.plan_grid(ego, grid, prm, want_records=False, want_volume=True)
    nd, nv, nt = 9, 6, 5
    cost = out["cost"].reshape(64, nd, nt, nv)            # (d outer, T, v inner)
    # np.linspace(-a, a, 9) is exactly symmetric only for some pairs (the reference sees 162 distinct costs of 270)
    exact = [i for i in range(nd // 2) if grid.d[i] == -grid.d[nd - 1 - i]]
    assert 0 in exact and len(exact) >= 2
    for i in exact:
        np.testing.assert_array_equal(cost[:, i], cost[:, nd - 1 - i])
    np.testing.assert_allclose(cost, cost[:, ::-1], rtol=1e-13)
    best = out["best_idx"]
    i_d = best // (nt * nv)
    flat = out["cost"]
    for b in range(64):
        mirror = (nd - 1 - i_d[b]) * nt * nv + best[b] % (nt * nv)
        if flat[b, mirror] == flat[b, best[b]]:
            assert best[b] >= mirror          # exact tie: the later candidate wins
