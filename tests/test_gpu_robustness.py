"""Error convention and degenerate inputs through the C ABI (int32 status codes, fiss_last_error, no UB;
NaN / Inf candidates are masked infeasible -- the reference's "exception => collision", frenet_optimal_planner.py:178-182)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(batch=8, obstacles=True):
    from fiss_plus_planner_b200 import synthetic as syn
    from fiss_plus_planner_b200.engine import FissEngine, fop_grid, make_params
    from fiss_plus_planner_b200.planners.common.cost.cost_function import CostFunction
    from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import FrenetOptimalPlannerSettings
    sc = syn.make_scene("cfg2_single_ego_8obs", batch=batch)
    veh = Vehicle(syn.vehicle_params())
    st = FrenetOptimalPlannerSettings(*sc.num_samples)
    st.min_t, st.max_t, st.highest_speed = sc.min_t, sc.max_t, sc.max_target_speed
    eng = FissEngine(0)
    eng.set_spline(sc.spline.device_table())
    if obstacles:
        eng.set_obstacles(sc.obs.xyth, sc.obs.lw, sc.obs.valid, sc.obs.final_time_step)
    return sc, eng, fop_grid(st, veh.w), make_params(st, veh, CostFunction("WX1").as_device_weights())


def test_non_finite_ego_states_are_infeasible_not_ub():
    from fiss_plus_planner_b200 import _shim
    sc, eng, grid, prm = _setup()
    ego = sc.ego.copy()
    ego[1, 0] = np.nan          # s0
    ego[2, 1] = np.inf          # s_d0
    ego[3, 3] = np.nan          # d0
    ego[4, 5] = -np.inf         # d_dd0
    ego[5, 0] = 1.0e9           # far beyond the reference line: no Cartesian point at all (n' = 0)
    out = eng.plan_grid(ego, grid, prm, want_records=True, want_volume=True)
    clean = eng.plan_grid(sc.ego, grid, prm, want_records=True, want_volume=True)
    for b in (2, 3, 4):
        assert out["best_idx"][b] == -1 and out["best_cost"][b] == np.inf
        assert not np.isfinite(out["cost"][b]).any() or ((out["flags"][b] & _shim.FLAG_INFEASIBLE_MASK) != 0).all()
    # s0 = NaN or far beyond the line: s never lands on the spline => n' = 0 everywhere; no obstacle can be hit
    # (has_collision's loop never runs, :173-176) and the Frenet-only cost decides, as in the reference
    for b in (1, 5):
        assert (((out["flags"][b] >> _shim.FLAG_NCART_SHIFT) & _shim.FLAG_NCART_MASK) == 0).all()
        assert ((out["flags"][b] & _shim.FLAG_COLLISION) == 0).all()
        assert out["best_idx"][b] >= 0 and np.isnan(out["records"][b][9]).all()      # a winner without a single x
    # neighbours of the poisoned problems are untouched
    for b in (0, 6, 7):
        np.testing.assert_array_equal(out["cost"][b], clean["cost"][b])
        np.testing.assert_array_equal(out["flags"][b], clean["flags"][b])
        assert out["best_idx"][b] == clean["best_idx"][b]
        np.testing.assert_array_equal(out["records"][b], clean["records"][b])


def test_pinned_and_pageable_buffers_give_identical_results():
    sc, eng, grid, prm = _setup(batch=64)
    a = eng.plan_grid(sc.ego, grid, prm, want_records=True, want_volume=True)
    ego_pin = eng.pinned_empty(sc.ego.shape, np.float64)
    ego_pin[...] = sc.ego
    outs = eng.alloc_plan_outputs(64, grid, want_records=True, want_volume=True, pinned=True)
    for _ in range(2):                                    # buffers are reusable
        b = eng.plan_grid(ego_pin, grid, prm, want_records=True, want_volume=True, out=outs)
        assert b is outs
        for k in ("best_idx", "best_cost", "meta", "records", "cost", "flags"):
            np.testing.assert_array_equal(a[k], b[k])


def test_status_codes_and_messages():
    import ctypes as C
    from fiss_plus_planner_b200 import _shim
    from fiss_plus_planner_b200._shim import FissError
    sc, eng, grid, prm = _setup(batch=2)
    lib = _shim.load()
    assert lib.fiss_create(99, C.byref(C.c_void_p())) == -1                        # FISS_ERR_INVALID: no such device
    assert b"device" in lib.fiss_last_error(None)
    with pytest.raises(FissError, match="tick_t"):
        bad = _shim.FissParams.from_buffer_copy(prm)
        bad.tick_t = 0.0
        eng.plan_grid(sc.ego, grid, bad)
    with pytest.raises(FissError, match="time_step_now"):
        bad = _shim.FissParams.from_buffer_copy(prm)
        bad.time_step_now = -1
        eng.plan_grid(sc.ego, grid, bad)
    end = grid.table()
    with pytest.raises(FissError, match="integral n"):
        e2 = end.copy()
        e2[3, 3] = 41.5
        eng.eval_end_states(sc.ego[0], e2, prm)
    with pytest.raises(FissError, match="ascending"):
        tab = sc.spline.device_table().copy()
        tab[0, 5] = tab[0, 3]
        eng.set_spline(tab)
    with pytest.raises(FissError, match="K >= 2"):
        eng.fit_splines(np.zeros((1, 1, 2)))
    # the handle is still usable after every error
    out = eng.plan_grid(sc.ego, grid, prm)
    assert out["best_idx"].shape == (2,)


def test_no_obstacles_and_many_obstacles():
    """M = 0 (:170-171: no obstacles => collision-free) and M = 200 (several mask words per (row, step))."""
    from fiss_plus_planner_b200 import _shim
    from fiss_plus_planner_b200 import synthetic as syn
    sc, eng, grid, prm = _setup(batch=4, obstacles=False)
    out0 = eng.plan_grid(sc.ego, grid, prm, want_volume=True)
    assert ((out0["flags"] & _shim.FLAG_COLLISION) == 0).all()
    big = syn.make_scene("cfg3_64obs", batch=4, num_obstacles=200)
    eng.set_obstacles(big.obs.xyth, big.obs.lw, big.obs.valid, big.obs.final_time_step)
    out = eng.plan_grid(sc.ego, grid, prm, want_volume=True)
    end = grid.table()
    import torch
    dev = torch.device("cuda", 0)
    cost_t = torch.empty(4 * len(end), dtype=torch.float64, device=dev)
    flags_t = torch.empty(4 * len(end), dtype=torch.int32, device=dev)
    eng.eval_candidates_dev(torch.tensor(sc.ego, dtype=torch.float64, device=dev), torch.tensor(end, dtype=torch.float64, device=dev),
                            prm, cost_t, flags_t, None, grid.n_stride, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out["flags"].ravel(), flags_t.cpu().numpy().astype(np.uint32))   # lattice == list kernel
    assert ((out["flags"] & _shim.FLAG_COLLISION) != 0).any()


def test_round2_entry_points_reject_bad_arguments():
    """fiss_plan_grid_dev / _submit / _wait, fiss_set_obstacles_waymo, fiss_allreduce_pick: status codes, messages, and a
    handle that keeps working afterwards."""
    import ctypes as C
    import torch
    from fiss_plus_planner_b200 import _shim
    from fiss_plus_planner_b200._shim import FissError
    sc, eng, grid, prm = _setup(batch=3)
    lib = _shim.load()
    h, g, pp = eng._h, C.byref(grid.c_struct), C.byref(prm)
    assert lib.fiss_plan_grid_dev(h, None, None, 3, g, pp, None, None, None, None, None, None, None, grid.n_stride) == -1
    assert b"plan_grid_dev" in lib.fiss_last_error(h)
    out = eng.alloc_plan_outputs(3, grid, want_records=True, want_volume=False, pinned=False)
    with pytest.raises(FissError, match="lane"):
        eng.plan_grid_submit(7, sc.ego, grid, prm, out)
    with pytest.raises(FissError, match="nothing in flight"):
        eng.plan_grid_wait(0)
    with pytest.raises(FissError, match="n_stride"):
        short = dict(out)
        short.pop("_ptrs", None)
        lib_rc = lib.fiss_plan_grid_submit(h, None, 0, _shim.ptr(sc.ego), 3, g, pp, _shim.ptr(out["best_idx"]),
                                           _shim.ptr(out["best_cost"]), _shim.ptr(out["meta"]), _shim.ptr(out["records"]), 3)
        eng._check(lib_rc, "fiss_plan_grid_submit")
    eng.plan_grid_submit(0, sc.ego, grid, prm, out)         # the failed submit left the lane free
    eng.plan_grid_wait(0)
    ref = eng.plan_grid(sc.ego, grid, prm, want_records=True)
    np.testing.assert_array_equal(out["best_idx"], ref["best_idx"])
    np.testing.assert_array_equal(out["records"], ref["records"])
    # Waymo tensors: wrong shapes are refused before anything is uploaded
    assert lib.fiss_set_obstacles_waymo(h, None, None, None, 4, 10, -1, None) == -1
    assert lib.fiss_set_obstacles_waymo(h, None, None, None, 0, 0, -1, None) == 0          # N = 0: the table is cleared
    assert not (eng.plan_grid(sc.ego, grid, prm, want_volume=True)["flags"] & 8).any()    # ... no collisions any more
    # cross-GPU pick without a communicator: nranks > 1 needs one; a handle without one is a single rank
    dev = torch.device("cuda", 0)
    idx = torch.zeros(3, dtype=torch.int32, device=dev)
    best = torch.zeros(3, dtype=torch.float64, device=dev)
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    assert lib.fiss_allreduce_pick(h, C.c_void_p(1), 2, 5, None, 3, 1, 0, 0, vp(idx), vp(best), None, None, 0) == -1   # rank >= size
    assert lib.fiss_allreduce_pick(h, None, 0, 0, None, 3, 0, 0, 0, vp(idx), vp(best), None, None, 0) == -1            # id map
    assert b"id map" in lib.fiss_last_error(h)
    assert lib.fiss_allreduce_pick(h, None, 0, 0, None, 3, 1 << 40, 0, 7, vp(idx), vp(best), None, None, 0) == 0
    torch.cuda.synchronize()
    assert idx.cpu().tolist() == [7, 7, 7]
    assert lib.fiss_comm_destroy(h) == 0                    # nothing to destroy: fine
