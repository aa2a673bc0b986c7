"""The launch plumbing around the kernels (B200): a plan step as ONE CUDA-graph launch (``fiss_plan_grid_dev``, and the
small-batch path of ``fiss_plan_grid_host``) and the streaming submit / wait pair -- results must be bit-identical to the
plain launch-by-launch path whatever changes between calls (time_step_now, ego states, batch size, lattice, scene)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene(name="cfg4_batch4096_32obs", batch=64, lattice=(9, 6, 5), t=(4.0, 5.0), m=32):
    from fiss_plus_planner_b200 import synthetic as syn
    from fiss_plus_planner_b200.engine import FissEngine, fop_grid, make_params
    from fiss_plus_planner_b200.planners.common.cost.cost_function import CostFunction
    from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import FrenetOptimalPlannerSettings
    sc = syn.make_scene(name, batch=batch, num_obstacles=m)
    veh = Vehicle(syn.vehicle_params())
    st = FrenetOptimalPlannerSettings(*lattice)
    st.min_t, st.max_t, st.highest_speed = t[0], t[1], sc.max_target_speed
    eng = FissEngine(0)
    eng.set_spline(sc.spline.device_table())
    eng.set_obstacles(sc.obs.xyth, sc.obs.lw, sc.obs.valid, sc.obs.final_time_step)
    grid = fop_grid(st, veh.w)
    weights = CostFunction("WX1").as_device_weights()
    mk = lambda now=0, **kw: make_params(st, veh, weights, time_step_now=now, **kw)  # noqa: E731
    return sc, eng, grid, mk


def _dev_buffers(b, grid, want_mat):
    import torch
    dev = torch.device("cuda", 0)
    c, ns = grid.num_candidates, grid.n_stride
    f64, i32 = torch.float64, torch.int32
    return dict(cost=torch.empty(b * c, dtype=f64, device=dev), flags=torch.empty(b * c, dtype=i32, device=dev),
                mat=torch.empty((5, b * c, ns), dtype=f64, device=dev) if want_mat else None,
                idx=torch.empty(b, dtype=i32, device=dev), best=torch.empty(b, dtype=f64, device=dev),
                meta=torch.empty((b, 2), dtype=i32, device=dev), rec=torch.empty((b, 16, ns), dtype=f64, device=dev))


def _host(t):
    return None if t is None else t.cpu().numpy()


@pytest.mark.parametrize("want_mat", [False, True])
def test_plan_grid_dev_graph_equals_separate_launches(want_mat):
    import torch
    sc, eng, grid, mk = _scene(batch=48)
    dev = torch.device("cuda", 0)
    s = torch.cuda.current_stream().cuda_stream
    end_t = torch.tensor(grid.table(), dtype=torch.float64, device=dev)
    a, b = _dev_buffers(48, grid, want_mat), _dev_buffers(48, grid, want_mat)
    launches0 = eng.launch_count
    n_calls = 0
    # the same buffers across calls (graph re-used, nodes patched), ego states and time_step_now changing
    for rep, now in enumerate((0, 0, 7, 7, 30, 95, 0)):
        ego = np.roll(sc.ego[:48], rep, axis=0).copy()
        ego_t = torch.tensor(ego, dtype=torch.float64, device=dev)
        prm = mk(now)
        eng.plan_grid_dev(ego_t, grid, prm, a["cost"], a["flags"], a["mat"], a["idx"], a["best"], a["meta"], a["rec"],
                          grid.n_stride, stream=s)
        eng.eval_grid_dev(ego_t, grid, prm, b["cost"], b["flags"], b["mat"], grid.n_stride, stream=s)
        eng.pick_winners_dev(ego_t, end_t, prm, b["cost"], b["flags"], b["idx"], b["best"], b["rec"], b["meta"],
                             grid.n_stride, stream=s)
        torch.cuda.synchronize()
        n_calls += 1
        for k in a:
            if a[k] is not None:
                np.testing.assert_array_equal(_host(a[k]), _host(b[k]), err_msg=f"{k} rep {rep} now {now}")
    assert eng.launch_count - launches0 == 4 * n_calls          # two kernels per step on either path
    # a different batch size on the same handle: the graph is rebuilt, results stay right
    ego_t = torch.tensor(sc.ego[:20].copy(), dtype=torch.float64, device=dev)
    a2, b2 = _dev_buffers(20, grid, want_mat), _dev_buffers(20, grid, want_mat)
    eng.plan_grid_dev(ego_t, grid, mk(3), a2["cost"], a2["flags"], a2["mat"], a2["idx"], a2["best"], a2["meta"], a2["rec"],
                      grid.n_stride, stream=s)
    eng.eval_grid_dev(ego_t, grid, mk(3), b2["cost"], b2["flags"], b2["mat"], grid.n_stride, stream=s)
    eng.pick_winners_dev(ego_t, end_t, mk(3), b2["cost"], b2["flags"], b2["idx"], b2["best"], b2["rec"], b2["meta"],
                         grid.n_stride, stream=s)
    torch.cuda.synchronize()
    for k in a2:
        if a2[k] is not None:
            np.testing.assert_array_equal(_host(a2[k]), _host(b2[k]), err_msg=k)


def test_host_small_batch_graph_tracks_every_change():
    """fiss_plan_grid_host's one-graph path (small batches) against the list-kernel path, while everything a caller can
    change between cycles changes: time step, ego state, what is asked for, the obstacle table, the lattice."""
    from fiss_plus_planner_b200 import synthetic as syn
    sc, eng, grid, mk = _scene("cfg2_single_ego_8obs", batch=6, m=8)
    _, _, grid2, _ = _scene("cfg2_single_ego_8obs", batch=1, lattice=(5, 4, 3), m=8)
    sc3 = syn.make_scene("cfg3_64obs", batch=1)
    steps = [dict(now=0), dict(now=1), dict(now=2, b=3), dict(now=3, b=3, rec=False), dict(now=4, vol=False),
             dict(now=5, grid=grid2), dict(now=6), dict(now=7, obstacles=sc3.obs), dict(now=40), dict(now=99), dict(now=0)]
    cur = grid
    for k, s in enumerate(steps):
        if "obstacles" in s:
            o = s["obstacles"]
            eng.set_obstacles(o.xyth, o.lw, o.valid, o.final_time_step)
        cur = s.get("grid", cur)
        b = s.get("b", 1)
        ego = sc.ego[k % 4:k % 4 + b].copy()
        prm = mk(s["now"])
        got = eng.plan_grid(ego, cur, prm, want_records=s.get("rec", True), want_volume=s.get("vol", True))
        ref = eng.plan_lattice(ego, cur.table(), prm, want_records=True, want_volume=True)      # generic list kernel
        np.testing.assert_array_equal(got["best_idx"], ref["best_idx"], err_msg=str(s))
        np.testing.assert_allclose(got["best_cost"], ref["best_cost"], rtol=1e-12)   # lattice vs list kernel: rounding
        np.testing.assert_array_equal(got["meta"], ref["meta"])
        if got["flags"] is not None:
            np.testing.assert_array_equal(got["flags"], ref["flags"])
            np.testing.assert_allclose(got["cost"], ref["cost"], rtol=1e-12)
        if got["records"] is not None:
            np.testing.assert_allclose(got["records"], ref["records"], rtol=1e-9, atol=1e-9, equal_nan=True)


def test_submit_wait_streams_batches():
    from fiss_plus_planner_b200.engine import FissError
    sc, eng, grid, mk = _scene(batch=6 * 96)
    prm = mk(0)
    batches = [np.ascontiguousarray(sc.ego[i * 96:(i + 1) * 96]) for i in range(6)]
    want = [eng.plan_grid(b, grid, prm, want_records=True, want_volume=False) for b in batches]
    for pinned in (True, False):
        outs = [eng.alloc_plan_outputs(96, grid, want_records=True, want_volume=False, pinned=pinned) for _ in range(2)]
        got = []
        eng.plan_grid_submit(0, batches[0], grid, prm, outs[0])
        for k in range(1, len(batches) + 1):
            if k < len(batches):
                eng.plan_grid_submit(k % 2, batches[k], grid, prm, outs[k % 2])     # in flight together with k - 1
            eng.plan_grid_wait((k - 1) % 2)
            got.append({f: np.array(v) for f, v in outs[(k - 1) % 2].items() if v is not None and not f.startswith("_")})
        for w, g in zip(want, got):
            for f in ("best_idx", "best_cost", "meta", "records"):
                np.testing.assert_array_equal(g[f], w[f], err_msg=f)
    # lane discipline
    out = eng.alloc_plan_outputs(96, grid, want_records=True, want_volume=False, pinned=False)
    eng.plan_grid_submit(1, batches[0], grid, prm, out)
    with pytest.raises(FissError):
        eng.plan_grid_submit(1, batches[1], grid, prm, out)
    eng.plan_grid_wait(1)
    with pytest.raises(FissError):
        eng.plan_grid_wait(1)
    with pytest.raises(FissError):
        eng.plan_grid_submit(4, batches[0], grid, prm, out)   # lanes 0 .. FISS_LANES - 1 = 3
