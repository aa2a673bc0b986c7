"""The N > 1 paths on real GPUs over NCCL (skipped on a one-GPU box; the host logic is covered on CPU by
test_multi_rank_gloo.py): problems sharded across ranks with no data-path collective, and one fine lattice split
across ranks with the all-reduce pick -- both against a single-GPU run of the same problems."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(rank, scene_name, batch):
    from fiss_plus_planner_b200 import synthetic as syn
    from fiss_plus_planner_b200.engine import FissEngine, fop_grid, make_params
    from fiss_plus_planner_b200.planners.common.cost.cost_function import CostFunction
    from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import FrenetOptimalPlannerSettings
    sc = syn.make_scene(scene_name, batch=batch)
    veh = Vehicle(syn.vehicle_params())
    st = FrenetOptimalPlannerSettings(*sc.num_samples)
    st.min_t, st.max_t, st.highest_speed = sc.min_t, sc.max_t, sc.max_target_speed
    eng = FissEngine(rank)
    eng.set_spline(sc.spline.device_table())
    eng.set_obstacles(sc.obs.xyth, sc.obs.lw, sc.obs.valid, sc.obs.final_time_step)
    return sc, eng, fop_grid(st, veh.w), make_params(st, veh, CostFunction("WX1").as_device_weights())


def _sym_case(rank):
    from fiss_plus_planner_b200.engine import LatticeGrid
    sc, eng, grid, prm = _setup(rank, "cfg2_single_ego_8obs", 2)
    eng.set_obstacles(None, np.zeros((0, 2)), None, 0)
    ego = sc.ego.copy()
    ego[:, 3:] = 0.0
    d = np.concatenate((-np.arange(8, 0, -1) * 0.125, np.arange(1, 9) * 0.125))       # 16 rows, exactly mirrored, no 0
    return ego, eng, LatticeGrid(d, grid.v, grid.T, grid.tick, "dtv"), prm


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from fiss_plus_planner_b200.batch import ShardedBatchPlanner, SplitLatticePlanner
        # config 4 style: 64 problems sharded, winners published with one all-reduce pair
        sc, eng, grid, prm = _setup(rank, "cfg4_batch4096_32obs", 64)
        sb = ShardedBatchPlanner(eng, grid, prm)
        loc = sb.plan_local(sc.ego)
        gi, gc = sb.gather_winners(loc, len(sc.ego))
        # config 5 style: one fine lattice split by lateral rows, all-reduce pick
        # (fiss_allreduce_pick: NCCL called from libfissgpu.so on the handle's own communicator).  Ego 1 is mirror
        # symmetric: its +-d candidates tie EXACTLY across the slabs of different ranks -- the last one must win
        sc5, eng5, grid5, prm5 = _setup(rank, "cfg5_fine_lattice", 3)
        sp = SplitLatticePlanner(eng5, grid5, prm5, axis="t")
        out5 = sp.plan(sc5.ego)
        assert eng5.comm_world == world and sp._use_device_path()
        out5b = sp.plan(sc5.ego[::-1].copy())        # buffers / graph re-used with other ego states
        for k in out5:
            np.testing.assert_array_equal(np.asarray(out5b[k])[::-1], np.asarray(out5[k]), err_msg=k)
        out5d = SplitLatticePlanner(eng5, grid5, prm5, axis="d").plan(sc5.ego)       # slabs of lateral rows
        # exact ties ACROSS ranks: an even number of lateral rows, a mirror-symmetric ego state and no obstacles make
        # the two innermost rows tie bit for bit -- they sit on different ranks, the +d one (larger id) must win
        sym = _sym_case(rank)
        out_sym = SplitLatticePlanner(sym[1], sym[2], sym[3], axis="d").plan(sym[0])
        q.put((rank, loc["problems"], gi, gc, {k: np.asarray(v) for k, v in out5.items()},
               {k: np.asarray(v) for k, v in out5d.items()}, {k: np.asarray(v) for k, v in out_sym.items()}))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_nccl_sharded_batch_and_split_lattice():
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sc, eng, grid, prm = _setup(0, "cfg4_batch4096_32obs", 64)
    ref = eng.plan_grid(sc.ego, grid, prm, want_records=False)
    sc5, eng5, grid5, prm5 = _setup(0, "cfg5_fine_lattice", 3)
    ref5 = eng5.plan_grid(sc5.ego, grid5, prm5, want_records=True)
    ego_s, eng_s, grid_s, prm_s = _sym_case(0)
    ref_s = eng_s.plan_grid(ego_s, grid_s, prm_s, want_records=True, want_volume=True)
    for b in range(len(ego_s)):   # the minimum is attained twice, by mirrored rows (so the cross-rank tie rule is exercised)
        feas = (ref_s["flags"][b] & 15) == 0
        ties = np.flatnonzero(feas & (ref_s["cost"][b] == ref_s["best_cost"][b]))
        assert len(ties) == 2 and ref_s["best_idx"][b] == ties.max()
        assert grid_s.table()[ties[0], 0] == -grid_s.table()[ties[1], 0] > 0 or grid_s.table()[ties[1], 0] > 0
    covered = []
    for rank, problems, gi, gc, out5, out5d, out_sym in res:
        for k in ("best_idx", "best_cost", "records", "meta"):
            np.testing.assert_array_equal(out5d[k], ref5[k], err_msg="lateral split: " + k)
            np.testing.assert_array_equal(out_sym[k], ref_s[k], err_msg="cross-rank tie: " + k)
        covered.append(problems)
        np.testing.assert_array_equal(gi, ref["best_idx"])
        np.testing.assert_array_equal(gc[gi >= 0], ref["best_cost"][gi >= 0])
        np.testing.assert_array_equal(out5["best_idx"], ref5["best_idx"])
        np.testing.assert_array_equal(out5["best_cost"], ref5["best_cost"])
        np.testing.assert_array_equal(out5["records"], ref5["records"])
        np.testing.assert_array_equal(out5["meta"], ref5["meta"])
    assert covered[0][0] == 0 and covered[-1][1] == 64 and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))


def test_allreduce_pick_single_rank_is_identity_plus_offset():
    """nranks = 1 (no NCCL needed, runs on the one-GPU box): the pack / select / unpack kernels alone must hand back the
    local winners with the id offset applied, bit-identical records, and NaN records where nothing is feasible."""
    import torch
    sc, eng, grid, prm = _setup(0, "cfg2_single_ego_8obs", 5)
    dev = torch.device("cuda", 0)
    ns = grid.n_stride
    ref = eng.plan_grid(sc.ego, grid, prm, want_records=True)
    idx = torch.tensor(ref["best_idx"], dtype=torch.int32, device=dev)
    cost = torch.tensor(ref["best_cost"], dtype=torch.float64, device=dev)
    meta = torch.tensor(ref["meta"], dtype=torch.int32, device=dev)
    rec = torch.tensor(ref["records"], dtype=torch.float64, device=dev)
    idx[3] = -1                                    # "nothing feasible" for problem 3
    cost[3] = float("inf")
    eng.allreduce_pick_dev(idx, cost, meta, rec, ns, id_offset=1000)
    torch.cuda.synchronize()
    want_idx = np.where(ref["best_idx"] >= 0, ref["best_idx"].astype(np.int64) + 1000, -1)
    want_idx[3] = -1
    np.testing.assert_array_equal(idx.cpu().numpy(), want_idx)
    keep = (np.arange(5) != 3) & (ref["best_idx"] >= 0)
    assert keep.sum() >= 3
    np.testing.assert_array_equal(cost.cpu().numpy()[keep], ref["best_cost"][keep])
    assert np.isinf(cost.cpu().numpy()[3])
    np.testing.assert_array_equal(rec.cpu().numpy()[keep], ref["records"][keep])      # NaN padding included, bit for bit
    assert np.isnan(rec.cpu().numpy()[3]).all()
    np.testing.assert_array_equal(meta.cpu().numpy()[keep], ref["meta"][keep])
    assert (meta.cpu().numpy()[3] == 0).all()
