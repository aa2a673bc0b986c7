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
        sc5, eng5, grid5, prm5 = _setup(rank, "cfg5_fine_lattice", 3)
        out5 = SplitLatticePlanner(eng5, grid5, prm5).plan(sc5.ego)
        q.put((rank, loc["problems"], gi, gc, {k: np.asarray(v) for k, v in out5.items()}))
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
    covered = []
    for rank, problems, gi, gc, out5 in res:
        covered.append(problems)
        np.testing.assert_array_equal(gi, ref["best_idx"])
        np.testing.assert_array_equal(gc[gi >= 0], ref["best_cost"][gi >= 0])
        np.testing.assert_array_equal(out5["best_idx"], ref5["best_idx"])
        np.testing.assert_array_equal(out5["best_cost"], ref5["best_cost"])
        np.testing.assert_array_equal(out5["records"], ref5["records"])
        np.testing.assert_array_equal(out5["meta"], ref5["meta"])
    assert covered[0][0] == 0 and covered[-1][1] == 64 and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
