"""N > 1 host logic on CPU: world_size-2 ``gloo`` runs of the sharding helpers, the cross-rank pick
(reference tie rule) and the split-lattice planner against a single-rank run of the same fake engine."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class FakeEngine(object):
    """Stands in for FissEngine.plan_grid on a box without a GPU: a deterministic cost over the lattice
    (with exact ties across lateral rows), infeasible candidates marked by +inf, and 'records' that encode
    which candidate they belong to."""

    def plan_grid(self, ego, grid, params, want_records=True, want_volume=False):
        tab = grid.table()
        b = len(ego)
        best_idx = np.full(b, -1, np.int32)
        best_cost = np.full(b, np.inf)
        meta = np.zeros((b, 2), np.int32)
        rec = np.full((b, 16, grid.n_stride), np.nan)
        for p in range(b):
            cost = np.abs(np.abs(tab[:, 0]) - ego[p, 3]) + 0.01 * np.abs(tab[:, 1] - ego[p, 1]) + 0.001 * tab[:, 2]
            cost = np.where(np.abs(tab[:, 0]) > ego[p, 0], np.inf, cost)        # "infeasible"
            if np.isfinite(cost).any():
                m = cost.min()
                c = int(np.flatnonzero(cost == m).max())                       # last minimum wins
                best_idx[p], best_cost[p] = c, m
                n = int(tab[c, 3])
                meta[p] = (n, n - 1)
                rec[p, :, :n] = tab[c, 0] * 1000 + tab[c, 1] * 10 + tab[c, 2] + np.arange(16)[:, None]
        return dict(best_idx=best_idx, best_cost=best_cost, meta=meta, records=rec if want_records else None,
                    cost=None, flags=None)


def _grid():
    from fiss_plus_planner_b200.engine import LatticeGrid
    return LatticeGrid(np.arange(-4, 5) * 0.2, np.linspace(0, 13, 6), np.linspace(4, 5, 5), 0.1, "dtv")   # exactly symmetric rows


def _ego():
    # d0 = 0.4 ties +-0.4 rows (different ranks at world 2); one problem with nothing feasible
    return np.array([[9.0, 3.0, 0.0, 0.4, 0.0, 0.0], [9.0, 13.0, 0.0, 0.0, 0.0, 0.0], [-1.0, 5.0, 0.0, 0.2, 0.0, 0.0],
                     [0.3, 7.0, 0.0, 0.8, 0.0, 0.0]])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fiss_plus_planner_b200.batch import (ShardedBatchPlanner, SplitLatticePlanner, allreduce_pick,
                                                    shard_range)
        # --- allreduce_pick: ties, missing winners
        cost = torch.tensor([1.0, 2.0, float("inf"), 5.0, 3.0], dtype=torch.float64)
        idx = torch.tensor([10, 7, -1, 2, 4], dtype=torch.int64)
        if rank == 1:
            cost = torch.tensor([1.0, 1.5, float("inf"), float("inf"), 3.0], dtype=torch.float64)
            idx = torch.tensor([3, 9, -1, -1, 40], dtype=torch.int64)
        allreduce_pick(cost, idx)
        pick = (cost.tolist(), idx.tolist())
        # --- split lattice
        grid, ego = _grid(), _ego()
        sp = SplitLatticePlanner(FakeEngine(), grid, None, axis="d")
        out = sp.plan(ego)
        sp_t = SplitLatticePlanner(FakeEngine(), grid, None, axis="t")            # slabs of horizons: strided global ids
        out_t = sp_t.plan(ego)
        # --- sharded batch + gather
        sb = ShardedBatchPlanner(FakeEngine(), grid, None)
        loc = sb.plan_local(ego)
        gi, gc = sb.gather_winners(loc, len(ego))
        q.put((rank, pick, {k: np.asarray(v) for k, v in out.items()}, (sp.lo, sp.hi), loc["problems"], gi, gc,
               shard_range(10, world, rank), {k: np.asarray(v) for k, v in out_t.items()}, (sp_t.lo, sp_t.hi)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world2_gloo_pick_split_and_shard():
    from fiss_plus_planner_b200.batch import SplitLatticePlanner
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0

    ref = SplitLatticePlanner(FakeEngine(), _grid(), None).plan(_ego())            # single rank, whole lattice
    for rank, pick, out, rows, problems, gi, gc, sr, out_t, horizons in res:
        for k in ("best_idx", "best_cost", "meta", "records"):
            np.testing.assert_array_equal(out_t[k], ref[k], err_msg="horizon split: " + k)
        # min cost; among equal minima the LARGEST index (frenet_optimal_planner.py:263-268); -1 when nobody has one
        assert pick == ([1.0, 1.5, float("inf"), 5.0, 3.0], [10, 9, -1, 2, 40])
        np.testing.assert_array_equal(out["best_idx"], ref["best_idx"])
        np.testing.assert_array_equal(out["best_cost"], ref["best_cost"])
        np.testing.assert_array_equal(out["meta"], ref["meta"])
        np.testing.assert_array_equal(out["records"], ref["records"])               # NaN pattern included
        np.testing.assert_array_equal(gi, ref["best_idx"])
        np.testing.assert_array_equal(gc, ref["best_cost"])
    assert ref["best_idx"][2] == -1 and np.isnan(ref["records"][2]).all()
    # problem 0 ties the d = -0.4 and d = +0.4 rows (different ranks): the larger id, i.e. the +0.4 row, wins
    assert _grid().table()[ref["best_idx"][0], 0] == pytest.approx(0.4)
    assert [r[3] for r in res] == [(0, 5), (5, 9)]                                  # lateral rows per rank
    assert [r[9] for r in res] == [(0, 3), (3, 5)]                                  # horizons per rank
    assert [r[4] for r in res] == [(0, 2), (2, 4)]                                  # problems per rank
    assert [r[7] for r in res] == [(0, 5), (5, 10)]


def test_split_id_maps_are_order_preserving_bijections():
    """Local candidate ids of every slab map onto the full lattice's ids: together a partition, each map monotone (so
    the slab's "last minimum" is the global numbering's)."""
    from fiss_plus_planner_b200.batch import SplitLatticePlanner
    from fiss_plus_planner_b200.engine import LatticeGrid
    grid = LatticeGrid(np.linspace(-1, 1, 33), np.linspace(0, 13, 17), np.linspace(8, 10, 9), 0.1, "dtv")
    full = grid.table()
    for axis in ("t", "d"):
        for world in (1, 2, 3, 4, 8, 12):
            seen = []
            for rank in range(world):
                sp = SplitLatticePlanner.__new__(SplitLatticePlanner)
                sp.world, sp.rank = world, rank
                import fiss_plus_planner_b200.batch as batch
                orig = batch._world
                batch._world = lambda group=None, w=world, r=rank: (w, r)
                try:
                    sp.__init__(FakeEngine(), grid, None, axis=axis)
                finally:
                    batch._world = orig
                if sp.local_grid is None:
                    continue
                g = sp.to_global(np.arange(sp.local_grid.num_candidates))
                assert (np.diff(g) > 0).all()
                np.testing.assert_array_equal(sp.local_grid.table(), full[g])      # the same end states, the same order
                seen.append(g)
            np.testing.assert_array_equal(np.sort(np.concatenate(seen)), np.arange(grid.num_candidates))


def test_shard_range_is_a_partition():
    from fiss_plus_planner_b200.batch import shard_range
    for n in (0, 1, 7, 8, 4096, 5049):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_range(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
