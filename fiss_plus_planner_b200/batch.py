"""Batched / multi-GPU front end: one process per GPU, ``torch.distributed`` for the plumbing.

The reference plans one ego state at a time in one Python thread (planning.py:120-128).  Independent
planning problems (ego states, scenarios) have no data flow between them, so they shard across
ranks with NO collective on the data path (BASELINE config 4):

    ``ShardedBatchPlanner``   rank r evaluates problems [lo_r, hi_r) of the batch on its own GPU.

A single problem with a big lattice (BASELINE config 5, B = 1) shards the lattice itself: each rank
evaluates a slab of lateral rows -- a contiguous candidate-id range in FrenetOptimalPlanner numbering
-- and the ranks agree on the winner with ONE tiny all-reduce pair that reproduces the reference's
tie rule (``min_cost >= cost`` scan => the LARGEST index among the minima,
frenet_optimal_planner.py:263-268):

    ``SplitLatticePlanner``   all_reduce(MIN) on the cost, all_reduce(MAX) on "my index if my cost is
                              the minimum else -1"; the owner of the winner broadcasts its record.

Tensors are device buffers only (NCCL needs them on the GPU); on CPU (gloo) the same reduction code
runs on host tensors, which is what the world_size-2 tests exercise.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from fiss_plus_planner_b200 import _shim
from fiss_plus_planner_b200.engine import FissEngine, LatticeGrid


def shard_range(n: int, world: int, rank: int) -> tuple:
    """Contiguous, balanced split of ``range(n)``: the first ``n % world`` ranks get one extra item."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def allreduce_pick(best_cost: torch.Tensor, best_idx: torch.Tensor, group=None):
    """Cross-rank argmin with the reference's tie rule, in place.

    ``best_cost [P]`` float64 (``+inf`` where the rank has no feasible candidate) and ``best_idx [P]``
    int64 GLOBAL candidate ids (``-1`` for none).  After the call every rank holds the global minimum
    cost and, among the ranks that attain it, the largest candidate id -- "last minimal-cost survivor
    wins" (frenet_optimal_planner.py:263-268).  16 bytes per problem on the wire."""
    world, _ = _world(group)
    if world == 1:
        return best_cost, best_idx
    mine = best_cost.clone()
    dist.all_reduce(best_cost, op=dist.ReduceOp.MIN, group=group)
    claim = torch.where((mine == best_cost) & (best_idx >= 0), best_idx, torch.full_like(best_idx, -1))
    dist.all_reduce(claim, op=dist.ReduceOp.MAX, group=group)
    best_idx.copy_(claim)
    return best_cost, best_idx


class ShardedBatchPlanner(object):
    """BASELINE config 4: a batch of ego states over one shared lattice, problems sharded across ranks."""

    def __init__(self, engine: FissEngine, grid: LatticeGrid, params, group=None):
        self.engine, self.grid, self.params, self.group = engine, grid, params, group
        self.world, self.rank = _world(group)

    def local_slice(self, batch: int) -> slice:
        lo, hi = shard_range(batch, self.world, self.rank)
        return slice(lo, hi)

    def plan_local(self, ego_all: np.ndarray, want_records: bool = True) -> dict:
        """Evaluate this rank's shard of ``ego_all [B, 6]`` (every rank passes the same array, or just
        its shard with ``already_sharded``).  No communication."""
        sl = self.local_slice(len(ego_all))
        out = self.engine.plan_grid(np.ascontiguousarray(ego_all[sl]), self.grid, self.params,
                                    want_records=want_records, want_volume=False)
        out["problems"] = (sl.start, sl.stop)
        return out

    def gather_winners(self, local: dict, batch: int, device=None):
        """Optional: publish every problem's (winner id, cost) to all ranks (one all_gather of 16 B/problem)."""
        idx = np.full(batch, -1, np.int64)
        cost = np.full(batch, np.inf)
        lo, hi = local["problems"]
        idx[lo:hi] = local["best_idx"]
        cost[lo:hi] = np.where(local["best_idx"] >= 0, local["best_cost"], np.inf)
        if self.world == 1:
            return idx, cost
        dev = device if device is not None else ("cuda" if dist.get_backend(self.group) == "nccl" else "cpu")
        t_idx = torch.tensor(idx, device=dev)
        t_cost = torch.tensor(cost, device=dev)
        # every problem is owned by exactly one rank: MAX over {-1, id} and MIN over {inf, cost} gather them
        dist.all_reduce(t_idx, op=dist.ReduceOp.MAX, group=self.group)
        dist.all_reduce(t_cost, op=dist.ReduceOp.MIN, group=self.group)
        return t_idx.cpu().numpy(), t_cost.cpu().numpy()


class SplitLatticePlanner(object):
    """BASELINE config 5 with few problems: the lattice's lateral rows are split across ranks.

    Requires FrenetOptimalPlanner numbering (``order == "dtv"``: d outermost), so a slab of lateral rows
    [i_lo, i_hi) is the contiguous global id range [i_lo * stride_d, i_hi * stride_d)."""

    def __init__(self, engine: FissEngine, grid: LatticeGrid, params, group=None):
        assert grid.order[0] == "d", "split along the outermost axis: use FrenetOptimalPlanner numbering"
        self.engine, self.grid, self.params, self.group = engine, grid, params, group
        self.world, self.rank = _world(group)
        self.i_lo, self.i_hi = shard_range(len(grid.d), self.world, self.rank)
        self.id_offset = self.i_lo * grid.strides[0]
        self.local_grid = (LatticeGrid(grid.d[self.i_lo:self.i_hi], grid.v, grid.T, grid.tick, grid.order)
                           if self.i_hi > self.i_lo else None)

    def owner_of(self, global_idx: int) -> int:
        i_d = int(global_idx) // self.grid.strides[0]
        for r in range(self.world):
            lo, hi = shard_range(len(self.grid.d), self.world, r)
            if lo <= i_d < hi:
                return r
        raise ValueError(global_idx)

    def plan(self, ego: np.ndarray, device=None) -> dict:
        """``ego [B, 6]`` (same on every rank) -> global winners on every rank."""
        ego = np.ascontiguousarray(np.atleast_2d(ego), dtype=np.float64)
        b = ego.shape[0]
        if self.local_grid is not None:
            out = self.engine.plan_grid(ego, self.local_grid, self.params, want_records=True, want_volume=False)
            idx = np.where(out["best_idx"] >= 0, out["best_idx"].astype(np.int64) + self.id_offset, -1)
            cost = np.where(out["best_idx"] >= 0, out["best_cost"], np.inf)
            rec, meta = out["records"], out["meta"]
        else:
            idx, cost = np.full(b, -1, np.int64), np.full(b, np.inf)
            rec = np.full((b, _shim.REC_ROWS, self.grid.n_stride), np.nan)
            meta = np.zeros((b, 2), np.int32)
        if self.world == 1:
            return dict(best_idx=idx, best_cost=cost, records=rec, meta=meta)
        dev = device if device is not None else ("cuda" if dist.get_backend(self.group) == "nccl" else "cpu")
        t_cost = torch.tensor(cost, device=dev)     # copies: the reduction is in place and must not alias idx / cost
        t_idx = torch.tensor(idx, device=dev)
        allreduce_pick(t_cost, t_idx, self.group)
        g_idx, g_cost = t_idx.cpu().numpy(), t_cost.cpu().numpy()
        # the winners' arrays live on their owners: zero everything else and sum (one all_reduce)
        n_stride = self.grid.n_stride
        mine = (idx == g_idx) & (g_idx >= 0)
        rec_full = np.zeros((b, _shim.REC_ROWS, n_stride))
        rec_full[mine] = np.nan_to_num(rec[mine, :, :n_stride], nan=0.0)
        nan_mask = np.zeros((b, _shim.REC_ROWS, n_stride))
        nan_mask[mine] = np.isnan(rec[mine, :, :n_stride])
        meta_full = np.where(mine[:, None], meta, 0).astype(np.int64)
        t_rec = torch.as_tensor(np.stack((rec_full, nan_mask)), device=dev)
        t_meta = torch.as_tensor(meta_full, device=dev)
        dist.all_reduce(t_rec, op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(t_meta, op=dist.ReduceOp.SUM, group=self.group)
        both = t_rec.cpu().numpy()
        records = np.where(both[1] > 0, np.nan, both[0])
        records[g_idx < 0] = np.nan
        return dict(best_idx=g_idx, best_cost=np.where(g_idx >= 0, g_cost, np.inf), records=records,
                    meta=t_meta.cpu().numpy().astype(np.int32))
