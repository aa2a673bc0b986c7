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

    ``SplitLatticePlanner``   on GPUs (NCCL): everything stays on the device -- ``fiss_plan_grid_dev`` picks the slab's
                              winner, ``fiss_allreduce_pick`` (C-ABI, NCCL called from libfissgpu.so on the kernels'
                              stream) agrees on the global one with ONE all-reduce and moves the winner's record with
                              a second; a single device->host copy at the end.
                              On CPU (gloo; the world_size-2 tests) the same rule runs on host tensors:
                              all_reduce(MIN) on the cost, all_reduce(MAX) on "my index if my cost is the minimum
                              else -1", records summed from their owners.
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
    """BASELINE config 5 with few problems: ONE lattice split across ranks, FrenetOptimalPlanner numbering (``"dtv"``).

    ``axis="t"`` (default) gives rank r a slab of horizons ``T[k_lo:k_hi]``: ALL the per-problem work divides -- the
    longitudinal rows are per (v, T), the lateral rows per (d, T) -- at the price of balance being bounded by the number
    of horizons (9 horizons on 8 ranks: 2, 1, 1, ...).  ``axis="d"`` gives a slab of lateral rows (a contiguous global id
    range): every rank then recomputes all the longitudinal rows, so only the lateral / per-candidate work divides.
    Local candidate ids map to global ones by ``(c // inner) * outer + c % inner + offset`` (``id_map``)."""

    def __init__(self, engine: FissEngine, grid: LatticeGrid, params, group=None, axis: str = "t"):
        assert grid.order == "dtv", "FrenetOptimalPlanner numbering (d outer, T, v inner) is required"
        assert axis in ("t", "d")
        self.engine, self.grid, self.params, self.group, self.axis = engine, grid, params, group, axis
        self.world, self.rank = _world(group)
        nd, nv, nt = grid.shape
        if axis == "d":
            self.lo, self.hi = shard_range(nd, self.world, self.rank)
            self.id_map = (1 << 40, 0, self.lo * grid.strides[0])
            self.local_grid = (LatticeGrid(grid.d[self.lo:self.hi], grid.v, grid.T, grid.tick, grid.order)
                               if self.hi > self.lo else None)
        else:
            self.lo, self.hi = shard_range(nt, self.world, self.rank)
            self.id_map = ((self.hi - self.lo) * nv if self.hi > self.lo else 1, nt * nv, self.lo * nv)
            self.local_grid = (LatticeGrid(grid.d, grid.v, grid.T[self.lo:self.hi], grid.tick, grid.order)
                               if self.hi > self.lo else None)
        self.id_offset = self.id_map[2]

    def to_global(self, local_idx):
        """Local candidate ids of this rank's slab -> ids of the full lattice (-1 stays -1)."""
        c = np.asarray(local_idx).astype(np.int64)
        inner, outer, offset = self.id_map
        return np.where(c >= 0, (c // inner) * outer + c % inner + offset, -1)

    # ------------------------------------------------------------------ device path (NCCL inside the C-ABI)
    def _device_buffers(self, b: int):
        key = (b, self.grid.n_stride, self.local_grid.num_candidates if self.local_grid is not None else 0)
        if getattr(self, "_dev_key", None) != key:
            dev = torch.device("cuda", self.engine.device)
            c, ns = key[2], key[1]
            f64, i32 = torch.float64, torch.int32
            self._dev = dict(ego=torch.empty((b, 6), dtype=f64, device=dev), cost=torch.empty(max(b * c, 1), dtype=f64, device=dev),
                             flags=torch.empty(max(b * c, 1), dtype=i32, device=dev), idx=torch.empty(b, dtype=i32, device=dev),
                             best=torch.empty(b, dtype=f64, device=dev), meta=torch.empty((b, 2), dtype=i32, device=dev),
                             rec=torch.empty((b, _shim.REC_ROWS, ns), dtype=f64, device=dev),
                             ego_pin=torch.empty((b, 6), dtype=f64).pin_memory())
            self._dev_key = key
        return self._dev

    def plan_step_dev(self, stream=None):
        """One cycle on the device buffers (``self._dev['ego']`` already holds the ego states): the slab's lattice kernel
        + pick / records, then the cross-GPU pick.  Asynchronous; the results are in ``self._dev`` (idx / best / meta /
        rec) on every rank."""
        d, ns = self._dev, self.grid.n_stride
        if self.local_grid is not None:
            self.engine.plan_grid_dev(d["ego"], self.local_grid, self.params, d["cost"], d["flags"], None, d["idx"], d["best"],
                                      d["meta"], d["rec"], ns, stream=stream)
        else:   # more ranks than lateral rows: this rank has no candidates
            d["idx"].fill_(-1)
            d["best"].fill_(float("inf"))
        if self.world > 1:
            inner, outer, offset = self.id_map
            self.engine.allreduce_pick_dev(d["idx"], d["best"], d["meta"], d["rec"], ns, offset, id_inner=inner, id_outer=outer,
                                           stream=stream)

    def _use_device_path(self) -> bool:
        return (self.world > 1 and dist.get_backend(self.group) == "nccl" and torch.cuda.is_available()
                and getattr(self.engine, "comm_world", 1) == self.world)

    def plan(self, ego: np.ndarray, device=None) -> dict:
        """``ego [B, 6]`` (same on every rank) -> global winners on every rank."""
        ego = np.ascontiguousarray(np.atleast_2d(ego), dtype=np.float64)
        b = ego.shape[0]
        if self.world > 1 and dist.get_backend(self.group) == "nccl" and getattr(self.engine, "comm_world", 1) != self.world:
            self.engine.comm_init(self.group)
        if self._use_device_path():
            d = self._device_buffers(b)
            with torch.cuda.device(self.engine.device):
                stream = torch.cuda.current_stream()
                d["ego_pin"].copy_(torch.from_numpy(ego))
                d["ego"].copy_(d["ego_pin"], non_blocking=True)
                self.plan_step_dev(stream=stream.cuda_stream)
                idx, cost = d["idx"].cpu().numpy().astype(np.int64), d["best"].cpu().numpy()   # syncs the stream
                return dict(best_idx=idx, best_cost=cost, records=d["rec"].cpu().numpy(), meta=d["meta"].cpu().numpy())
        if self.local_grid is not None:
            out = self.engine.plan_grid(ego, self.local_grid, self.params, want_records=True, want_volume=False)
            idx = self.to_global(out["best_idx"])
            cost = np.where(out["best_idx"] >= 0, out["best_cost"], np.inf)
            rec, meta = out["records"], out["meta"]
            if rec.shape[2] < self.grid.n_stride:       # a slab of short horizons: pad the rows to the full lattice's pitch
                pad = np.full((b, _shim.REC_ROWS, self.grid.n_stride), np.nan)
                pad[:, :, :rec.shape[2]] = rec
                rec = pad
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
