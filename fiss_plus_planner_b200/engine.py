"""``FissEngine``: thin Python owner of one ``fiss_handle`` (one GPU, one stream).

Host-side marshalling only: NumPy arrays in, NumPy arrays out; the arithmetic is in
csrc/fiss_kernels.cuh.  The planners (``planners/*.py``) and the batched / multi-GPU front end
(``batch.py``) sit on top of this class.  Device-pointer methods take torch tensors used purely as
device buffers (``.data_ptr()``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from fiss_plus_planner_b200 import _shim
from fiss_plus_planner_b200._shim import FissError, FissGrid, FissParams


def end_state_table(d, v, T, tick: float) -> np.ndarray:
    """``[C, 4] = (d_end, v_end, T, n)`` with ``n = len(np.arange(0, T, tick))``."""
    d, v, T = np.broadcast_arrays(np.asarray(d, np.float64), np.asarray(v, np.float64), np.asarray(T, np.float64))
    # len(np.arange(0, T, tick)) = ceil((T - 0) / tick) evaluated in float64 -- the same two IEEE
    # operations as fiss_arange_len in the C ABI (tests/test_abi_exports.py pins both against NumPy)
    n = np.ceil((T.ravel() - 0.0) / float(tick))
    return np.ascontiguousarray(np.column_stack((d.ravel(), v.ravel(), T.ravel(), n)))


class LatticeGrid(object):
    """A product lattice ``d_end x v_end x T`` plus its candidate numbering (``struct fiss_grid``).

    ``order`` names the loop nest that numbers the candidates, outermost first: ``"dtv"`` is
    FrenetOptimalPlanner's (frenet_optimal_planner.py:75,78,89), ``"dvt"`` FissPlanner's grid
    ``[i_d][j_v][k_t]`` (fiss_planner.py:48,60,70)."""

    def __init__(self, d, v, T, tick: float, order: str = "dtv"):
        self.d = np.ascontiguousarray(d, dtype=np.float64)
        self.v = np.ascontiguousarray(v, dtype=np.float64)
        self.T = np.ascontiguousarray(T, dtype=np.float64)
        assert sorted(order) == ["d", "t", "v"], order
        self.order = order
        self.tick = float(tick)
        size = {"d": len(self.d), "v": len(self.v), "t": len(self.T)}
        stride, acc = {}, 1
        for ax in reversed(order):
            stride[ax] = acc
            acc *= size[ax]
        self.strides = (stride["d"], stride["v"], stride["t"])
        self.n = np.array([_shim.arange_len(t, tick) for t in self.T], dtype=np.int64)
        self.num_candidates = acc
        self.n_stride = int(self.n.max())
        self._table = None
        dp = C.POINTER(C.c_double)
        self.c_struct = FissGrid(self.d.ctypes.data_as(dp), self.v.ctypes.data_as(dp), self.T.ctypes.data_as(dp),
                                 len(self.d), len(self.v), len(self.T), *self.strides)

    @property
    def shape(self):
        return len(self.d), len(self.v), len(self.T)

    def index3(self, c):
        """flat candidate id -> (i_d, j_v, k_t)."""
        c = np.asarray(c)
        sd, sv, st = self.strides
        return (c // sd) % len(self.d), (c // sv) % len(self.v), (c // st) % len(self.T)

    def table(self) -> np.ndarray:
        """The expanded ``[C, 4] = (d_end, v_end, T, n)`` end-state table in candidate order."""
        if self._table is None:
            i, j, k = self.index3(np.arange(self.num_candidates))
            self._table = np.ascontiguousarray(np.column_stack((self.d[i], self.v[j], self.T[k],
                                                                self.n[k].astype(np.float64))))
        return self._table


def fop_grid(settings, vehicle_w: float) -> LatticeGrid:
    """FrenetOptimalPlanner's lattice (frenet_optimal_planner.py:72-78,89) as a ``LatticeGrid``."""
    sw = settings.max_road_width - vehicle_w
    return LatticeGrid(np.linspace(-sw / 2, sw / 2, settings.num_width),
                       np.linspace(settings.lowest_speed, settings.highest_speed, settings.num_speed),
                       np.linspace(settings.min_t, settings.max_t, settings.num_t), settings.tick_t, "dtv")


def fiss_grid(settings, vehicle_w: float) -> LatticeGrid:
    """FissPlanner's lattice (fiss_planner.py:40-70: width + 0.3, grid [i_d][j_v][k_t])."""
    sw = settings.max_road_width - vehicle_w + 0.3
    return LatticeGrid(np.linspace(-sw / 2, sw / 2, settings.num_width),
                       np.linspace(settings.lowest_speed, settings.highest_speed, settings.num_speed),
                       np.linspace(settings.min_t, settings.max_t, settings.num_t), settings.tick_t, "dvt")


def fop_lattice(settings, vehicle_w: float) -> np.ndarray:
    """End states in FrenetOptimalPlanner.calc_frenet_paths order: d outer, T middle, v inner
    (frenet_optimal_planner.py:72-78,89)."""
    sw = settings.max_road_width - vehicle_w
    ds = np.linspace(-sw / 2, sw / 2, settings.num_width)
    Ts = np.linspace(settings.min_t, settings.max_t, settings.num_t)
    vs = np.linspace(settings.lowest_speed, settings.highest_speed, settings.num_speed)
    dd, TT, vv = np.meshgrid(ds, Ts, vs, indexing="ij")
    return end_state_table(dd, vv, TT, settings.tick_t)


def fiss_lattice(settings, vehicle_w: float):
    """FISS grid ``[i_d][j_v][k_t]`` (fiss_planner.py:40-70): returns (table [C,4] in i,j,k order,
    d samples, v samples, t samples, resolutions[3])."""
    sw = settings.max_road_width - vehicle_w + 0.3
    ds, rd = np.linspace(-sw / 2, sw / 2, settings.num_width, retstep=True)
    vs, rv = np.linspace(settings.lowest_speed, settings.highest_speed, settings.num_speed, retstep=True)
    ts, rt = np.linspace(settings.min_t, settings.max_t, settings.num_t, retstep=True)
    dd, vv, tt = np.meshgrid(ds, vs, ts, indexing="ij")
    return end_state_table(dd, vv, tt, settings.tick_t), ds, vs, ts, np.array([rd, rv, rt])


def make_params(settings, vehicle, weights: dict, time_step_now: int = 0, check_res: int = 2,
                check_curvature: bool = False, collide_all: bool = False) -> FissParams:
    return FissParams(
        tick_t=settings.tick_t, target_speed=settings.highest_speed, max_speed=vehicle.max_speed,
        max_accel=vehicle.max_accel, max_curvature=float(getattr(vehicle, "max_curvature", np.inf)),
        ego_length=vehicle.l, ego_width=vehicle.w, cost_time_offset=weights["time_offset"],
        w_speed=weights["w_V"], w_accel=weights["w_A"], w_jerk=weights["w_J"], w_offset=weights["w_LC"],
        time_step_now=int(time_step_now), check_res=int(check_res), check_curvature=int(check_curvature),
        collide_all=int(collide_all))


def decode_flags(flags: np.ndarray):
    """flags word -> (constraint_ok, collision, n_cart)."""
    flags = np.asarray(flags)
    constraint_ok = (flags & (_shim.FLAG_SPEED | _shim.FLAG_ACCEL | _shim.FLAG_CURVATURE)) == 0
    collision = (flags & _shim.FLAG_COLLISION) != 0
    n_cart = (flags >> _shim.FLAG_NCART_SHIFT) & _shim.FLAG_NCART_MASK
    return constraint_ok, collision, n_cart.astype(np.int64)


class FissEngine:
    def __init__(self, device: int = 0):
        self._lib = _shim.load()
        self._h = C.c_void_p()
        rc = self._lib.fiss_create(int(device), C.byref(self._h))
        if rc != _shim.FISS_OK:
            raise FissError(f"fiss_create failed ({rc}): {self._lib.fiss_last_error(None).decode()}")
        self.device = int(device)
        self.num_obstacles = 0
        # scene versions: bumped whenever the installed reference line / obstacle table changes, so that several planners
        # (or a planner and its lazy candidate bundles) sharing one engine can tell that the tables are no longer theirs
        self.spline_token = 0
        self.obstacle_token = 0
        self.comm_world, self.comm_rank = 1, 0
        self._ego_stage = np.zeros((self._STAGE_ROWS, 6), np.float64)
        self._ego_stage_ptr = _shim.ptr(self._ego_stage)

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.fiss_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != _shim.FISS_OK:
            raise FissError(f"{what} failed ({rc}): {self._lib.fiss_last_error(self._h).decode()}")

    @property
    def launch_count(self) -> int:
        return int(self._lib.fiss_launch_count(self._h))

    @staticmethod
    def _stream(stream):
        return C.c_void_p(int(stream) if stream else 0)

    # ------------------------------------------------------------------ scene tables
    def set_spline(self, table: np.ndarray, stream=None):
        table = np.ascontiguousarray(table, dtype=np.float64)
        assert table.ndim == 2 and table.shape[0] == 9, "spline table must be [9, K]"
        self._check(self._lib.fiss_set_spline(self._h, self._stream(stream), _shim.ptr(table), table.shape[1]),
                    "fiss_set_spline")
        self.spline_token += 1

    def fit_splines(self, lanes, install: int = 0, stream=None) -> np.ndarray:
        """Natural cubic splines of ``lanes [L, K, 2]`` (or one lane ``[K, 2]``) fitted on the device (Thomas
        recurrence, one CTA per lane).  Returns the coefficient tables ``[L, 9, K]`` (the layout of
        ``CubicSpline2D.device_table()``); lane ``install`` (>= 0) becomes the engine's reference line."""
        lanes = np.ascontiguousarray(lanes, dtype=np.float64)
        if lanes.ndim == 2:
            lanes = lanes[None]
        n_lanes, k, two = lanes.shape
        assert two == 2, "lanes must be [L, K, 2]"
        tables = np.empty((n_lanes, 9, k), np.float64)
        self._check(self._lib.fiss_fit_splines_host(self._h, self._stream(stream), _shim.ptr(lanes), n_lanes, k,
                                                    _shim.ptr(tables), int(install)), "fiss_fit_splines_host")
        if install >= 0:
            self.spline_token += 1
        return tables

    def frame_samples(self, s_end: float, step: float = 0.1, stream=None) -> np.ndarray:
        """``[m, 4] = (x, y, yaw, curvature)`` of the installed reference line at ``np.arange(0, s_end, step)``
        (generate_frenet_frame's polyline, frenet_optimal_planner.py:274-278), evaluated on the device."""
        m = _shim.arange_len(s_end, step)
        ref = np.empty((m, 4), np.float64)
        self._check(self._lib.fiss_frame_samples_host(self._h, self._stream(stream), float(step), m, _shim.ptr(ref)),
                    "fiss_frame_samples_host")
        return ref

    def set_obstacles(self, xyth, lw, valid, final_time_step: int, stream=None):
        if xyth is None or len(lw) == 0:
            self._check(self._lib.fiss_set_obstacles(self._h, self._stream(stream), None, None, None, 0, 0,
                                                     int(final_time_step)), "fiss_set_obstacles")
            self.num_obstacles = 0
            self.obstacle_token += 1
            return
        xyth = np.ascontiguousarray(xyth, dtype=np.float64)
        lw = np.ascontiguousarray(lw, dtype=np.float64)
        valid = np.ascontiguousarray(valid, dtype=np.uint8)
        m, t = valid.shape
        assert xyth.shape == (m, t, 3) and lw.shape == (m, 2)
        self._check(self._lib.fiss_set_obstacles(self._h, self._stream(stream), _shim.ptr(xyth), _shim.ptr(lw),
                                                 _shim.ptr(valid), m, t, int(final_time_step)), "fiss_set_obstacles")
        self.num_obstacles = m
        self.obstacle_token += 1

    def set_obstacles_waymo(self, trajs, mask, final_time_step: int = -1, stream=None) -> int:
        """Obstacle table straight from the Waymo wire format (``trajs [N, T, 11]`` float32, ``mask [N, T]``) with the
        keep / cut rules of ``convert_waymo_obstacle_to_cr`` (waymo_interface.py:24-76): returns the number of agents
        kept as obstacles.  ``final_time_step < 0`` takes it from the first kept agent, like the reference's list."""
        trajs = np.ascontiguousarray(trajs, dtype=np.float32)
        mask = np.ascontiguousarray(np.asarray(mask) != 0, dtype=np.uint8)
        n, t, f = trajs.shape
        assert f == 11 and mask.shape == (n, t)
        kept = C.c_int32(0)
        self._check(self._lib.fiss_set_obstacles_waymo(self._h, self._stream(stream), _shim.ptr(trajs), _shim.ptr(mask),
                                                       n, t, int(final_time_step), C.byref(kept)), "fiss_set_obstacles_waymo")
        self.num_obstacles = int(kept.value)
        self.obstacle_token += 1
        return self.num_obstacles

    # ------------------------------------------------------------------ host-pointer calls
    def plan_lattice(self, ego: np.ndarray, end: np.ndarray, params: FissParams, want_records: bool = True,
                     want_volume: bool = False, stream=None) -> dict:
        """plan() for ``ego [B, 6]`` over the shared end-state table ``end [C, 4]``."""
        ego = np.ascontiguousarray(np.atleast_2d(ego), dtype=np.float64)
        end = np.ascontiguousarray(end, dtype=np.float64)
        b, c = ego.shape[0], end.shape[0]
        n_stride = int(end[:, 3].max())
        best_idx = np.empty(b, np.int32)
        best_cost = np.empty(b, np.float64)
        meta = np.empty((b, 2), np.int32)
        records = np.empty((b, _shim.REC_ROWS, n_stride), np.float64) if want_records else None
        cost = np.empty((b, c), np.float64) if want_volume else None
        flags = np.empty((b, c), np.uint32) if want_volume else None
        self._check(self._lib.fiss_plan_lattice_host(
            self._h, self._stream(stream), _shim.ptr(ego), b, _shim.ptr(end), c, C.byref(params), _shim.ptr(best_idx),
            _shim.ptr(best_cost), _shim.ptr(meta), _shim.ptr(records), n_stride, _shim.ptr(cost), _shim.ptr(flags)),
            "fiss_plan_lattice_host")
        return dict(best_idx=best_idx, best_cost=best_cost, meta=meta, records=records, cost=cost, flags=flags)

    @staticmethod
    def pinned_empty(shape, dtype) -> np.ndarray:
        """Page-locked host array (a torch pinned tensor viewed as NumPy; the view keeps the tensor alive).  The
        ``*_host`` entry points DMA straight from / into such arrays instead of staging + memcpy."""
        import torch
        tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32, np.dtype(np.uint32): torch.int32}[np.dtype(dtype)]
        arr = torch.empty(tuple(int(v) for v in np.atleast_1d(shape)), dtype=tdt, pin_memory=True).numpy()
        return arr.view(dtype)

    def alloc_plan_outputs(self, batch: int, grid: LatticeGrid, want_records: bool = True, want_volume: bool = False,
                           pinned: bool = True) -> dict:
        """Reusable output buffers for ``plan_grid(..., out=)``: winners, (n, n') and optionally the winners' full
        records and the whole cost / flags volume."""
        mk = self.pinned_empty if pinned else (lambda shape, dt: np.empty(shape, dt))
        b, c, ns = int(batch), grid.num_candidates, grid.n_stride
        return dict(best_idx=mk((b,), np.int32), best_cost=mk((b,), np.float64), meta=mk((b, 2), np.int32),
                    records=mk((b, _shim.REC_ROWS, ns), np.float64) if want_records else None,
                    cost=mk((b, c), np.float64) if want_volume else None,
                    flags=mk((b, c), np.uint32) if want_volume else None)

    _OUT_KEYS = ("best_idx", "best_cost", "meta", "records", "cost", "flags")
    _STAGE_ROWS = 64

    def _out_ptrs(self, out: dict):
        """ctypes pointers of an output set, made once (``ndarray.ctypes`` costs microseconds per array -- more than the
        rest of the call's host side); kept in the dict under ``"_ptrs"``."""
        ptrs = out.get("_ptrs")
        if ptrs is None:
            ptrs = tuple(_shim.ptr(out[k]) for k in self._OUT_KEYS)
            out["_ptrs"] = ptrs
        return ptrs

    def plan_grid(self, ego: np.ndarray, grid: LatticeGrid, params: FissParams, want_records: bool = True,
                  want_volume: bool = False, stream=None, out: dict = None) -> dict:
        """plan() for ``ego [B, 6]`` over a product lattice (the lattice kernel).  ``out`` (from
        ``alloc_plan_outputs``) is filled in place and returned: no allocation on the call path, and pinned
        buffers -- ``ego`` included -- are the DMA endpoints themselves."""
        b = 1 if getattr(ego, "ndim", 2) == 1 else len(ego)
        if b <= self._STAGE_ROWS:
            # small batches (the plan-cycle latency path): into the engine's own staging rows, whose pointer is cached
            self._ego_stage[:b] = ego
            ego_ptr = self._ego_stage_ptr
        else:
            if not (isinstance(ego, np.ndarray) and ego.ndim == 2 and ego.dtype == np.float64 and ego.flags.c_contiguous):
                ego = np.ascontiguousarray(np.atleast_2d(ego), dtype=np.float64)
            ego_ptr = _shim.ptr(ego)
        n_stride = grid.n_stride
        if out is None:
            out = self.alloc_plan_outputs(b, grid, want_records, want_volume, pinned=False)
        else:
            assert out["best_idx"].shape == (b,) and (out["records"] is None or out["records"].shape[2] == n_stride)
        p = self._out_ptrs(out)
        rc = self._lib.fiss_plan_grid_host(self._h, self._stream(stream), ego_ptr, b, C.byref(grid.c_struct), C.byref(params),
                                           p[0], p[1], p[2], p[3], n_stride, p[4], p[5])
        if rc != _shim.FISS_OK:
            self._check(rc, "fiss_plan_grid_host")
        return out

    def plan_grid_submit(self, lane: int, ego: np.ndarray, grid: LatticeGrid, params: FissParams, out: dict, stream=None):
        """Streaming ``plan_grid``: enqueue batch ``ego [B, 6]`` on ``lane`` (0 .. 3) and return at once; the results land
        in ``out`` (from ``alloc_plan_outputs(..., want_volume=False)``) when ``plan_grid_wait(lane)`` returns.  The
        copy-back of one lane overlaps the kernels of the others."""
        assert isinstance(ego, np.ndarray) and ego.ndim == 2 and ego.dtype == np.float64 and ego.flags.c_contiguous
        p = self._out_ptrs(out)
        self._check(self._lib.fiss_plan_grid_submit(
            self._h, self._stream(stream), int(lane), _shim.ptr(ego), ego.shape[0], C.byref(grid.c_struct), C.byref(params),
            p[0], p[1], p[2], p[3], grid.n_stride), "fiss_plan_grid_submit")

    def plan_grid_wait(self, lane: int):
        self._check(self._lib.fiss_plan_grid_wait(self._h, int(lane)), "fiss_plan_grid_wait")

    def eval_end_states(self, ego6: np.ndarray, end: np.ndarray, params: FissParams, want_records: bool = False,
                        stream=None) -> dict:
        ego6 = np.ascontiguousarray(ego6, dtype=np.float64).reshape(6)
        end = np.ascontiguousarray(end, dtype=np.float64)
        n_end = end.shape[0]
        n_stride = int(end[:, 3].max())
        cost = np.empty(n_end, np.float64)
        flags = np.empty(n_end, np.uint32)
        records = np.empty((n_end, _shim.REC_ROWS, n_stride), np.float64) if want_records else None
        self._check(self._lib.fiss_eval_end_states_host(
            self._h, self._stream(stream), _shim.ptr(ego6), _shim.ptr(end), n_end, C.byref(params), _shim.ptr(cost),
            _shim.ptr(flags), _shim.ptr(records), n_stride), "fiss_eval_end_states_host")
        return dict(cost=cost, flags=flags, records=records)

    # ------------------------------------------------------------------ device-pointer calls (torch tensors as buffers)
    def eval_candidates_dev(self, ego_t, end_t, params: FissParams, cost_t, flags_t, mat_t, n_stride: int, stream=None):
        b, c = ego_t.shape[0], end_t.shape[0]
        self._check(self._lib.fiss_eval_candidates_dev(
            self._h, self._stream(stream), C.c_void_p(ego_t.data_ptr()), b, C.c_void_p(end_t.data_ptr()), c,
            C.byref(params), C.c_void_p(cost_t.data_ptr()), C.c_void_p(flags_t.data_ptr()),
            C.c_void_p(mat_t.data_ptr()) if mat_t is not None else None, int(n_stride)), "fiss_eval_candidates_dev")

    def eval_grid_dev(self, ego_t, grid: LatticeGrid, params: FissParams, cost_t, flags_t, mat_t, n_stride: int,
                      stream=None):
        self._check(self._lib.fiss_eval_grid_dev(
            self._h, self._stream(stream), C.c_void_p(ego_t.data_ptr()), ego_t.shape[0], C.byref(grid.c_struct),
            C.byref(params), C.c_void_p(cost_t.data_ptr()), C.c_void_p(flags_t.data_ptr()),
            C.c_void_p(mat_t.data_ptr()) if mat_t is not None else None, int(n_stride)), "fiss_eval_grid_dev")

    def plan_grid_dev(self, ego_t, grid: LatticeGrid, params: FissParams, cost_t, flags_t, mat_t, best_idx_t, best_cost_t,
                      meta_t, records_t, n_stride: int, stream=None):
        """One plan step on device buffers (lattice kernel + pick / records) -- one graph launch from the second call on."""
        vp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
        self._check(self._lib.fiss_plan_grid_dev(
            self._h, self._stream(stream), vp(ego_t), ego_t.shape[0], C.byref(grid.c_struct), C.byref(params), vp(cost_t),
            vp(flags_t), vp(mat_t), vp(best_idx_t), vp(best_cost_t), vp(meta_t), vp(records_t), int(n_stride)),
            "fiss_plan_grid_dev")

    # ------------------------------------------------------------------ multi-GPU pick (NCCL inside the C-ABI)
    def comm_init(self, group=None, nccl_path: str = None):
        """Create the handle's own NCCL communicator over the ranks of ``group`` (torch.distributed is only the
        out-of-band channel for the 128-byte unique id).  ``fiss_allreduce_pick`` then runs on it."""
        import torch
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        path = nccl_path.encode() if nccl_path else None
        ident = (C.c_ubyte * 128)()
        if rank == 0:
            self._check(self._lib.fiss_comm_unique_id(ident, path), "fiss_comm_unique_id")
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = (C.c_ubyte * 128).from_buffer_copy(box[0])
        torch.cuda.synchronize(self.device)
        self._check(self._lib.fiss_comm_init(self._h, ident, world, rank, path), "fiss_comm_init")
        self.comm_world, self.comm_rank = world, rank

    def comm_destroy(self):
        self._check(self._lib.fiss_comm_destroy(self._h), "fiss_comm_destroy")
        self.comm_world, self.comm_rank = 1, 0

    def allreduce_pick_dev(self, best_idx_t, best_cost_t, meta_t, records_t, n_stride: int, id_offset: int,
                           id_inner: int = 1 << 40, id_outer: int = 0, stream=None):
        """In place on device buffers: local slab winners -> the global winners on every rank (``fiss_allreduce_pick``
        on the handle's communicator): one all-reduce for the pick, one for the winners' records, no host sync."""
        vp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
        self._check(self._lib.fiss_allreduce_pick(self._h, None, 0, 0, self._stream(stream), best_idx_t.shape[0],
                                                  int(id_inner), int(id_outer), int(id_offset), vp(best_idx_t), vp(best_cost_t), vp(meta_t), vp(records_t),
                                                  int(n_stride)), "fiss_allreduce_pick")

    def pick_winners_dev(self, ego_t, end_t, params: FissParams, cost_t, flags_t, best_idx_t, best_cost_t,
                         records_t, meta_t, n_stride: int, stream=None):
        b, c = ego_t.shape[0], end_t.shape[0]
        self._check(self._lib.fiss_pick_winners_dev(
            self._h, self._stream(stream), C.c_void_p(ego_t.data_ptr()), b, C.c_void_p(end_t.data_ptr()), c,
            C.byref(params), C.c_void_p(cost_t.data_ptr()), C.c_void_p(flags_t.data_ptr()),
            C.c_void_p(best_idx_t.data_ptr()), C.c_void_p(best_cost_t.data_ptr()),
            C.c_void_p(records_t.data_ptr()) if records_t is not None else None,
            C.c_void_p(meta_t.data_ptr()) if meta_t is not None else None, int(n_stride)), "fiss_pick_winners_dev")
