"""FrenetOptimalPlanner on the B200 lattice engine -- drop-in for the reference class.

Reference: planners/frenet_optimal_planner.py (``Stats`` :15-36, ``FrenetOptimalPlannerSettings``
:38-56, ``FrenetOptimalPlanner`` :58-278).  Same constructor, attributes (``settings, vehicle,
cubic_spline, best_traj, all_trajs, stats``) and methods (``generate_frenet_frame``, ``plan``).

What differs is where the work happens.  The reference walks the (d, T, v) lattice in three nested
Python loops and then runs calc_global_paths / check_constraints / check_collisions over Python
lists.  Here ``plan()`` marshals six ego numbers, the lattice axes and one ``fiss_params`` struct,
and ONE call into libfissgpu.so (``fiss_plan_grid_host``: the lattice kernel) evaluates every
candidate, picks the winner with the reference's tie rule and returns the winner's arrays.  ``all_trajs`` stays
populated (the GIF renderer reads it, planning.py:352-355) but lazily: a cycle's candidate bundle
is only materialised on the device and copied back when somebody indexes it.
"""
from __future__ import annotations

import collections.abc

import numpy as np

from fiss_plus_planner_b200 import _shim
from fiss_plus_planner_b200.engine import FissEngine, decode_flags, fop_grid, make_params
from fiss_plus_planner_b200.planners.common.cost.cost_function import CostFunction
from fiss_plus_planner_b200.planners.common.geometry.cubic_spline import CubicSpline2D
from fiss_plus_planner_b200.planners.common.scenario.frenet import FrenetState, FrenetTrajectory
from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle


class Stats(object):
    def __init__(self):
        self.num_iter = 0
        self.num_trajs_generated = 0
        self.num_trajs_validated = 0
        self.num_collison_checks = 0

    def __add__(self, other):
        # accumulates in place and returns self, like the reference (:23-29)
        self.num_iter += other.num_iter
        self.num_trajs_generated += other.num_trajs_generated
        self.num_trajs_validated += other.num_trajs_validated
        self.num_collison_checks += other.num_collison_checks
        return self

    def average(self, value: int):
        self.num_iter /= value
        self.num_trajs_generated /= value
        self.num_trajs_validated /= value
        self.num_collison_checks /= value
        return self


class FrenetOptimalPlannerSettings(object):
    def __init__(self, num_width: int = 5, num_speed: int = 5, num_t: int = 5):
        self.tick_t = 0.1
        self.max_road_width = 3.5
        self.num_width = num_width
        self.highest_speed = 13.4112
        self.lowest_speed = 0.0
        self.num_speed = num_speed
        self.min_t = 8.0
        self.max_t = 10.0
        self.num_t = num_t
        self.check_obstacle = True   # declared, never read -- as in the reference (:54-55)
        self.check_boundary = True


# ------------------------------------------------------------------------------------------------
class ObstacleTable(object):
    """Dense device-ready obstacle predictions (the wire format of ``fiss_set_obstacles``)."""

    def __init__(self, xyth, lw, valid, final_time_step):
        self.xyth = np.ascontiguousarray(xyth, dtype=np.float64)
        self.lw = np.ascontiguousarray(lw, dtype=np.float64)
        self.valid = np.ascontiguousarray(valid, dtype=np.uint8)
        self.final_time_step = int(final_time_step)

    def __len__(self):
        return len(self.lw)


def _rectangle_size(shape):
    """(length, width) of a CommonRoad ``Rectangle``-like shape about the origin.  The reference intersects the true
    ``shapely_object`` (:189); the device predicate is rectangle-vs-rectangle, so anything that is not an axis-aligned
    rectangle centred on the origin is refused rather than silently replaced by its bounding box."""
    center = getattr(shape, "center", None)
    if center is not None and np.any(np.asarray(center, dtype=np.float64) != 0.0):
        raise NotImplementedError("obstacle shapes with a non-zero center are not supported by the device collision check")
    if getattr(shape, "orientation", 0.0) not in (0.0, None):
        raise NotImplementedError("obstacle shapes with a non-zero orientation are not supported by the device collision check")
    if hasattr(shape, "length") and hasattr(shape, "width"):
        return float(shape.length), float(shape.width)
    ring = getattr(shape, "shapely_object", None)
    if ring is not None:
        pts = np.asarray(getattr(ring, "pts", None) if hasattr(ring, "pts") else ring.exterior.coords, dtype=np.float64)
        if len(pts) >= 2 and np.array_equal(pts[0], pts[-1]):
            pts = pts[:-1]
        hl, hw = pts[:, 0].max(), pts[:, 1].max()
        corners = {(-hl, -hw), (-hl, hw), (hl, hw), (hl, -hw)}
        if len(pts) == 4 and {tuple(v) for v in pts} == corners:
            return float(2.0 * hl), float(2.0 * hw)
    raise NotImplementedError("obstacle_shape must be an axis-aligned rectangle about the origin (CommonRoad Rectangle); "
                              "circles, polygons and shape groups are not supported by the device collision check")


def marshal_obstacles(obstacles) -> ObstacleTable:
    """CommonRoad-style obstacle objects -> dense table.  Reads exactly what has_collision reads
    (:173,185-189): ``obstacles[0].prediction.final_time_step``, ``state_at_time(t)`` (None = absent)
    with ``.position`` / ``.orientation``, and the rectangle size."""
    if isinstance(obstacles, ObstacleTable):
        return obstacles
    if hasattr(obstacles, "xyth") and hasattr(obstacles, "lw"):     # synthetic.ObstacleSet
        return ObstacleTable(obstacles.xyth, obstacles.lw, obstacles.valid, obstacles.final_time_step)
    m = len(obstacles)
    if m == 0:
        return ObstacleTable(np.zeros((0, 0, 3)), np.zeros((0, 2)), np.zeros((0, 0)), 0)
    final = int(obstacles[0].prediction.final_time_step)
    horizon = final
    for ob in obstacles:
        pred = getattr(ob, "prediction", None)
        if pred is not None and getattr(pred, "final_time_step", None) is not None:
            horizon = max(horizon, int(pred.final_time_step))
    t_obs = horizon + 1
    xyth = np.zeros((m, t_obs, 3))
    valid = np.zeros((m, t_obs), dtype=np.uint8)
    lw = np.zeros((m, 2))
    for j, ob in enumerate(obstacles):
        lw[j] = _rectangle_size(ob.obstacle_shape)
        for t in range(t_obs):
            st = ob.state_at_time(t)
            if st is not None:
                xyth[j, t] = (st.position[0], st.position[1], st.orientation)
                valid[j, t] = 1
    return ObstacleTable(xyth, lw, valid, final)


class CandidateBundle(collections.abc.Sequence):
    """One cycle's candidates as a lazy sequence of ``FrenetTrajectory``.

    Costs and masks are already on the host; the per-step arrays are produced by one
    ``fiss_eval_end_states_host`` call (full records) the first time an element is touched.  The bundle remembers the
    reference line of its cycle (``spline_table`` + the engine's ``spline_token`` at that time): if the engine holds
    another line by then (a later ``generate_frenet_frame``, another planner on a shared engine) the bundle's own line
    is re-installed first, so the arrays are never computed against the wrong road."""

    def __init__(self, engine: FissEngine, ego6, end, params, cost, flags, idx3=None, spline_table=None,
                 spline_token=None):
        self._engine, self._ego6, self._end, self._params = engine, np.array(ego6), np.array(end), params
        self.cost = np.array(cost)
        self.flags = np.array(flags)
        self._idx3 = idx3
        self._items = None
        self._spline_table = spline_table
        self._spline_token = engine.spline_token if spline_token is None else spline_token

    def __len__(self):
        return len(self.cost)

    def _materialize(self):
        if self._items is None:
            if len(self.cost) == 0:
                self._items = []
                return self._items
            if self._engine.spline_token != self._spline_token:
                if self._spline_table is None:
                    raise RuntimeError("the engine's reference line changed since this cycle and the bundle holds no copy")
                self._engine.set_spline(self._spline_table)   # bumps the token: the owning planner re-installs its own
                self._spline_token = self._engine.spline_token
            out = self._engine.eval_end_states(self._ego6, self._end, self._params, want_records=True)
            _, _, n_cart = decode_flags(out["flags"])
            items = []
            for c in range(len(self.cost)):
                tr = FrenetTrajectory().fill_from_device_record(out["records"][c], int(self._end[c, 3]),
                                                                int(n_cart[c]), float(self.cost[c]))
                if self._idx3 is not None:
                    tr.idx = np.array(self._idx3[c])
                items.append(tr)
            self._items = items
        return self._items

    def __getitem__(self, i):
        return self._materialize()[i]


# ------------------------------------------------------------------------------------------------
class FrenetOptimalPlanner(object):
    def __init__(self, planner_settings: FrenetOptimalPlannerSettings, ego_vehicle: Vehicle, scenario=None,
                 device: int = 0, engine: FissEngine = None):
        self.settings = planner_settings
        self.vehicle = ego_vehicle
        self.cost_function = CostFunction("WX1")
        self.cubic_spline = None
        self.best_traj = None
        self.all_trajs = []
        self.stats = Stats()
        # device side
        self._device = device
        self._engine = engine
        self._obstacle_key = None
        self._obstacle_token = None
        self._obstacle_ref = None
        self._spline_table = None
        self._spline_token = None
        self._lattice_key = None
        self._lattice = None
        self._grid = None
        self._params_key = None
        self._params_proto = None
        self._out = None
        self._out_key = None
        self._out_grid = None
        self.check_curvature = False  # optional third mask bit; off = reference behaviour (:145-150)

    # -- device plumbing -------------------------------------------------------------------------
    @property
    def engine(self) -> FissEngine:
        if self._engine is None:
            self._engine = FissEngine(self._device)
        return self._engine

    def _install_spline(self, table):
        self.engine.set_spline(table)
        self._spline_table = table
        self._spline_token = self.engine.spline_token

    def _bundle(self, ego6, end, prm, cost, flags, idx3=None) -> "CandidateBundle":
        return CandidateBundle(self.engine, ego6, end, prm, cost, flags, idx3=idx3, spline_table=self._spline_table,
                               spline_token=self._spline_token)

    def invalidate_obstacles(self):
        """Force the next ``plan()`` to re-marshal the obstacle list (needed only when a prediction object was
        modified IN PLACE: a different list, or a list whose elements were replaced, is detected by itself)."""
        self._obstacle_key = None

    @staticmethod
    def _obstacle_identity(obstacles):
        if isinstance(obstacles, (list, tuple)):
            return (id(obstacles), tuple(id(ob) for ob in obstacles))
        return (id(obstacles), getattr(obstacles, "version", 0))

    def _upload_obstacles(self, obstacles):
        """Scene tables before a cycle.  Predictions are immutable objects, so the list is re-marshalled only when it
        is a different list or holds different elements (the reference re-reads ``state_at_time`` every cycle, :185-189);
        the engine's tokens tell when somebody else -- another planner sharing the engine, a lazy candidate bundle --
        has replaced the tables since, in which case this planner's own are installed again."""
        eng = self.engine
        if self._spline_table is not None and eng.spline_token != self._spline_token:
            self._install_spline(self._spline_table)
        key = self._obstacle_identity(obstacles)
        if key != self._obstacle_key or eng.obstacle_token != self._obstacle_token:
            if hasattr(obstacles, "upload_to"):       # device-ready wire formats (waymo_interface.WaymoObstacles)
                obstacles.upload_to(eng)
            else:
                tab = marshal_obstacles(obstacles)
                if len(tab) == 0:
                    eng.set_obstacles(None, np.zeros((0, 2)), None, 0)
                else:
                    eng.set_obstacles(tab.xyth, tab.lw, tab.valid, tab.final_time_step)
            self._obstacle_key = key
            self._obstacle_token = eng.obstacle_token
            # keep the list AND its elements alive, so that their ids cannot be reused by new objects
            self._obstacle_ref = (obstacles, list(obstacles) if isinstance(obstacles, (list, tuple)) else None)

    def _end_states(self) -> np.ndarray:
        st = self.settings
        key = (st.max_road_width, self.vehicle.w, st.num_width, st.min_t, st.max_t, st.num_t, st.lowest_speed,
               st.highest_speed, st.num_speed, st.tick_t)
        if key != self._lattice_key:
            self._grid = fop_grid(st, self.vehicle.w)
            self._lattice = self._grid.table()
            self._lattice_key = key
        return self._lattice

    def _lattice_grid(self):
        """The same lattice as a ``LatticeGrid`` (axes + numbering) for the lattice kernel."""
        self._end_states()
        return self._grid

    def _params(self, time_step_now: int, collide_all: bool = False):
        """``struct fiss_params`` of this cycle.  Building the 16-field ctypes struct costs several microseconds -- a tenth
        of a plan cycle --, so it is rebuilt only when a setting, a vehicle limit or a switch changed; a cycle gets its
        own copy (the candidate bundles keep theirs) with ``time_step_now`` filled in."""
        st, veh = self.settings, self.vehicle
        key = (st.tick_t, st.highest_speed, veh.max_speed, veh.max_accel, getattr(veh, "max_curvature", None), veh.l, veh.w,
               self.check_curvature, collide_all)
        if key != self._params_key:
            self._params_proto = make_params(st, veh, self.cost_function.as_device_weights(), 0, check_res=2,
                                             check_curvature=self.check_curvature, collide_all=collide_all)
            self._params_key = key
        prm = type(self._params_proto).from_buffer_copy(self._params_proto)
        prm.time_step_now = int(time_step_now)
        return prm

    def _plan_outputs(self, grid, batch: int = 1, want_records: bool = True) -> dict:
        """The planner's reusable output set for ``plan_grid`` (winners, records, cost / flags volume): allocating six
        arrays and taking their addresses every cycle cost more host time than the call itself."""
        if self._out is None or self._out_key != (id(grid), batch, want_records):
            self._out = self.engine.alloc_plan_outputs(batch, grid, want_records=want_records, want_volume=True, pinned=False)
            self._out_key = (id(grid), batch, want_records)
            self._out_grid = grid           # keeps id(grid) alive
        return self._out

    def _trajectory_from_record(self, rec, meta, cost) -> FrenetTrajectory:
        return FrenetTrajectory().fill_from_device_record(rec, int(meta[0]), int(meta[1]), float(cost))

    # -- reference API ---------------------------------------------------------------------------
    def generate_frenet_frame(self, centerline_pts: np.ndarray, fit: str = "host"):
        """Reference-line set-up (frenet_optimal_planner.py:272-278): spline through the centre line, table upload,
        and the 0.1 m polyline ``[m, 4] = (x, y, yaw, curvature)`` that ``FrenetState.from_state`` consumes.

        ``fit="host"`` (default): the reference's own dense float64 solve on the host -- coefficients identical to
        the reference's, which the bit-exact mask parity rests on.  ``fit="device"``: the Thomas-recurrence fit and
        the resampling run on the GPU (``fiss_fit_splines_host`` / ``fiss_frame_samples_host``); equal to the host
        path to rounding (~1e-15 relative)."""
        if fit == "device":
            table = self.engine.fit_splines(np.asarray(centerline_pts, dtype=np.float64)[:, :2], install=0)[0]
            self._spline_table, self._spline_token = table, self.engine.spline_token
            self._obstacle_key = None
            self.cubic_spline = CubicSpline2D.from_device_table(table)
            return self.cubic_spline, self.engine.frame_samples(self.cubic_spline.s[-1], 0.1)
        if fit != "host":
            raise ValueError("fit must be 'host' or 'device'")
        self.cubic_spline = CubicSpline2D(centerline_pts[:, 0], centerline_pts[:, 1])
        self._install_spline(self.cubic_spline.device_table())
        self._obstacle_key = None
        s = np.arange(0, self.cubic_spline.s[-1], 0.1)
        ref_xy = [self.cubic_spline.calc_position(i_s) for i_s in s]
        ref_yaw = [self.cubic_spline.calc_yaw(i_s) for i_s in s]
        ref_rk = [self.cubic_spline.calc_curvature(i_s) for i_s in s]
        return self.cubic_spline, np.column_stack((ref_xy, ref_yaw, ref_rk))

    def plan(self, frenet_state: FrenetState, max_target_speed: float, obstacles: list, time_step_now: int = 0) -> FrenetTrajectory:
        self.stats = Stats()
        self.settings.highest_speed = max_target_speed      # mutates settings, as the reference does (:250)
        self._upload_obstacles(obstacles)
        end = self._end_states()
        prm = self._params(time_step_now)
        ego6 = frenet_state.as_ego6() if hasattr(frenet_state, "as_ego6") else np.array(
            [frenet_state.s, frenet_state.s_d, frenet_state.s_dd, frenet_state.d, frenet_state.d_d, frenet_state.d_dd])
        grid = self._lattice_grid()
        out = self.engine.plan_grid(ego6, grid, prm, want_records=True, want_volume=True, out=self._plan_outputs(grid))

        n_cand = len(end)
        self.stats.num_trajs_generated = n_cand
        self.stats.num_trajs_validated = n_cand
        self.stats.num_collison_checks = n_cand
        self.all_trajs.append(self._bundle(ego6, end, prm, out["cost"][0], out["flags"][0]))

        best = int(out["best_idx"][0])
        if best >= 0:
            traj = self._trajectory_from_record(out["records"][0], out["meta"][0], out["best_cost"][0])
            traj.lattice_index = best
            self.best_traj = traj
        # nothing feasible: the previous cycle's best_traj is returned unchanged, like the reference (:263-270)
        return self.best_traj
