"""ORACLE (test infrastructure, not product code): CPU restatement of the reference hot path.

A plain Python / NumPy float64 restatement of SS47816/fiss_plus_planner's lattice sampling ->
quintic / quartic evaluation -> cost -> Frenet->Cartesian -> constraint and collision masks ->
argmin, plus the four planners' search loops around it.  Each function names the reference
file:line it follows (paths relative to /root/reference).  It keeps the reference's scalar
evaluation order (Python ``**`` powers, sequential ``sum``) so that FP64 results are the
reference's to the last bit wherever NumPy's LAPACK is the same; it deliberately does NOT keep
the reference's ``copy.deepcopy`` of the lateral arrays (frenet_optimal_planner.py:90), which is
an implementation cost, not part of the algorithm -- the timed "port" baseline is therefore a
little FASTER than the real reference.

PINNING: tests/test_oracle_golden.py checks this module against tests/golden/*.npz, which were
produced by executing the reference's own modules in the build container
(tests/golden/make_golden.py).  Polynomials, cost, spline, global path, constraint mask,
argmin / tie rules and the FOP+/FISS/FISS+ searches are pinned that way.  The collision
predicate is shapely/GEOS in the reference (third-party, absent): here and in the goldens it is
oracle/sat_geometry.py -- COLLISION PARITY UNPINNED at the GEOS boundary.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (fiss_plus_planner_b200/) never does.
"""
from __future__ import annotations

import bisect
import heapq
import math
import time

import numpy as np

from oracle import sat_geometry as sat

# cost weights, CostFunction("WX1") -- planners/common/cost/cost_function.py:6-12
W_V, W_A, W_J, W_LC = 1, 0.1, 0.1, 10


# ----------------------------------------------------------------------------------- settings
class Settings:
    """FrenetOptimalPlannerSettings / FissPlannerSettings / FissPlusPlannerSettings
    (frenet_optimal_planner.py:38-56, fiss_planner.py:13-18, fiss_plus_planner.py:15-22)."""

    def __init__(self, num_width=5, num_speed=5, num_t=5, refine_iters=3):
        self.tick_t = 0.1
        self.max_road_width = 3.5
        self.num_width = num_width
        self.highest_speed = 13.4112
        self.lowest_speed = 0.0
        self.num_speed = num_speed
        self.min_t = 8.0
        self.max_t = 10.0
        self.num_t = num_t
        self.w_heuristic = 10.0
        self.refine_trajectory = True
        self.max_refine_iters = refine_iters
        self.decaying_factor = 0.5


class Stats:
    """frenet_optimal_planner.py:15-36."""

    def __init__(self):
        self.num_iter = 0
        self.num_trajs_generated = 0
        self.num_trajs_validated = 0
        self.num_collison_checks = 0

    def as_tuple(self):
        return (self.num_iter, self.num_trajs_generated, self.num_trajs_validated, self.num_collison_checks)


# ----------------------------------------------------------------------------------- spline
class Spline1D:
    """Natural cubic spline, dense solve -- planners/common/geometry/cubic_spline.py:19-43,118-142."""

    def __init__(self, x, y):
        h = np.diff(x)
        n = len(x)
        self.x = list(x)
        self.a = [v for v in y]
        A = np.zeros((n, n))
        A[0, 0] = 1.0
        for i in range(n - 1):
            if i != n - 2:
                A[i + 1, i + 1] = 2.0 * (h[i] + h[i + 1])
            A[i + 1, i] = h[i]
            A[i, i + 1] = h[i]
        A[0, 1] = 0.0
        A[n - 1, n - 2] = 0.0
        A[n - 1, n - 1] = 1.0
        B = np.zeros(n)
        for i in range(n - 2):
            B[i + 1] = 3.0 * (self.a[i + 2] - self.a[i + 1]) / h[i + 1] - 3.0 * (self.a[i + 1] - self.a[i]) / h[i]
        self.c = np.linalg.solve(A, B)
        self.b, self.d = [], []
        for i in range(n - 1):
            self.d.append((self.c[i + 1] - self.c[i]) / (3.0 * h[i]))
            self.b.append(1.0 / h[i] * (self.a[i + 1] - self.a[i]) - h[i] / 3.0 * (2.0 * self.c[i] + self.c[i + 1]))

    def position(self, x):
        """cubic_spline.py:45-66 (None outside [x0, xK]; IndexError at x == xK is the reference's)."""
        if x < self.x[0] or x > self.x[-1]:
            return None
        i = bisect.bisect(self.x, x) - 1
        dx = x - self.x[i]
        return self.a[i] + self.b[i] * dx + self.c[i] * dx ** 2.0 + self.d[i] * dx ** 3.0

    def first_derivative(self, x):
        """cubic_spline.py:68-88."""
        if x < self.x[0] or x > self.x[-1]:
            return None
        i = bisect.bisect(self.x, x) - 1
        dx = x - self.x[i]
        return self.b[i] + 2.0 * self.c[i] * dx + 3.0 * self.d[i] * dx ** 2.0


class Spline2D:
    """Arc-length parametrised 2-D spline -- cubic_spline.py:157-168,170-190,214-232."""

    def __init__(self, x, y):
        ds = np.hypot(np.diff(x), np.diff(y))
        s = [0]
        s.extend(np.cumsum(ds))
        self.s = s
        self.sx = Spline1D(s, x)
        self.sy = Spline1D(s, y)

    def position(self, s):
        return self.sx.position(s), self.sy.position(s)

    def yaw(self, s):
        return math.atan2(self.sy.first_derivative(s), self.sx.first_derivative(s))

    def table(self) -> np.ndarray:
        """[9, K]: knots, ax..dx, ay..dy (b, d zero-padded) -- the layout the device consumes."""
        k = len(self.s)
        tab = np.zeros((9, k))
        tab[0] = self.s
        for r, sp in ((1, self.sx), (5, self.sy)):
            tab[r] = sp.a
            tab[r + 1, :k - 1] = sp.b
            tab[r + 2] = sp.c
            tab[r + 3, :k - 1] = sp.d
        return tab


# ----------------------------------------------------------------------------------- polynomials
def quartic_coeffs(xs, vxs, axs, vxe, axe, T):
    """planners/common/geometry/polynomial.py:5-19 (2x2 LAPACK solve)."""
    a0, a1, a2 = xs, vxs, axs / 2.0
    A = np.array([[3 * T ** 2, 4 * T ** 3], [6 * T, 12 * T ** 2]])
    b = np.array([vxe - a1 - 2 * a2 * T, axe - 2 * a2])
    a3, a4 = np.linalg.solve(A, b)
    return a0, a1, a2, a3, a4


def quintic_coeffs(xs, vxs, axs, xe, vxe, axe, T):
    """planners/common/geometry/polynomial.py:45-62 (3x3 LAPACK solve)."""
    a0, a1, a2 = xs, vxs, axs / 2.0
    A = np.array([[T ** 3, T ** 4, T ** 5], [3 * T ** 2, 4 * T ** 3, 5 * T ** 4], [6 * T, 12 * T ** 2, 20 * T ** 3]])
    b = np.array([xe - a0 - a1 * T - a2 * T ** 2, vxe - a1 - 2 * a2 * T, axe - 2 * a2])
    a3, a4, a5 = np.linalg.solve(A, b)
    return a0, a1, a2, a3, a4, a5


def quartic_eval(c, ts):
    """polynomial.py:21-41: power-form evaluation, one scalar at a time."""
    a0, a1, a2, a3, a4 = c
    p = [a0 + a1 * t + a2 * t ** 2 + a3 * t ** 3 + a4 * t ** 4 for t in ts]
    p1 = [a1 + 2 * a2 * t + 3 * a3 * t ** 2 + 4 * a4 * t ** 3 for t in ts]
    p2 = [2 * a2 + 6 * a3 * t + 12 * a4 * t ** 2 for t in ts]
    p3 = [6 * a3 + 24 * a4 * t for t in ts]
    return p, p1, p2, p3


def quintic_eval(c, ts):
    """polynomial.py:64-84."""
    a0, a1, a2, a3, a4, a5 = c
    p = [a0 + a1 * t + a2 * t ** 2 + a3 * t ** 3 + a4 * t ** 4 + a5 * t ** 5 for t in ts]
    p1 = [a1 + 2 * a2 * t + 3 * a3 * t ** 2 + 4 * a4 * t ** 3 + 5 * a5 * t ** 4 for t in ts]
    p2 = [2 * a2 + 6 * a3 * t + 12 * a4 * t ** 2 + 20 * a5 * t ** 3 for t in ts]
    p3 = [6 * a3 + 24 * a4 * t + 60 * a5 * t ** 2 for t in ts]
    return p, p1, p2, p3


# ----------------------------------------------------------------------------------- candidate
class Traj:
    """The fields of FrenetTrajectory the path touches (planners/common/scenario/frenet.py:113-148)."""
    __slots__ = ("idx", "end", "is_generated", "cost_est", "cost_heu", "cost_final", "t",
                 "s", "s_d", "s_dd", "s_ddd", "d", "d_d", "d_dd", "d_ddd",
                 "x", "y", "yaw", "ds", "c", "c_d", "c_dd", "seq")

    def __init__(self):
        self.idx = np.array([-1, -1, -1])
        self.end = None            # (d, v, T)
        self.is_generated = False
        self.cost_est = 0.0
        self.cost_heu = 0.0
        self.cost_final = 0.0
        self.t = []
        self.s, self.s_d, self.s_dd, self.s_ddd = [], [], [], []
        self.d, self.d_d, self.d_dd, self.d_ddd = [], [], [], []
        self.x, self.y, self.yaw, self.ds, self.c, self.c_d, self.c_dd = [], [], [], [], [], [], []
        self.seq = -1

    def __lt__(self, other):      # frenet.py:156 -- cost-only ordering
        return self.cost_final < other.cost_final


def cost_total(tr: Traj, target_speed: float) -> float:
    """planners/common/cost/cost_function.py:41-50 (sequential Python sums of NumPy squares)."""
    cost_time = 10.0 - tr.t[-1]
    cost_speed = W_V * sum(np.power(np.subtract(tr.s_d, target_speed), 2))
    cost_accel = W_A * sum(np.power(tr.s_dd, 2)) + W_A * sum(np.power(tr.d_dd, 2))
    cost_jerk = W_J * sum(np.power(tr.s_ddd, 2)) + W_J * sum(np.power(tr.d_ddd, 2))
    cost_offset = W_LC * sum(np.power(tr.d, 2))
    return (cost_time + 0.0 + cost_speed + cost_accel + cost_jerk + cost_offset) / len(tr.t)


def generate(tr: Traj, ego6, d_end, v_end, T, tick, target_speed) -> Traj:
    """Inner body shared by calc_frenet_paths (frenet_optimal_planner.py:79-99),
    generate_trajectory (fiss_planner.py:110-128) and generate_trajectory_by_end_state
    (fiss_plus_planner.py:181-196): time grid, lateral quintic to (d_end, 0, 0), longitudinal
    quartic to (v_end, 0), cost."""
    s0, s_d0, s_dd0, d0, d_d0, d_dd0 = ego6
    tr.t = [t for t in np.arange(0.0, T, tick)]
    tr.d, tr.d_d, tr.d_dd, tr.d_ddd = quintic_eval(quintic_coeffs(d0, d_d0, d_dd0, d_end, 0.0, 0.0, T), tr.t)
    tr.s, tr.s_d, tr.s_dd, tr.s_ddd = quartic_eval(quartic_coeffs(s0, s_d0, s_dd0, v_end, 0.0, T), tr.t)
    tr.cost_final = cost_total(tr, target_speed)
    tr.end = (d_end, v_end, T)
    tr.is_generated = True
    return tr


def to_global(tr: Traj, spline: Spline2D, tick: float) -> Traj:
    """calc_global_paths for one candidate (frenet_optimal_planner.py:108-136)."""
    xs, ys = [], []
    for i in range(len(tr.s)):
        ix, iy = spline.position(tr.s[i])
        if ix is None:
            break
        i_yaw = spline.yaw(tr.s[i])
        di = tr.d[i]
        xs.append(ix + di * math.cos(i_yaw + math.pi / 2.0))
        ys.append(iy + di * math.sin(i_yaw + math.pi / 2.0))
    tr.x, tr.y = xs, ys
    if len(xs) >= 2:
        tr.x = np.array(xs)
        tr.y = np.array(ys)
        x_d = np.diff(tr.x)
        y_d = np.diff(tr.y)
        yaw = np.arctan2(y_d, x_d)
        tr.ds = np.hypot(x_d, y_d)
        tr.yaw = np.append(yaw, yaw[-1])
        tr.c = np.divide(np.diff(tr.yaw), tr.ds)
        tr.c_d = np.divide(np.diff(tr.c), tick)
        tr.c_dd = np.divide(np.diff(tr.c_d), tick)
    return tr


def passes_constraints(tr: Traj, max_speed: float, max_accel: float, max_curvature: float | None = None) -> bool:
    """check_constraints (frenet_optimal_planner.py:140-160): speed (signed) and |accel| only.  ``max_curvature`` (default
    off) switches on the first of the three checks the reference carries commented out (:145-146,
    ``any([abs(c) > self.vehicle.max_curvature for c in traj.c])``) -- north_star's optional curvature mask."""
    if max_curvature is not None and any([abs(c) > max_curvature for c in tr.c]):
        return False
    if any([v > max_speed for v in tr.s_d]):
        return False
    if any([abs(a) > max_accel for a in tr.s_dd]):
        return False
    return True


def curvature_ok(tr: Traj, max_curvature: float) -> bool:
    """The curvature check alone (frenet_optimal_planner.py:145-146, commented out in the reference)."""
    return not any([abs(c) > max_curvature for c in tr.c])


class ObstacleTable:
    """Dense obstacle predictions: xyth [M,T,3], lw [M,2], valid [M,T], final_time_step.
    Rings are placed once (they do not depend on the candidate): construct_polygon on
    ``obstacle.obstacle_shape.shapely_object`` (frenet_optimal_planner.py:189)."""

    def __init__(self, xyth, lw, valid, final_time_step):
        self.xyth, self.lw, self.valid = np.asarray(xyth), np.asarray(lw), np.asarray(valid, dtype=bool)
        self.final_time_step = int(final_time_step)
        self.m = len(self.lw)
        self._rings = {}

    def rings_at(self, t: int):
        """[m_valid, 4, 2] placed rings of the obstacles that have a state at step ``t``."""
        if t not in self._rings:
            out = []
            if 0 <= t < self.valid.shape[1]:
                for j in range(self.m):
                    if self.valid[j, t]:
                        out.append(sat.place(sat.obstacle_ring(*self.lw[j]), *self.xyth[j, t]))
            self._rings[t] = np.array(out).reshape(-1, 4, 2)
        return self._rings[t]


def waymo_obstacle_table(waymo_trajs: np.ndarray, waymo_traj_masks: np.ndarray) -> ObstacleTable | None:
    """convert_waymo_obstacle_to_cr (planners/waymo_interface/waymo_interface.py:24-76) restated onto the dense table:
    per agent, the rectangle from row 0 (:33-34), the initial state at step 0 whatever the mask says (:36-40), then states
    1, 2, ... until the first masked step (``break``, :45-54); an agent whose state list is empty is not appended (:58), so
    it has no state at any step and the later agents move up.  What the planner then reads off the list:
    ``obstacles[0].prediction.final_time_step`` = the time step of the first kept agent's last state
    (frenet_optimal_planner.py:173) and ``state_at_time(t)`` = initial state at 0, trajectory state in [1, t_end], else
    ``None`` (:187-189; commonroad-io behaviour, restated from memory -- SURVEY 8(f) f-2).  Returns ``None`` for an empty
    list (``has_collision`` then answers False, :170-171).  Values are widened to float64, as shapely does."""
    n_obs, n_t, _ = waymo_trajs.shape
    rows, sizes, valids = [], [], []
    for i in range(n_obs):
        traj = waymo_trajs[i]
        if traj.shape[0] <= 0:
            continue
        states = []
        for t in range(1, n_t):
            if waymo_traj_masks[i, t]:
                states.append(t)
            else:
                break
        if states:
            valid = np.zeros(n_t, dtype=bool)
            valid[0] = True
            valid[states] = True
            xyth = np.zeros((n_t, 3))
            for t in np.flatnonzero(valid):
                xyth[t] = (float(traj[t, 0]), float(traj[t, 1]), float(traj[t, 6]))
            rows.append(xyth)
            sizes.append((float(traj[0, 3]), float(traj[0, 4])))
            valids.append(valid)
    if not rows:
        return None
    final_time_step = int(np.flatnonzero(valids[0])[-1])
    return ObstacleTable(np.array(rows), np.array(sizes), np.array(valids), final_time_step)


def has_collision(tr: Traj, obs: ObstacleTable | None, ego_ring: np.ndarray, now: int, check_res: int = 2) -> bool:
    """has_collision (frenet_optimal_planner.py:168-195): even steps below
    min(n', final_time_step - now); ego rectangle at (x_i, y_i, yaw_i) against every obstacle
    with a state at i + now; an exception while building the ego polygon counts as a collision."""
    if obs is None or obs.m <= 0:
        return False
    t_step_max = min(len(tr.x), obs.final_time_step - now)
    for i in range(t_step_max):
        if i % check_res == 0:
            try:
                ego = sat.place(ego_ring, tr.x[i], tr.y[i], tr.yaw[i])
            except Exception:
                return True
            rings = obs.rings_at(i + now)
            if len(rings) and sat.sat_closed_many(ego, rings).any():
                return True
    return False


def arange_len(T: float, tick: float) -> int:
    """len(np.arange(0.0, T, tick)) -- the step count n of a candidate."""
    return len(np.arange(0.0, T, tick))


# ----------------------------------------------------------------------------------- planners
class FopOracle:
    """FrenetOptimalPlanner (frenet_optimal_planner.py:58-278)."""

    def __init__(self, settings: Settings, ego_l, ego_w, max_speed, max_accel):
        self.settings = settings
        self.ego_l, self.ego_w = ego_l, ego_w
        self.max_speed, self.max_accel = max_speed, max_accel
        self.ego_ring = sat.ego_ring(ego_l, ego_w)
        self.spline = None
        self.best_traj = None
        self.stats = Stats()
        self.last_all = []

    def generate_frenet_frame(self, centerline_pts):
        self.spline = Spline2D(centerline_pts[:, 0], centerline_pts[:, 1])
        return self.spline

    def lattice(self):
        """(d, T, v) enumeration of calc_frenet_paths (:72-78,89): d outer, T middle, v inner."""
        st = self.settings
        sw = st.max_road_width - self.ego_w
        out = []
        for di in np.linspace(-sw / 2, sw / 2, st.num_width):
            for Ti in np.linspace(st.min_t, st.max_t, st.num_t):
                for tv in np.linspace(st.lowest_speed, st.highest_speed, st.num_speed):
                    out.append((di, tv, Ti))
        return out

    def sample_all(self, ego6):
        st = self.settings
        trajs = []
        for seq, (di, tv, Ti) in enumerate(self.lattice()):
            tr = generate(Traj(), ego6, di, tv, Ti, st.tick_t, st.highest_speed)
            tr.seq = seq
            trajs.append(tr)
        return trajs

    def plan(self, ego6, max_target_speed, obs: ObstacleTable | None, now=0):
        """plan (:247-270): whole lattice, masks, last minimal-cost survivor; stale best kept."""
        self.stats = Stats()
        self.settings.highest_speed = max_target_speed
        trajs = [to_global(tr, self.spline, self.settings.tick_t) for tr in self.sample_all(ego6)]
        self.last_all = trajs
        self.stats.num_trajs_generated = self.stats.num_trajs_validated = self.stats.num_collison_checks = len(trajs)
        self.last_constraint_ok = [passes_constraints(tr, self.max_speed, self.max_accel) for tr in trajs]
        survivors = [tr for tr, ok in zip(trajs, self.last_constraint_ok) if ok]
        self.last_collision = {tr.seq: has_collision(tr, obs, self.ego_ring, now) for tr in survivors}
        survivors = [tr for tr in survivors if not self.last_collision[tr.seq]]
        min_cost = float("inf")
        for tr in survivors:
            if min_cost >= tr.cost_final:
                min_cost = tr.cost_final
                self.best_traj = tr
        return self.best_traj


class FopPlusOracle(FopOracle):
    """FopPlusPlanner.plan (fop_plus_planner.py:16-40): best-first lazy validation."""

    def plan(self, ego6, max_target_speed, obs, now=0):
        self.stats = Stats()
        self.settings.highest_speed = max_target_speed
        trajs = [to_global(tr, self.spline, self.settings.tick_t) for tr in self.sample_all(ego6)]
        self.last_all = trajs
        self.stats.num_trajs_generated = len(trajs)
        heap = []
        for tr in trajs:
            heapq.heappush(heap, tr)          # queue.PriorityQueue is heapq underneath
        while heap:
            self.stats.num_iter += 1
            cand = heapq.heappop(heap)
            ok = passes_constraints(cand, self.max_speed, self.max_accel)
            self.stats.num_trajs_validated += 1
            safe = ok and not has_collision(cand, obs, self.ego_ring, now)
            self.stats.num_collison_checks += 1
            if safe:
                self.best_traj = cand
                return cand
        return None


class FissOracle(FopOracle):
    """FissPlanner (fiss_planner.py:20-269)."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.sampling_res = np.empty(3)
        self.sampling_min = np.empty(3)
        self.sampling_max = np.empty(3)
        self.prev_best_idx = None
        self.grid = []
        self.heap = []
        self.ego6 = None
        self.generated_log = []

    def sample_end_states(self):
        """sample_end_frenet_states (fiss_planner.py:33-99): [i_d][j_v][k_t] placeholders with
        cost_est = lateral + horizon + speed terms (+ history heuristic)."""
        st = self.settings
        max_sqr = np.power(st.num_width, 2) + np.power(st.num_speed, 2) + np.power(st.num_t, 2)
        sw = st.max_road_width - self.ego_w + 0.3
        left, right = -sw / 2, sw / 2
        self.sampling_min[0], self.sampling_max[0] = left, right
        d_samples, self.sampling_res[0] = np.linspace(left, right, st.num_width, retstep=True)
        grid = []
        for i, d in enumerate(d_samples):
            plane = []
            lat_norm = max(np.power(left, 2), np.power(right, 2))
            est_lat = np.power(d, 2) / lat_norm
            self.sampling_min[1], self.sampling_max[1] = st.lowest_speed, st.highest_speed
            v_samples, self.sampling_res[1] = np.linspace(st.lowest_speed, st.highest_speed, st.num_speed, retstep=True)
            for j, v in enumerate(v_samples):
                row = []
                est_speed = np.power(st.highest_speed - v, 2) / np.power(st.highest_speed - st.lowest_speed, 2)
                self.sampling_min[2], self.sampling_max[2] = st.min_t, st.max_t
                t_samples, self.sampling_res[2] = np.linspace(st.min_t, st.max_t, st.num_t, retstep=True)
                for k, t in enumerate(t_samples):
                    est_time = 1.0 - (t - st.min_t) / (st.max_t - st.min_t)
                    est = est_lat + est_time + est_speed
                    if self.prev_best_idx is not None:
                        p = self.prev_best_idx
                        sqr = np.power(i - p[0], 2) + np.power(j - p[1], 2) + np.power(k - p[2], 2)
                        heu = st.w_heuristic * sqr / max_sqr
                    else:
                        heu = 0.0
                    tr = Traj()
                    tr.idx = np.array([i, j, k])
                    tr.end = (d, v, t)
                    tr.cost_heu = heu
                    tr.cost_est = est + heu
                    row.append(tr)
                plane.append(row)
            grid.append(plane)
        self.grid = grid
        self.sizes = np.array([len(grid), len(grid[0]), len(grid[0][0])])
        return grid

    def generate_at(self, idx):
        """generate_trajectory (fiss_planner.py:101-138)."""
        tr = self.grid[idx[0]][idx[1]][idx[2]]
        if tr.is_generated:
            return False, tr.cost_final
        self.stats.num_trajs_generated += 1
        tr.idx = idx
        d, v, T = tr.end
        generate(tr, self.ego6, d, v, T, self.settings.tick_t, self.settings.highest_speed)
        self.generated_log.append(tr)
        heapq.heappush(self.heap, (tr.cost_final, tr.idx))
        return True, tr.cost_final

    def initial_guess(self):
        """find_initial_guess (fiss_planner.py:140-150): min cost_est, '<=' so the last one wins."""
        best, lo = None, float("inf")
        for plane in self.grid:
            for row in plane:
                for tr in row:
                    if not tr.is_generated and tr.cost_est <= lo:
                        lo = tr.cost_est
                        best = tr.idx
        return best

    def gradients(self, idx):
        """find_gradients (fiss_planner.py:152-172)."""
        _, centre = self.generate_at(idx)
        g = np.empty(3)
        for dim in range(3):
            nb = idx.copy()
            if idx[dim] < self.sizes[dim] - 1:
                nb[dim] += 1
                _, cost = self.generate_at(nb)
                g[dim] = cost - centre
                if g[dim] >= 0 and idx[dim] == 0:
                    g[dim] = 0.0
            else:
                nb[dim] -= 1
                _, cost = self.generate_at(nb)
                g[dim] = centre - cost
                if g[dim] <= 0 and idx[dim] == self.sizes[dim] - 1:
                    g[dim] = 0.0
        return g

    def explore_next(self, idx):
        """explore_next_sample (fiss_planner.py:174-188)."""
        if self.grid[idx[0]][idx[1]][idx[2]].is_generated:
            return True, idx
        nxt = idx.copy()
        g = self.gradients(idx)
        for dim in range(3):
            nxt[dim] += -1 if g[dim] > 0.0 else +1
        return False, np.clip(nxt, 0, self.sizes - 1)

    def _validate(self, cand, obs, now):
        """shared validation tail (fiss_planner.py:233-260): global path, constraints, collision."""
        self.stats.num_trajs_validated += 1
        to_global(cand, self.spline, self.settings.tick_t)
        if not passes_constraints(cand, self.max_speed, self.max_accel):
            return False
        hit = has_collision(cand, obs, self.ego_ring, now)
        self.stats.num_collison_checks += 1
        return not hit

    def _reset(self, ego6, max_target_speed):
        self.stats = Stats()
        self.settings.highest_speed = max_target_speed
        self.ego6 = ego6
        self.heap = []
        self.best_traj = None
        self.generated_log = []
        self.sample_end_states()

    def plan(self, ego6, max_target_speed, obs, now=0):
        """plan (fiss_planner.py:190-269)."""
        self._reset(ego6, max_target_speed)
        found = False
        while not found:
            self.stats.num_iter += 1
            if not self.heap:
                best_idx = self.initial_guess()
                if best_idx is None:
                    break
            else:
                best_idx = self.heap[0][1]
            converged = False
            while not converged:
                converged, best_idx = self.explore_next(best_idx)
            if self.heap:
                _, idx = heapq.heappop(self.heap)
                cand = self.grid[idx[0]][idx[1]][idx[2]]
                if self._validate(cand, obs, now):
                    found = True
                    self.best_traj = cand
                    self.prev_best_idx = cand.idx
                    break
            else:
                break
        return self.best_traj


class FissPlusOracle(FissOracle):
    """FissPlusPlanner (fiss_plus_planner.py:24-326)."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.frontier = []
        self.refined = []

    def explore_neighbors(self, idx):
        """explore_neighbors (fiss_plus_planner.py:30-59)."""
        _, centre = self.generate_at(idx)
        lo = centre
        best = idx.copy()
        is_min = True
        for dim in range(3):
            for step, exists in ((-1, idx[dim] >= 1), (+1, idx[dim] < self.sizes[dim] - 1)):
                if not exists:
                    continue
                nb = idx.copy()
                nb[dim] += step
                is_new, cost = self.generate_at(nb)
                if is_new and cost <= centre:
                    heapq.heappush(self.frontier, (cost, nb))
                if cost <= lo:
                    lo = cost
                    best = nb
                    is_min = False
        return is_min, best

    def generate_by_end_state(self, d, v, T):
        """generate_trajectory_by_end_state (fiss_plus_planner.py:172-205)."""
        self.stats.num_trajs_generated += 1
        tr = generate(Traj(), self.ego6, d, v, T, self.settings.tick_t, self.settings.highest_speed)
        heapq.heappush(self.refined, tr)
        self.refined_log.append(tr)
        return tr.cost_final

    def gradient_descent(self, x, res, decay):
        """gradient_decent (fiss_plus_planner.py:207-277): 6 clipped neighbours + one normalised step."""
        dJ = np.empty(3)
        dx = np.empty(3)
        for dim in range(3):
            xl = x.copy()
            xl[dim] -= res[dim]
            xl = np.clip(xl, self.sampling_min, self.sampling_max)
            Jl = self.generate_by_end_state(xl[0], xl[1], xl[2])
            xr = x.copy()
            xr[dim] += res[dim]
            xr = np.clip(xr, self.sampling_min, self.sampling_max)
            Jr = self.generate_by_end_state(xr[0], xr[1], xr[2])
            dJ[dim] = Jr - Jl
            dx[dim] = xr[dim] - xl[dim]
        grad = dJ / dx
        res *= decay                      # in place: aliases self.sampling_res (:261,282)
        x_new = np.clip(x - res * grad / np.linalg.norm(grad), self.sampling_min, self.sampling_max)
        J_new = self.generate_by_end_state(x_new[0], x_new[1], x_new[2])
        return True, J_new, x_new, res

    def refine(self, best, obs, now):
        """refine_solution (fiss_plus_planner.py:279-326).  The reference also breaks out of the
        refinement rounds on wall-clock time (:296-299; budget = settings.time_limit minus the coarse
        search time, :153-156); the oracle models an unlimited budget (goldens use time_limit=1e9)."""
        res = self.sampling_res
        x = np.array([best.end[0], best.end[1], best.end[2]])
        for _ in range(self.settings.max_refine_iters):
            ok, _, x, res = self.gradient_descent(x, res, self.settings.decaying_factor)
            if not ok:
                break
        while self.refined:
            cand = heapq.heappop(self.refined)
            if cand.cost_final > best.cost_final:
                break
            if self._validate(cand, obs, now):
                return cand
        return None

    def plan(self, ego6, max_target_speed, obs, now=0):
        """plan (fiss_plus_planner.py:61-170)."""
        self._reset(ego6, max_target_speed)
        self.frontier, self.refined, self.refined_log = [], [], []
        found = False
        while not found:
            self.stats.num_iter += 1
            if not self.heap:
                best_idx = self.initial_guess()
                if best_idx is None:
                    break
            else:
                best_idx = self.heap[0][1]
            converged = False
            while not converged:
                _, best_idx = self.explore_neighbors(best_idx)
                if not self.frontier:
                    converged = True
                else:
                    _, best_idx = heapq.heappop(self.frontier)
            if self.heap:
                _, idx = heapq.heappop(self.heap)
                cand = self.grid[idx[0]][idx[1]][idx[2]]
                if self._validate(cand, obs, now):
                    found = True
                    self.best_traj = cand
                    self.prev_best_idx = cand.idx
                    break
            else:
                break
        if found and self.settings.refine_trajectory:
            better = self.refine(self.best_traj, obs, now)
            if better is not None:
                self.best_traj = better
        return self.best_traj


PLANNERS = {"FOP": FopOracle, "FOP+": FopPlusOracle, "FISS": FissOracle, "FISS+": FissPlusOracle}


# ----------------------------------------------------------------------------------- dense API
def dense_lattice_eval(ego6, lattice, spline: Spline2D, obs: ObstacleTable | None, *, tick, target_speed,
                       max_speed, max_accel, ego_l, ego_w, now=0, max_curvature=None):
    """Every candidate of ``lattice`` [(d, v, T)...] through the whole path; returns the arrays the
    GPU parity tests compare: cost, n, n', constraint-ok, collision, winner (last minimal survivor)."""
    ring = sat.ego_ring(ego_l, ego_w)
    trajs = []
    for seq, (d, v, T) in enumerate(lattice):
        tr = to_global(generate(Traj(), ego6, d, v, T, tick, target_speed), spline, tick)
        tr.seq = seq
        trajs.append(tr)
    cost = np.array([tr.cost_final for tr in trajs])
    n = np.array([len(tr.t) for tr in trajs])
    n_cart = np.array([len(tr.x) for tr in trajs])
    ok = np.array([passes_constraints(tr, max_speed, max_accel, max_curvature) for tr in trajs])
    coll = np.array([has_collision(tr, obs, ring, now) for tr in trajs])
    best, lo = -1, float("inf")
    for i in range(len(trajs)):
        if ok[i] and not coll[i] and lo >= cost[i]:
            lo, best = cost[i], i
    return dict(trajs=trajs, cost=cost, n=n, n_cart=n_cart, constraint_ok=ok, collision=coll, best=best)


def timed_fop_cycles(ego_batch, settings_kw, centerline, obs_arrays, *, min_t, max_t, max_target_speed,
                     ego_l, ego_w, max_speed, max_accel, now=0):
    """CPU-baseline worker: FOP ``plan()`` over a list of ego states; returns (seconds, candidates)."""
    st = Settings(**settings_kw)
    st.min_t, st.max_t = min_t, max_t
    pl = FopOracle(st, ego_l, ego_w, max_speed, max_accel)
    pl.generate_frenet_frame(centerline)
    obs = ObstacleTable(*obs_arrays) if obs_arrays is not None else None
    t0 = time.perf_counter()
    n = 0
    for ego6 in ego_batch:
        pl.plan(tuple(ego6), max_target_speed, obs, now)
        n += len(pl.last_all)
    return time.perf_counter() - t0, n


# ----------------------------------------------------------------------------------- CPU-baseline workers
_WORKER = {}


def baseline_worker_init(centerline, obs_arrays, settings_kw, min_t, max_t, max_target_speed, ego_l, ego_w,
                         max_speed, max_accel):
    """Per-process setup for the timed CPU baseline: spline fit + obstacle table, done once."""
    st = Settings(**settings_kw)
    st.min_t, st.max_t, st.highest_speed = min_t, max_t, max_target_speed
    pl = FopOracle(st, ego_l, ego_w, max_speed, max_accel)
    pl.generate_frenet_frame(centerline)
    _WORKER["pl"] = pl
    _WORKER["lattice"] = pl.lattice()
    _WORKER["obs"] = ObstacleTable(*obs_arrays) if obs_arrays is not None else None


def baseline_worker_slice(task):
    """FOP semantics (plan(), frenet_optimal_planner.py:252-259) on lattice[lo:hi] of one ego state:
    generate, convert, constraint mask, collision on the survivors.  Returns (#candidates, best cost, best id)."""
    ego6, lo, hi, now = task
    pl = _WORKER["pl"]
    st = pl.settings
    best, best_cost = -1, float("inf")
    for seq in range(lo, hi):
        d, v, T = _WORKER["lattice"][seq]
        tr = to_global(generate(Traj(), ego6, d, v, T, st.tick_t, st.highest_speed), pl.spline, st.tick_t)
        if passes_constraints(tr, pl.max_speed, pl.max_accel) and not has_collision(tr, _WORKER["obs"], pl.ego_ring, now):
            if best_cost >= tr.cost_final:
                best_cost, best = tr.cost_final, seq
    return hi - lo, best_cost, best
