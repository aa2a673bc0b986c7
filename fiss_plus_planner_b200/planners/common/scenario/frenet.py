"""Frenet / Cartesian data carriers of the drop-in API.

Same class names, fields and method signatures as the reference's
planners/common/scenario/frenet.py (``State`` :6-13, ``FrenetState`` :15-99,
``FrenetTrajectory`` :101-219), because ``planning.py:135-138`` and user code read them
(``best.x``, ``best.state_at_time_step(1)`` ...).  The difference is where the numbers come
from: a ``FrenetTrajectory`` here is filled from the device's winner / candidate records
(``FrenetTrajectory.from_device_record``) as float64 NumPy arrays, never computed on the host.
"""
from __future__ import annotations

import copy

import numpy as np

from fiss_plus_planner_b200.planners.common.geometry.math_utils import unifyAngleRange
from fiss_plus_planner_b200.planners.common.scenario.lane import LaneType


class State(object):
    def __init__(self, t: float = 0.0, x: float = 0.0, y: float = 0.0, yaw: float = 0.0, v: float = 0.0, a: float = 0.0):
        self.t = t
        self.x = x
        self.y = y
        self.yaw = yaw
        self.v = v
        self.a = a


class FrenetState(object):
    def __init__(self, t: float = 0.0,
                 s: float = 0.0, s_d: float = 0.0, s_dd: float = 0.0, s_ddd: float = 0.0,
                 d: float = 0.0, d_d: float = 0.0, d_dd: float = 0.0, d_ddd: float = 0.0):
        self.t = t
        self.s = s
        self.s_d = s_d
        self.s_dd = s_dd
        self.s_ddd = s_ddd
        self.d = d
        self.d_d = d_d
        self.d_dd = d_dd
        self.d_ddd = d_ddd

    def __str__(self):
        return f'FrenetState with d={self.d:.2f}, s_d={self.s_d:.2f}, t={self.t:.2f}'

    def as_ego6(self) -> np.ndarray:
        """(s, s_d, s_dd, d, d_d, d_dd): the six numbers the device path reads."""
        return np.array([self.s, self.s_d, self.s_dd, self.d, self.d_d, self.d_dd], dtype=np.float64)

    def from_state(self, state: State, polyline: np.ndarray):
        """Cartesian -> Frenet against the 0.1 m polyline ``[m, >=3] = (x, y, yaw, ...)``.

        Follows frenet.py:32-99 step by step: nearest point, "next" waypoint by heading test,
        projection on the prev->next chord, sign convention ``wp_yaw <= x_yaw -> d < 0``
        (CommonRoad), arc length as the sum of chord lengths up to the previous waypoint,
        ``s_dd = d_dd = 0``.  Returns ``state`` like the reference does (:99).
        """
        px = polyline[:, 0]
        py = polyline[:, 1]
        nearest = int(np.argmin(np.hypot(px - state.x, py - state.y)))
        heading = np.arctan2(py[nearest] - state.y, px[nearest] - state.x)
        angle = abs(state.yaw - heading)
        angle = min(2 * np.pi - angle, angle)
        nxt = nearest + 1 if angle > np.pi / 2 else nearest
        if nxt < 1:
            nxt = 1
        elif nxt >= polyline.shape[0]:
            nxt = polyline.shape[0] - 1
        prv = max(nxt - 1, 0)

        n_x = px[nxt] - px[prv]
        n_y = py[nxt] - py[prv]
        x_x = state.x - px[prv]
        x_y = state.y - py[prv]
        x_yaw = np.arctan2(x_y, x_x)
        proj_norm = (x_x * n_x + x_y * n_y) / (n_x * n_x + n_y * n_y)
        proj_x = proj_norm * n_x
        proj_y = proj_norm * n_y

        self.d = np.hypot(x_x - proj_x, x_y - proj_y)
        wp_yaw = polyline[prv, 2]
        delta_yaw = unifyAngleRange(state.yaw - wp_yaw)
        if wp_yaw <= x_yaw:
            self.d *= -1

        self.s = 0
        for i in range(prv):
            self.s += np.hypot(px[i + 1] - px[i], py[i + 1] - py[i])

        self.t = state.t
        self.s_d = state.v * np.cos(delta_yaw)
        self.s_dd = 0.0
        self.s_ddd = 0.0
        self.d_d = state.v * np.sin(delta_yaw)
        self.d_dd = 0.0
        self.d_ddd = 0.0
        return state


# order of the rows in a device "full record" (fiss_abi.h: FISS_REC_*)
RECORD_FIELDS = ("t", "s", "s_d", "s_dd", "s_ddd", "d", "d_d", "d_dd", "d_ddd",
                 "x", "y", "yaw", "ds", "c", "c_d", "c_dd")


class FrenetTrajectory(object):
    """One candidate: lattice index, flags, costs, Frenet arrays (length n), Cartesian arrays
    (``x, y, yaw`` length n', ``ds, c`` n'-1, ``c_d`` n'-2, ``c_dd`` n'-3) -- frenet.py:113-148."""

    def __init__(self):
        self.idx = np.array([-1, -1, -1])
        self.lane_id = -1
        self.lane_type = LaneType.UNDEFINED

        self.is_generated = False
        self.is_searched = False
        self.constraint_passed = False
        self.collision_passed = False
        self.end_state = None

        self.cost_fix = 0.0
        self.cost_dyn = 0.0
        self.cost_heu = 0.0
        self.cost_est = 0.0
        self.cost_final = 0.0

        self.t = []
        self.s = []
        self.s_d = []
        self.s_dd = []
        self.s_ddd = []
        self.d = []
        self.d_d = []
        self.d_dd = []
        self.d_ddd = []
        self.x = []
        self.y = []
        self.yaw = []
        self.ds = []
        self.c = []
        self.c_d = []
        self.c_dd = []

    # cost-ordered comparisons (frenet.py:150-166)
    def __eq__(self, other):
        return self.cost_final == other.cost_final

    def __ne__(self, other):
        return self.cost_final != other.cost_final

    def __lt__(self, other):
        return self.cost_final < other.cost_final

    def __le__(self, other):
        return self.cost_final <= other.cost_final

    def __gt__(self, other):
        return self.cost_final > other.cost_final

    def __ge__(self, other):
        return self.cost_final >= other.cost_final

    __hash__ = object.__hash__

    def __repr__(self):
        return "%f" % (self.cost_final)

    def __str__(self):
        return (f'FrenetTrajectory with cost_final={self.cost_final:.2f},  d={self.end_state.d:.2f}, '
                f's_d={self.end_state.s_d:.2f}, t={self.end_state.t:.2f}')

    def fill_from_device_record(self, rec: np.ndarray, n: int, n_cart: int, cost: float) -> "FrenetTrajectory":
        """``rec`` is one ``[16, n_stride]`` float64 full record written by the GPU.

        Frenet rows keep length ``n``; Cartesian rows are cut to the truncation length ``n_cart``
        (= n', the first step whose arc length leaves the reference line) with the reference's
        ragged lengths: ``ds, c`` n'-1, ``c_d`` n'-2, ``c_dd`` n'-3; for n' < 2 they stay empty
        (frenet_optimal_planner.py:121-134).
        """
        rec = rec.copy()            # ONE private copy; the fields are views into it (the caller's buffer is reused)
        (self.t, self.s, self.s_d, self.s_dd, self.s_ddd, self.d, self.d_d, self.d_dd, self.d_ddd) = rec[:9, :n]
        self.x = rec[9, :n_cart]
        self.y = rec[10, :n_cart]
        if n_cart >= 2:
            self.yaw = rec[11, :n_cart]
            self.ds = rec[12, :n_cart - 1]
            self.c = rec[13, :n_cart - 1]
            self.c_d = rec[14, :max(n_cart - 2, 0)]
            self.c_dd = rec[15, :max(n_cart - 3, 0)]
        else:
            self.yaw, self.ds, self.c, self.c_d, self.c_dd = [], [], [], [], []
        self.cost_final = cost
        self.is_generated = True
        return self

    def state_at_time_step(self, t: int) -> State:
        assert t < len(self.s) and t >= 0
        return State(self.t[t], self.x[t], self.y[t], self.yaw[t], self.s_d[t], self.s_dd[t])

    def frenet_state_at_time_step(self, t: int) -> FrenetState:
        assert t < len(self.s) and t >= 0
        return FrenetState(self.t[t],
                           self.s[t], self.s_d[t], self.s_dd[t], self.s_ddd[t],
                           self.d[t], self.d_d[t], self.d_dd[t], self.d_ddd[t])

    def forward_t_steps(self, steps: int):
        """Copy with the first ``steps`` samples dropped (frenet.py:198-219; c_d/c_dd untouched)."""
        if steps < 0 or steps >= len(self.t):
            return None
        out = copy.deepcopy(self)
        for name in ("t", "s", "s_d", "s_dd", "s_ddd", "d", "d_d", "d_dd", "d_ddd", "x", "y", "yaw", "ds", "c"):
            setattr(out, name, getattr(out, name)[steps:])
        return out
