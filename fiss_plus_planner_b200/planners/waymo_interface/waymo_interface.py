"""Waymo ndarray wire format <-> the planners (reference: planners/waymo_interface/waymo_interface.py:24-90,146).

The reference converts every agent row of the Waymo tensors ``[N_obs, N_t, 11] = (x, y, z, l, w, h, heading, vx, vy,
valid, type)`` into CommonRoad ``DynamicObstacle`` objects (``convert_waymo_obstacle_to_cr``, :24-76), which the planner
then walks with ``state_at_time`` and turns into shapely polygons every cycle.  Here the tensors ARE the wire format of
the device obstacle table: ``convert_waymo_obstacle_to_cr`` returns a ``WaymoObstacles`` sequence
that ``plan()`` uploads with ONE call (``fiss_set_obstacles_waymo``: H2D of the tensor + a prep kernel) -- no per-agent
Python objects on the way.  The keep / cut rules are the reference's:

* agent ``i`` becomes an obstacle iff its trajectory (steps 1.. up to, excluding, the first masked step, :45-54) is
  non-empty, i.e. iff ``mask[i, 1]`` (:58); dropped agents vanish and the ones behind move up;
* a kept agent has its initial state at step 0 (never masked, :36-40) and states 1..t_end; ``state_at_time`` is ``None``
  behind that; rectangle ``(length, width) = traj[0, 3], traj[0, 4]`` (:33-34);
* ``obstacles[0].prediction.final_time_step`` (frenet_optimal_planner.py:173) is the FIRST KEPT agent's ``t_end``.

The sequence still behaves like the reference's list for other consumers (``len``, iteration, indexing give
``DynamicObstacle`` stand-ins built lazily from the same arrays).
"""
from __future__ import annotations

import collections.abc

import numpy as np

from fiss_plus_planner_b200.planners.commonroad_interface.commonroad_lite import (CustomState, DynamicObstacle, Rectangle,
                                                                                   Trajectory, TrajectoryPrediction)


def waymo_keep_rules(masks: np.ndarray):
    """(kept agent ids [M], t_end [M]) of ``convert_waymo_obstacle_to_cr`` (:45-58): kept iff ``mask[i, 1]``;
    ``t_end`` = the last step of the unbroken run of set mask entries starting at step 1."""
    masks = np.asarray(masks) != 0
    n, t = masks.shape
    if t < 2:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    keep = np.flatnonzero(masks[:, 1])
    run = masks[keep, 1:]
    # length of the leading run of True per row
    broken = ~run
    first_false = np.where(broken.any(axis=1), broken.argmax(axis=1), run.shape[1])
    return keep.astype(np.int64), first_false.astype(np.int64)   # t_end = 1 + (run length - 1) = run length


class WaymoObstacles(collections.abc.Sequence):
    """The obstacle list of one Waymo scenario, kept in its wire format.  ``plan(..., obstacles=this)`` uploads it with
    ``fiss_set_obstacles_waymo`` (float32 tensors) or, for other dtypes, as the dense table the same rules produce."""

    def __init__(self, waymo_trajs: np.ndarray, waymo_traj_masks: np.ndarray):
        trajs = np.asarray(waymo_trajs)
        assert trajs.ndim == 3 and trajs.shape[2] >= 7, "waymo_trajs must be [N_obs, N_t, >=7]"
        self.trajs = trajs
        self.masks = np.ascontiguousarray(np.asarray(waymo_traj_masks) != 0, dtype=np.uint8)
        assert self.masks.shape == trajs.shape[:2]
        self.keep, self.t_end = waymo_keep_rules(self.masks)
        self.version = 0
        self._objects = {}

    def __len__(self):
        return len(self.keep)

    @property
    def final_time_step(self) -> int:
        """``obstacles[0].prediction.final_time_step``; 0 for an empty list (never read then, :170-171)."""
        return int(self.t_end[0]) if len(self.keep) else 0

    def dense_table(self):
        """``(xyth [M, T, 3], lw [M, 2], valid [M, T], final_time_step)``: what ``marshal_obstacles`` would read off the
        reference's obstacle objects (``state_at_time`` for every step)."""
        m, t = len(self.keep), self.trajs.shape[1]
        tr = self.trajs[self.keep].astype(np.float64)
        valid = (np.arange(t)[None, :] <= self.t_end[:, None]).astype(np.uint8)
        xyth = np.where(valid[:, :, None] != 0, tr[:, :, [0, 1, 6]], 0.0)
        lw = np.ascontiguousarray(tr[:, 0, 3:5]) if m else np.zeros((0, 2))
        return xyth, lw, valid, self.final_time_step

    def upload_to(self, engine):
        if len(self.keep) == 0:
            engine.set_obstacles(None, np.zeros((0, 2)), None, 0)
        elif self.trajs.dtype == np.float32 and self.trajs.shape[2] == 11:
            kept = engine.set_obstacles_waymo(self.trajs, self.masks, -1)
            assert kept == len(self.keep)
        else:   # float64 tensors would lose bits in the float32 wire format: same rules, dense table
            engine.set_obstacles(*self.dense_table())

    def __getitem__(self, j):
        """The ``DynamicObstacle`` the reference would have built for kept agent ``j`` (:31-73), on demand."""
        if isinstance(j, slice):
            return [self[k] for k in range(*j.indices(len(self)))]
        j = range(len(self))[j]
        if j not in self._objects:
            i, tr = int(self.keep[j]), self.trajs[self.keep[j]]
            shape = Rectangle(width=tr[0, 4], length=tr[0, 3])
            init = CustomState(position=np.array([tr[0, 0], tr[0, 1]]), velocity=np.hypot(tr[0, 7], tr[0, 8]),
                               orientation=tr[0, 6], time_step=0)
            states = [CustomState(position=np.array([tr[t, 0], tr[t, 1]]), velocity=np.hypot(tr[t, 7], tr[t, 8]),
                                  orientation=tr[t, 6], time_step=t) for t in range(1, int(self.t_end[j]) + 1)]
            self._objects[j] = DynamicObstacle(i, "car", shape, init, TrajectoryPrediction(Trajectory(1, states), shape))
        return self._objects[j]


def convert_waymo_obstacle_to_cr(waymo_trajs: np.ndarray, waymo_traj_masks: np.ndarray) -> WaymoObstacles:
    """Drop-in for waymo_interface.py:24-76: same arguments; the returned sequence has the reference list's length,
    order and element attributes, and uploads to the device in one call."""
    return WaymoObstacles(waymo_trajs, waymo_traj_masks)


def convert_cr_traj_to_waymo(cr_traj, vehicle) -> np.ndarray:
    """Planned trajectory -> Waymo rows ``[n', 11]`` (waymo_interface.py:78-90)."""
    x = np.asarray(cr_traj.x)
    waymo_traj = np.vstack([x, cr_traj.y, np.zeros_like(x), np.full_like(x, vehicle.l), np.full_like(x, vehicle.w),
                            np.full_like(x, vehicle.h), cr_traj.yaw, np.zeros_like(x), np.zeros_like(x),
                            np.full_like(x, 1, dtype=int), np.full_like(x, 1, dtype=int)])
    return np.transpose(waymo_traj)
