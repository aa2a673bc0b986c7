"""Seeded synthetic scenes for the BASELINE.json configs.

Everything is host NumPy float64 and deterministic in ``seed``; the same arrays feed the CUDA
path, the oracle and the golden-vector generator, so all three see bit-identical inputs.
There is no dataset behind the benchmark (no network): ``"data": "synthetic"``.
"""
from __future__ import annotations

import dataclasses
import types

import numpy as np

from fiss_plus_planner_b200.planners.common.geometry.cubic_spline import CubicSpline2D

SEED_BASE = 20231001

# VW Vanagon footprint (SMP/maneuver_automaton/maneuver_automaton.py:47-48) and the
# commonroad-vehicle-models limits quoted in SURVEY 8(c)(ii)
EGO_L, EGO_W = 4.569, 1.844
EGO_V_MAX, EGO_A_MAX = 41.7, 11.5


def vehicle_params(l=EGO_L, w=EGO_W, v_max=EGO_V_MAX, a_max=EGO_A_MAX):
    """Duck-typed stand-in for ``VehicleParameterMapping['VW_VANAGON'].value`` (planning.py:297-298)."""
    return types.SimpleNamespace(
        l=l, w=w, a=1.1508, b=1.3211, T_f=1.5, T_r=1.5,
        longitudinal=types.SimpleNamespace(v_max=v_max, a_max=a_max),
        steering=types.SimpleNamespace(max=1.023, v_max=0.4, kappa_dot_max=0.4, kappa_dot_dot_max=20.0))


def reference_line(num_knots: int = 81, spacing: float = 5.0) -> np.ndarray:
    """Centerline ``[K, 2]`` with a Flensburg-like world offset (stresses FP precision)."""
    k = np.arange(num_knots, dtype=np.float64)
    return np.column_stack((spacing * k + 465.7, 3.0 * np.sin(spacing * k / 30.0) - 304.8))


def ego_states(rng: np.random.Generator, batch: int, symmetric_first: bool = False) -> np.ndarray:
    """``[B, 6] = (s0, s_d0, s_dd0, d0, d_d0, d_dd0)``."""
    ego = np.column_stack((rng.uniform(5, 50, batch), rng.uniform(2, 13, batch), rng.uniform(-1, 1, batch),
                           rng.uniform(-0.8, 0.8, batch), rng.uniform(-0.3, 0.3, batch),
                           rng.uniform(-0.2, 0.2, batch)))
    if symmetric_first:
        ego[0, 3:] = 0.0  # d0 = d_d0 = d_dd0 = 0: +-d lattice points tie exactly
    return np.ascontiguousarray(ego)


@dataclasses.dataclass
class ObstacleSet:
    xyth: np.ndarray          # [M, T_obs, 3] x, y, orientation
    lw: np.ndarray            # [M, 2] length, width
    valid: np.ndarray         # [M, T_obs] bool: state_at_time(t) is not None
    final_time_step: int      # obstacles[0].prediction.final_time_step


def obstacles(rng: np.random.Generator, spline: CubicSpline2D, num: int, t_obs: int = 100,
              dt: float = 0.1, early_end_frac: float = 0.2) -> ObstacleSet:
    """Rectangles in constant Frenet motion, mapped through ``spline`` to world poses."""
    lw = np.column_stack((rng.uniform(3.5, 7.5, num), rng.uniform(1.6, 2.2, num)))
    s0 = rng.uniform(0, 150, num)
    d0 = rng.uniform(-4, 4, num)
    v0 = rng.uniform(0, 12, num)
    # Deviation from the letter of SURVEY 8(d): an obstacle that shares the ego corridor
    # (|d| < 3 m) starts ahead of every ego start (s0 <= 50 m) instead of anywhere in [0, 150].
    # As written, ~70 % of problems with M=32 would collide at step 0 and the early exit would
    # flatter the collision stage; this way collisions happen when fast candidates catch up.
    in_corridor = np.abs(d0) < 3.0
    s0 = np.where(in_corridor, 57.0 + (150.0 - 57.0) * (s0 / 150.0), s0)
    xyth = np.zeros((num, t_obs, 3))
    for j in range(num):
        for t in range(t_obs):
            s = s0[j] + v0[j] * (t * dt)
            px, py = spline.calc_position(s)
            yaw = spline.calc_yaw(s)
            xyth[j, t] = (px - d0[j] * np.sin(yaw), py + d0[j] * np.cos(yaw), yaw)
    valid = np.ones((num, t_obs), dtype=bool)
    ends_early = rng.uniform(0, 1, num) < early_end_frac
    end_step = rng.integers(10, t_obs, num)
    for j in range(num):
        if ends_early[j]:
            valid[j, end_step[j]:] = False
    return ObstacleSet(np.ascontiguousarray(xyth), np.ascontiguousarray(lw), valid, t_obs)


@dataclasses.dataclass
class Scene:
    centerline: np.ndarray
    spline: CubicSpline2D
    ego: np.ndarray
    obs: ObstacleSet
    num_samples: tuple        # (num_width, num_speed, num_t)
    min_t: float
    max_t: float
    max_target_speed: float = 13.4112
    time_step_now: int = 0


# name -> (config #, lattice, (min_t, max_t), obstacles, default batch)
CONFIGS = {
    "cfg1_demo_substitute": (1, (5, 5, 5), (8.0, 10.0), 27, 1),
    "cfg2_single_ego_8obs": (2, (9, 6, 5), (4.0, 5.0), 8, 1),
    "cfg3_64obs": (3, (9, 6, 5), (4.0, 5.0), 64, 1),
    "cfg4_batch4096_32obs": (4, (9, 6, 5), (4.0, 5.0), 32, 4096),
    "cfg5_fine_lattice": (5, (33, 17, 9), (8.0, 10.0), 32, 1),
}


def make_scene(name: str, batch: int | None = None, num_obstacles: int | None = None,
               seed: int | None = None, symmetric_first: bool = False) -> Scene:
    cfg, lattice, (min_t, max_t), m, b = CONFIGS[name]
    rng = np.random.default_rng(SEED_BASE + cfg if seed is None else seed)
    line = reference_line()
    spline = CubicSpline2D(line[:, 0], line[:, 1])
    ego = ego_states(rng, b if batch is None else batch, symmetric_first)
    obs = obstacles(rng, spline, m if num_obstacles is None else num_obstacles)
    return Scene(line, spline, ego, obs, lattice, min_t, max_t)
