"""Host-side natural cubic spline for the reference line (one-time setup, float64).

Mirrors the public surface of the reference's ``CubicSpline1D`` / ``CubicSpline2D``
(planners/common/geometry/cubic_spline.py:5-233): same constructor arguments, same attribute
names (``x, y, nx, a, b, c, d`` / ``s, sx, sy, ds``), same ``calc_*`` methods with the same
out-of-range -> ``None`` behaviour.  The fit stays on the host: it
runs once per scenario and its coefficients must be the ones the reference computes, so the same
dense ``nx x nx`` system is assembled (cubic_spline.py:118-142) and handed to the same LAPACK
``gesv`` through ``np.linalg.solve`` (cubic_spline.py:36).  What is new here is
``CubicSpline2D.device_table()``: the ``[9, K]`` float64 table (knots + 4 x-coefficients + 4
y-coefficients, one row each, segment-major) that ``fiss_set_spline`` uploads and the CUDA kernels
stage into shared memory.  Per-step evaluation for the planners happens on the GPU, not here.
"""
from __future__ import annotations

import bisect
import math

import numpy as np


class CubicSpline1D:
    def __init__(self, x, y):
        xs = np.asarray(x, dtype=np.float64)
        ys = np.asarray(y, dtype=np.float64)
        h = np.diff(xs)
        if np.any(h < 0):
            raise ValueError("x coordinates must be sorted in ascending order")
        self.x = x
        self.y = y
        self.nx = n = len(xs)
        self._knots = xs

        # Natural spline: second-derivative coefficients c from the dense system A c = B.
        A = np.zeros((n, n))
        B = np.zeros(n)
        A[0, 0] = 1.0
        A[n - 1, n - 1] = 1.0
        inner = np.arange(1, n - 1)
        A[inner, inner] = 2.0 * (h[:-1] + h[1:])
        A[inner, inner - 1] = h[:-1]
        A[inner, inner + 1] = h[1:]
        A[0, 1] = 0.0
        A[n - 1, n - 2] = 0.0
        B[1:n - 1] = 3.0 * (ys[2:] - ys[1:-1]) / h[1:] - 3.0 * (ys[1:-1] - ys[:-2]) / h[:-1]
        c = np.linalg.solve(A, B)

        self.a = [v for v in ys]
        self.c = c
        self.d = list((c[1:] - c[:-1]) / (3.0 * h))
        self.b = list(1.0 / h * (ys[1:] - ys[:-1]) - h / 3.0 * (2.0 * c[:-1] + c[1:]))

    def _segment(self, x):
        return bisect.bisect(self.x, x) - 1

    def _outside(self, x) -> bool:
        return x < self.x[0] or x > self.x[-1]

    def calc_position(self, x):
        if self._outside(x):
            return None
        i = self._segment(x)
        dx = x - self.x[i]
        return self.a[i] + self.b[i] * dx + self.c[i] * dx ** 2.0 + self.d[i] * dx ** 3.0

    def calc_first_derivative(self, x):
        if self._outside(x):
            return None
        i = self._segment(x)
        dx = x - self.x[i]
        return self.b[i] + 2.0 * self.c[i] * dx + 3.0 * self.d[i] * dx ** 2.0

    def calc_second_derivative(self, x):
        if self._outside(x):
            return None
        i = self._segment(x)
        dx = x - self.x[i]
        return 2.0 * self.c[i] + 6.0 * self.d[i] * dx

    def coefficient_rows(self) -> np.ndarray:
        """[4, K] rows a, b, c, d; b and d are padded with 0 in the (never evaluated) last slot."""
        k = self.nx
        out = np.zeros((4, k))
        out[0] = self.a
        out[1, :k - 1] = self.b
        out[2] = self.c
        out[3, :k - 1] = self.d
        return out


class _TableSpline1D(CubicSpline1D):
    """``CubicSpline1D`` interface over coefficient rows that were fitted elsewhere (the device)."""

    def __init__(self, knots, rows):
        k = len(knots)
        self.x = [float(v) for v in knots]
        self.y = [float(v) for v in rows[0]]
        self.nx = k
        self._knots = np.asarray(knots, dtype=np.float64)
        self.a = [float(v) for v in rows[0]]
        self.b = [float(v) for v in rows[1][:k - 1]]
        self.c = np.array(rows[2], dtype=np.float64)
        self.d = [float(v) for v in rows[3][:k - 1]]


class CubicSpline2D:
    @classmethod
    def from_device_table(cls, table: np.ndarray) -> "CubicSpline2D":
        """Host view of a ``[9, K]`` table produced by ``FissEngine.fit_splines`` (same ``calc_*`` interface)."""
        table = np.asarray(table, dtype=np.float64)
        self = cls.__new__(cls)
        self.s = [float(v) for v in table[0]]
        self.ds = np.diff(table[0])
        self.sx = _TableSpline1D(table[0], table[1:5])
        self.sy = _TableSpline1D(table[0], table[5:9])
        return self

    def __init__(self, x, y):
        dx = np.diff(x)
        dy = np.diff(y)
        self.ds = np.hypot(dx, dy)
        s = [0]
        s.extend(np.cumsum(self.ds))
        self.s = s
        self.sx = CubicSpline1D(self.s, x)
        self.sy = CubicSpline1D(self.s, y)

    def calc_position(self, s):
        return self.sx.calc_position(s), self.sy.calc_position(s)

    def calc_curvature(self, s):
        dx = self.sx.calc_first_derivative(s)
        ddx = self.sx.calc_second_derivative(s)
        dy = self.sy.calc_first_derivative(s)
        ddy = self.sy.calc_second_derivative(s)
        return (ddy * dx - ddx * dy) / ((dx ** 2 + dy ** 2) ** (3 / 2))

    def calc_yaw(self, s):
        dx = self.sx.calc_first_derivative(s)
        dy = self.sy.calc_first_derivative(s)
        return math.atan2(dy, dx)

    def device_table(self) -> np.ndarray:
        """[9, K] float64, C-contiguous: knots, ax, bx, cx, dx, ay, by, cy, dy."""
        k = len(self.s)
        tab = np.empty((9, k), dtype=np.float64)
        tab[0] = np.asarray(self.s, dtype=np.float64)
        tab[1:5] = self.sx.coefficient_rows()
        tab[5:9] = self.sy.coefficient_rows()
        return tab
