"""The knot index of the lattice kernel's spline segment search (csrc/fiss_abi.cu::fiss_set_spline builds it,
csrc/fiss_kernels.cuh::lut_window / spline_frame2 use it), restated in Python with the same floating-point
expressions: for every abscissa in range the bracket taken from the index plus the host's iteration bound must end on
the segment `bisect_right(knots, s) - 1` that the reference's `CubicSpline1D.__search_index` returns
(cubic_spline.py:112-116).  The CUDA path itself is checked against the goldens in tests/test_gpu_parity.py."""
import bisect

import numpy as np
import pytest

CELLS = 256


def build_index(knots):
    K, s0, s1 = len(knots), knots[0], knots[-1]
    inv_h = CELLS / (s1 - s0)
    lut, i = [0] * (CELLS + 1), 0
    for c in range(CELLS + 1):
        start = s0 + c * ((s1 - s0) / CELLS)
        while i + 1 < K and knots[i + 1] <= start:
            i += 1
        lut[c] = i                                   # the largest i with knots[i] <= start of cell c
    span = 1
    for c in range(CELLS):
        lo, hi = lut[max(c - 1, 0)], min(lut[min(c + 2, CELLS)] + 1, K - 1)
        span = max(span, hi - lo)
    iters = 0
    while (1 << iters) < span:
        iters += 1
    return lut, inv_h, iters


def search(knots, lut, inv_h, iters, s):
    K = len(knots)
    c = int((s - knots[0]) * inv_h)
    c = max(0, min(c, CELLS - 1))
    lo, hi = lut[max(c - 1, 0)], min(lut[min(c + 2, CELLS)] + 1, K - 1)
    for _ in range(iters):                           # the kernel's bisection step, unchanged
        mid = (lo + hi) >> 1
        le, is_open = knots[mid] <= s, hi - lo > 1
        if le and is_open:
            lo = mid
        if not le and is_open:
            hi = mid
    return lo


def _cases():
    rng = np.random.default_rng(20231001)
    x, y = 5.0 * np.arange(81) + 465.7, 3.0 * np.sin(5.0 * np.arange(81) / 30.0) - 304.8
    yield "bench line", np.concatenate([[0.0], np.cumsum(np.hypot(np.diff(x), np.diff(y)))]), 1
    yield "irregular", np.concatenate([[0.0], np.cumsum(rng.uniform(0.05, 9.0, 300))]), None
    yield "clustered", np.concatenate([[0.0], np.cumsum(rng.choice([0.01, 0.01, 0.01, 25.0], 500))]), None
    yield "two knots", np.array([0.0, 1.0]), 0
    yield "duplicates", np.concatenate([[3.0], 3.0 + np.cumsum(np.r_[np.zeros(3), rng.uniform(0.5, 1.0, 40)])]), None


@pytest.mark.parametrize("name,knots,want_iters", list(_cases()), ids=[c[0] for c in _cases()])
def test_index_bracket_ends_on_the_bisect_segment(name, knots, want_iters):
    knots = [float(k) for k in knots]
    lut, inv_h, iters = build_index(knots)
    if want_iters is not None:
        assert iters == want_iters                  # evenly spaced knots: one step instead of log2 K
    rng = np.random.default_rng(7)
    probes = list(rng.uniform(knots[0], knots[-1], 20000)) + knots[:-1]
    probes += [float(np.nextafter(k, -np.inf)) for k in knots[1:]] + [float(np.nextafter(k, np.inf)) for k in knots[:-1]]
    for s in probes:
        if knots[0] <= s < knots[-1]:
            assert search(knots, lut, inv_h, iters, s) == bisect.bisect_right(knots, s) - 1, (name, s)
