#!/usr/bin/env python
"""Fit the polynomial used by fiss::fast_atan2 (csrc/fiss_math.cuh): atan(q) = q * P(q^2) on q in [0, 1].

Chebyshev interpolation of g(u) = atan(sqrt(u))/sqrt(u) on [0, 1] in 80-digit arithmetic (mpmath), converted
to the monomial basis, rounded to double; then the error of the ROUNDED polynomial evaluated in float64
Horner form is measured against mpmath.  Prints the coefficient table for the header.
"""
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 80


def g(u):
    if u == 0:
        return mp.mpf(1)
    r = mp.sqrt(u)
    return mp.atan(r) / r


def cheb_fit(deg):
    n = deg + 1
    nodes = [(mp.cos(mp.pi * (2 * k + 1) / (2 * n)) + 1) / 2 for k in range(n)]      # Chebyshev nodes on [0, 1]
    vals = [g(x) for x in nodes]
    # interpolating polynomial in the monomial basis by solving the Vandermonde system exactly enough
    A = mp.matrix(n, n)
    for i, x in enumerate(nodes):
        for j in range(n):
            A[i, j] = x ** j
    c = mp.lu_solve(A, mp.matrix(vals))
    return [c[j] for j in range(n)]


def max_err(coef, samples=20001):
    c = np.array([float(v) for v in coef])
    q = np.linspace(0.0, 1.0, samples)
    u = q * q
    p = np.full_like(u, c[-1])
    for k in range(len(c) - 2, -1, -1):
        p = p * u + c[k]                     # (FMA on the device: one rounding less)
    got = q * p
    worst = 0.0
    for qi, gi in zip(q[:: max(1, samples // 4001)], got[:: max(1, samples // 4001)]):
        ref = mp.atan(mp.mpf(float(qi)))
        err = abs(mp.mpf(float(gi)) - ref)
        rel = err / ref if ref != 0 else err
        worst = max(worst, float(rel))
    return worst


if __name__ == "__main__":
    for deg in ([int(a) for a in sys.argv[1:]] or [17, 18, 19, 20, 21, 22]):
        coef = cheb_fit(deg)
        print(f"// degree {deg} in q^2: max relative error of the float64 Horner evaluation {max_err(coef):.3e}")
        if len(sys.argv) > 1:
            for k, v in enumerate(coef):
                print(f"    {float(v)!r},  // q^{2 * k + 1}")
