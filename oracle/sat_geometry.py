"""ORACLE (test infrastructure, not product code): the collision predicate in FP64 on the CPU.

Restates what ``FrenetOptimalPlanner.construct_polygon`` / ``has_collision`` ask of shapely 2.0.0
(GEOS) at /root/reference/planners/frenet_optimal_planner.py:162-166,179,189,191:

* ``affinity.translate(polygon, xoff=x, yoff=y)``                     -> ``translate``
* ``affinity.rotate(polygon, yaw, use_radians=True)`` (origin='center' = centre of the
  bounding box of the translated polygon; |cos|,|sin| < 2.5e-16 snapped to 0; affine map
  ``x' = a x + b y + xoff``)                                           -> ``rotate_about_bbox_center``
* ``Polygon.intersects`` on two convex rings: closed-set intersection (touching counts)
                                                                       -> ``sat_closed`` / ``sat_closed_many``

shapely / GEOS are third-party dependencies pinned in environment.yml:184 and are NOT under
/root/reference and NOT installed here, so this predicate cannot be checked against GEOS:
COLLISION PARITY IS UNPINNED at that boundary.  The same functions back the
shapely stub used when the real reference is executed to produce tests/golden/*.npz, so
"reference" collision masks in the goldens are reference control flow + this predicate.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.
"""
from __future__ import annotations

import math

import numpy as np


def ego_ring(length: float, width: float) -> np.ndarray:
    """Ego footprint, clockwise from front-left (planners/common/vehicle/vehicle.py:24-30)."""
    hl, hw = length / 2, width / 2
    return np.array([(hl, hw), (hl, -hw), (-hl, -hw), (-hl, hw)], dtype=np.float64)


def obstacle_ring(length: float, width: float) -> np.ndarray:
    """CommonRoad ``Rectangle(length, width).shapely_object`` vertex order, centred on the origin."""
    hl, hw = 0.5 * length, 0.5 * width
    return np.array([(-hl, -hw), (-hl, hw), (hl, hw), (hl, -hw)], dtype=np.float64)


def translate(ring: np.ndarray, xoff: float, yoff: float) -> np.ndarray:
    return np.column_stack((ring[:, 0] + xoff, ring[:, 1] + yoff))


def rotate_about_bbox_center(ring: np.ndarray, angle: float) -> np.ndarray:
    cosp = math.cos(angle)
    sinp = math.sin(angle)
    if abs(cosp) < 2.5e-16:
        cosp = 0.0
    if abs(sinp) < 2.5e-16:
        sinp = 0.0
    x0 = (ring[:, 0].min() + ring[:, 0].max()) / 2.0
    y0 = (ring[:, 1].min() + ring[:, 1].max()) / 2.0
    xoff = x0 - x0 * cosp + y0 * sinp
    yoff = y0 - x0 * sinp - y0 * cosp
    x = ring[:, 0]
    y = ring[:, 1]
    return np.column_stack((cosp * x + (-sinp) * y + xoff, sinp * x + cosp * y + yoff))


def place(ring: np.ndarray, x: float, y: float, yaw: float) -> np.ndarray:
    """construct_polygon (frenet_optimal_planner.py:162-166): translate, then rotate."""
    return rotate_about_bbox_center(translate(ring, x, y), yaw)


def sat_closed(p: np.ndarray, q: np.ndarray) -> bool:
    """Closed-set separating-axis test on two convex rings ``[n, 2]``; True = they intersect."""
    for poly in (p, q):
        n = len(poly)
        for i in range(n):
            ex = poly[(i + 1) % n, 0] - poly[i, 0]
            ey = poly[(i + 1) % n, 1] - poly[i, 1]
            ax, ay = -ey, ex
            pp = p[:, 0] * ax + p[:, 1] * ay
            qq = q[:, 0] * ax + q[:, 1] * ay
            if pp.max() < qq.min() or qq.max() < pp.min():
                return False
    return True


def sat_closed_many(p: np.ndarray, qs: np.ndarray) -> np.ndarray:
    """``sat_closed(p, qs[j])`` for every j, same arithmetic, vectorised over ``qs [M, 4, 2]``."""
    m = len(qs)
    hit = np.ones(m, dtype=bool)
    # axes from the edges of p (shared by all pairs)
    n = len(p)
    for i in range(n):
        ex = p[(i + 1) % n, 0] - p[i, 0]
        ey = p[(i + 1) % n, 1] - p[i, 1]
        ax, ay = -ey, ex
        pp = p[:, 0] * ax + p[:, 1] * ay
        qq = qs[:, :, 0] * ax + qs[:, :, 1] * ay
        hit &= ~((pp.max() < qq.min(axis=1)) | (qq.max(axis=1) < pp.min()))
    # axes from the edges of each q
    nq = qs.shape[1]
    for i in range(nq):
        ex = qs[:, (i + 1) % nq, 0] - qs[:, i, 0]
        ey = qs[:, (i + 1) % nq, 1] - qs[:, i, 1]
        ax, ay = -ey, ex
        pp = p[None, :, 0] * ax[:, None] + p[None, :, 1] * ay[:, None]
        qq = qs[:, :, 0] * ax[:, None] + qs[:, :, 1] * ay[:, None]
        hit &= ~((pp.max(axis=1) < qq.min(axis=1)) | (qq.max(axis=1) < pp.min(axis=1)))
    return hit
