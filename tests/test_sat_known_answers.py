"""Known-answer tests of the collision predicate itself (CPU): oracle/sat_geometry.py -- the stand-in for
shapely/GEOS behind the goldens AND the oracle -- against answers that follow from exact geometry (tests/sat_cases.py):
touching edges and vertices, containment, one-ulp near misses.  The CUDA ``rect_sat`` gets the same cases through the
C-ABI in tests/test_gpu_sat_known_answers.py."""
import math

import numpy as np
import pytest

import sat_cases as sc
from oracle import sat_geometry as sat

CASES = sc.cases()


def _ego():
    return sat.place(sat.ego_ring(sc.EGO_L, sc.EGO_W), sc.EGO_X, sc.EGO_Y, 0.0)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_sat_closed_known_answer(case):
    name, cx, cy, th, length, width, expected = case
    obstacle = sat.place(sat.obstacle_ring(length, width), cx, cy, th)
    ego = _ego()
    assert sat.sat_closed(ego, obstacle) is expected
    assert sat.sat_closed(obstacle, ego) is expected                      # symmetric
    assert bool(sat.sat_closed_many(ego, obstacle[None])[0]) is expected  # the vectorised twin the oracle calls


def test_cases_cover_both_answers_and_are_exact():
    assert sum(c[-1] for c in CASES) >= 12 and sum(not c[-1] for c in CASES) >= 10
    # the axis-aligned placements are exact: the placed rings have the corner coordinates one computes by hand
    ego = _ego()
    np.testing.assert_array_equal(np.sort(np.unique(ego[:, 0])), [14.0, 18.0])
    np.testing.assert_array_equal(np.sort(np.unique(ego[:, 1])), [-0.5, 1.5])
    ring = sat.place(sat.obstacle_ring(2.0, 1.0), 18.5, 2.5, math.pi / 2)   # cos(pi/2) = 6e-17 snaps to 0
    np.testing.assert_array_equal(np.sort(np.unique(ring[:, 0])), [18.0, 19.0])
    np.testing.assert_array_equal(np.sort(np.unique(ring[:, 1])), [1.5, 3.5])


def test_rotation_about_bbox_centre_equals_rotation_about_own_centre_for_centred_rectangles():
    """construct_polygon translates, then rotates about the BOUNDING-BOX centre (shapely default origin='center',
    frenet_optimal_planner.py:163-164).  For a rectangle centred on the origin that is the rectangle's own centre, which
    is what the CUDA predicate assumes (centre / axis form)."""
    rng = np.random.default_rng(5)
    for _ in range(50):
        l, w = rng.uniform(1, 8, 2)
        x, y = rng.uniform(-500, 500, 2)
        th = rng.uniform(-math.pi, math.pi)
        ring = sat.place(sat.obstacle_ring(l, w), x, y, th)
        np.testing.assert_allclose(ring.mean(axis=0), (x, y), atol=1e-12)
        c, s = math.cos(th), math.sin(th)
        want = np.array([(x + c * px - s * py, y + s * px + c * py) for px, py in sat.obstacle_ring(l, w)])
        np.testing.assert_allclose(ring, want, atol=1e-12)
