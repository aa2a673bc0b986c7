"""Known-answer cases for the collision predicate (shared by the CPU test of oracle/sat_geometry.py and the GPU test of
the CUDA ``rect_sat`` behind the C-ABI).

The reference's predicate is GEOS ``Polygon.intersects`` on two rectangles (frenet_optimal_planner.py:179,189,191):
closed-set intersection, touching counts.  GEOS is absent here, so these answers come from GEOMETRY, not from any
implementation: every coordinate below is a dyadic rational (exact in binary floating point) and every placement is
axis-aligned or rotated by exactly pi/2 (shapely snaps |cos| < 2.5e-16 to 0), which makes "touching", "one ulp apart"
and "one ulp overlapping" exact statements about the inputs.  For the UNROTATED placements they survive the reference's
own arithmetic (translate only: one exact addition per coordinate).  A rotation goes through shapely's affine map
``x' = cos x - sin y + xoff`` with ``xoff = x0 - x0 cos + y0 sin`` about the bounding-box centre, which rounds at the
magnitude of the world coordinates -- a one-ulp shift of the centre does not survive it (found by these tests) -- so
the near-miss cases of rotated obstacles use a gap of 2^-40 m (~250 ulp here): tiny, but above the reference's own
rounding.  Rotations by pi/4 are not exact at all; those cases keep a 1e-9 m margin.

The ego is the rectangle EGO_L x EGO_W centred at (EGO_X, EGO_Y) with heading 0.  A case is
``(name, obstacle centre x, y, theta, length, width, expected)``.
"""
import math

import numpy as np

EGO_L, EGO_W = 4.0, 2.0
EGO_X, EGO_Y = 16.0, 0.5
HL, HW = EGO_L / 2, EGO_W / 2
TINY = 2.0 ** -40


def ulps(v, k):
    """v moved by k units in the last place (k may be negative)."""
    for _ in range(abs(k)):
        v = np.nextafter(v, math.inf if k > 0 else -math.inf)
    return float(v)


def cases():
    out = []
    # --- obstacle 2 x 1 (half extents 1, 0.5) ahead of the ego on the x axis: faces touch at x = EGO_X + 2
    cx_touch = EGO_X + HL + 1.0            # 19.0
    out += [("edge_touch_x", cx_touch, EGO_Y, 0.0, 2.0, 1.0, True),
            ("edge_gap_1ulp_x", ulps(cx_touch, +1), EGO_Y, 0.0, 2.0, 1.0, False),
            ("edge_overlap_1ulp_x", ulps(cx_touch, -1), EGO_Y, 0.0, 2.0, 1.0, True),
            ("edge_gap_4ulp_x", ulps(cx_touch, +4), EGO_Y, 0.0, 2.0, 1.0, False)]
    # --- beside the ego: faces touch at y = EGO_Y + 1
    cy_touch = EGO_Y + HW + 0.5            # 2.0
    out += [("edge_touch_y", EGO_X, cy_touch, 0.0, 2.0, 1.0, True),
            ("edge_gap_1ulp_y", EGO_X, ulps(cy_touch, +1), 0.0, 2.0, 1.0, False),
            ("edge_overlap_1ulp_y", EGO_X, ulps(cy_touch, -1), 0.0, 2.0, 1.0, True),
            ("edge_touch_y_below", EGO_X, EGO_Y - HW - 0.5, 0.0, 2.0, 1.0, True),
            ("edge_gap_1ulp_y_below", EGO_X, ulps(EGO_Y - HW - 0.5, -1), 0.0, 2.0, 1.0, False)]
    # --- corner to corner: the ego's front-left corner (18, 1.5) is the obstacle's rear-right corner
    out += [("corner_touch", cx_touch, cy_touch, 0.0, 2.0, 1.0, True),
            ("corner_gap_1ulp_x", ulps(cx_touch, +1), cy_touch, 0.0, 2.0, 1.0, False),
            ("corner_gap_1ulp_y", cx_touch, ulps(cy_touch, +1), 0.0, 2.0, 1.0, False),
            ("corner_overlap_1ulp", ulps(cx_touch, -1), ulps(cy_touch, -1), 0.0, 2.0, 1.0, True)]
    # --- partial face contact: only half of the faces overlap along y
    out += [("edge_touch_partial", cx_touch, EGO_Y + 1.0, 0.0, 2.0, 1.0, True),
            ("edge_gap_partial", ulps(cx_touch, +1), EGO_Y + 1.0, 0.0, 2.0, 1.0, False)]
    # --- rotated by exactly pi/2: a 2 x 1 rectangle becomes 1 (x) by 2 (y); faces touch at centre x = 18.5
    h = math.pi / 2
    out += [("rot90_edge_touch", EGO_X + HL + 0.5, EGO_Y, h, 2.0, 1.0, True),
            ("rot90_edge_gap_tiny", EGO_X + HL + 0.5 + TINY, EGO_Y, h, 2.0, 1.0, False),
            ("rot90_edge_overlap_tiny", EGO_X + HL + 0.5 - TINY, EGO_Y, h, 2.0, 1.0, True),
            ("rot270_edge_touch", EGO_X - HL - 0.5, EGO_Y, -h, 2.0, 1.0, True),
            ("rot90_vertex_on_edge", EGO_X + HL + 0.5, EGO_Y + HW + 1.0, h, 2.0, 1.0, True),   # corner (18, 1.5) on both
            ("rot90_vertex_gap_tiny", EGO_X + HL + 0.5, EGO_Y + HW + 1.0 + TINY, h, 2.0, 1.0, False)]
    # --- containment (no boundary crossing at all: `intersects` is still True)
    out += [("obstacle_inside_ego", EGO_X + 0.25, EGO_Y - 0.125, 0.0, 1.0, 0.5, True),
            ("ego_inside_obstacle", EGO_X - 0.5, EGO_Y + 0.25, 0.0, 16.0, 8.0, True),
            ("ego_inside_obstacle_rot90", EGO_X, EGO_Y, h, 8.0, 16.0, True),
            ("same_rectangle", EGO_X, EGO_Y, 0.0, EGO_L, EGO_W, True)]
    # --- a square of side 2 rotated by pi/4 (half diagonal sqrt 2) pointing its vertex at the ego's front face
    r2 = math.sqrt(2.0)
    q = math.pi / 4
    out += [("rot45_vertex_near_hit", EGO_X + HL + r2 - 1e-9, EGO_Y, q, 2.0, 2.0, True),
            ("rot45_vertex_near_miss", EGO_X + HL + r2 + 1e-9, EGO_Y, q, 2.0, 2.0, False),
            # ... and at the ego's corner: the diamond's edge x + y = const passes 1e-9 inside / outside (18, 1.5)
            ("rot45_edge_near_hit", EGO_X + HL + 1.0 - 1e-9, EGO_Y + HW + r2 - 1.0, q, 2.0, 2.0, True),
            ("rot45_edge_near_miss", EGO_X + HL + 1.0 + 1e-9, EGO_Y + HW + r2 - 1.0, q, 2.0, 2.0, False)]
    # --- circles overlap, rectangles do not (the circle test alone would be wrong)
    out += [("diag_far_corner", EGO_X + HL + 0.5 + 2.0 ** -10, EGO_Y + HW + 1.0 + 2.0 ** -10, 0.0, 1.0, 2.0, False),
            ("long_thin_beside", EGO_X, EGO_Y + HW + 0.25 + 2.0 ** -20, 0.0, 12.0, 0.5, False),
            ("long_thin_touch", EGO_X, EGO_Y + HW + 0.25, 0.0, 12.0, 0.5, True)]
    return out
