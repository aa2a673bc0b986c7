"""Harness that imports the UNMODIFIED reference (SS47816/fiss_plus_planner) from /root/reference.

Test infrastructure only -- used by ``make_golden.py`` (here, in the build container) to run the
reference's own Python modules on seeded inputs and freeze their outputs under ``tests/golden/``.
Nothing in the product, in ``-m gpu`` tests, in ``smoke()`` or in ``bench.py`` imports this file:
/root/reference does not exist on the GPU box.

What is stubbed, and why:

* ``commonroad.scenario.scenario.Scenario`` / ``commonroad.scenario.obstacle.Obstacle`` --
  annotation-only uses (frenet_optimal_planner.py:5,59; fop_plus_planner.py:3-4,16). Empty classes.
* ``shapely`` (GEOS) -- not installed and no network.  The three calls on the hot path
  (``Polygon(coords)``, ``affinity.translate``, ``affinity.rotate(origin='center',
  use_radians=True)``, ``Polygon.intersects``; frenet_optimal_planner.py:163-164,191 and
  vehicle.py:31) are replaced by a NumPy stand-in: the affine maps follow shapely 2.0's published
  formulas and ``intersects`` is a closed-set separating-axis test on convex polygons in FP64.
  COLLISION PARITY IS THEREFORE UNPINNED AT THE GEOS BOUNDARY; every other stage is pinned by
  executing reference code.
* obstacles are duck-typed: ``.prediction.final_time_step``, ``.state_at_time(t)`` returning
  ``None`` or an object with ``.position`` / ``.orientation``, ``.obstacle_shape.shapely_object``.
"""
from __future__ import annotations

import math
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


# --------------------------------------------------------------------------- shapely stand-in
# The arithmetic lives in oracle/sat_geometry.py so that the goldens and the oracle share ONE
# definition of the (unpinned) collision predicate.
import os as _os
sys.path.insert(0, _os.path.abspath(_os.path.join(_os.path.dirname(__file__), "..", "..")))
from oracle import sat_geometry as _sat  # noqa: E402


class Polygon:
    """Convex polygon given by its exterior ring (closing vertex optional)."""

    def __init__(self, coords):
        pts = np.asarray(coords, dtype=np.float64)
        if len(pts) >= 2 and np.array_equal(pts[0], pts[-1]):
            pts = pts[:-1]
        self.pts = pts

    def intersects(self, other: "Polygon") -> bool:
        return _sat.sat_closed(self.pts, other.pts)


def _wrap(pts) -> Polygon:
    out = Polygon.__new__(Polygon)
    out.pts = pts
    return out


def _translate(geom: Polygon, xoff=0.0, yoff=0.0, zoff=0.0) -> Polygon:
    return _wrap(_sat.translate(geom.pts, xoff, yoff))


def _rotate(geom: Polygon, angle, origin="center", use_radians=False) -> Polygon:
    assert origin == "center" and use_radians
    return _wrap(_sat.rotate_about_bbox_center(geom.pts, angle))


def install_stubs():
    """Put the commonroad / shapely stand-ins in sys.modules and the reference on sys.path."""
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "shapely" in sys.modules and getattr(sys.modules["shapely"], "_fiss_stub", False):
        return
    shapely = types.ModuleType("shapely")
    shapely._fiss_stub = True
    affinity = types.ModuleType("shapely.affinity")
    affinity.translate = _translate
    affinity.rotate = _rotate
    geometry = types.ModuleType("shapely.geometry")
    geometry.Polygon = Polygon
    shapely.Polygon = Polygon
    shapely.affinity = affinity
    shapely.geometry = geometry
    sys.modules["shapely"] = shapely
    sys.modules["shapely.affinity"] = affinity
    sys.modules["shapely.geometry"] = geometry

    cr = types.ModuleType("commonroad")
    cr_s = types.ModuleType("commonroad.scenario")
    cr_ss = types.ModuleType("commonroad.scenario.scenario")
    cr_so = types.ModuleType("commonroad.scenario.obstacle")
    cr_ss.Scenario = type("Scenario", (), {})
    cr_so.Obstacle = type("Obstacle", (), {})
    cr.scenario = cr_s
    cr_s.scenario = cr_ss
    cr_s.obstacle = cr_so
    sys.modules.update({"commonroad": cr, "commonroad.scenario": cr_s,
                        "commonroad.scenario.scenario": cr_ss, "commonroad.scenario.obstacle": cr_so})


# --------------------------------------------------------------------------- duck-typed inputs
class _ObsState:
    def __init__(self, x, y, th):
        self.position = np.array([x, y])
        self.orientation = th


class RefObstacle:
    """Rectangle obstacle with a dense (x, y, theta) prediction and a validity mask."""

    def __init__(self, xyth: np.ndarray, valid: np.ndarray, length: float, width: float, final_time_step: int):
        self._xyth = xyth
        self._valid = valid
        hl, hw = length / 2.0, width / 2.0
        # CommonRoad Rectangle(length, width).shapely_object: vertices about the origin
        self.obstacle_shape = types.SimpleNamespace(
            shapely_object=Polygon([(-hl, -hw), (-hl, hw), (hl, hw), (hl, -hw), (-hl, -hw)]))
        self.prediction = types.SimpleNamespace(final_time_step=final_time_step)

    def state_at_time(self, t: int):
        if t < 0 or t >= len(self._valid) or not self._valid[t]:
            return None
        return _ObsState(*self._xyth[t])


def make_ref_obstacles(xyth, lw, valid, final_time_step):
    """xyth [M,T,3], lw [M,2], valid [M,T] -> list of RefObstacle."""
    return [RefObstacle(xyth[j], valid[j], lw[j, 0], lw[j, 1], final_time_step) for j in range(len(lw))]


def make_vehicle_params(l=4.569, w=1.844, v_max=41.7, a_max=11.5):
    """Namespace with the attributes vehicle.py:17-46 reads (values: SURVEY 8(c)(ii))."""
    return types.SimpleNamespace(
        l=l, w=w, a=1.1508, b=1.3211, T_f=1.5, T_r=1.5,
        longitudinal=types.SimpleNamespace(v_max=v_max, a_max=a_max),
        steering=types.SimpleNamespace(max=1.023, v_max=0.4, kappa_dot_max=0.4, kappa_dot_dot_max=20.0))


def load_reference():
    """Import and return the reference's planner modules as a namespace."""
    install_stubs()
    from planners.common.scenario.frenet import FrenetState, FrenetTrajectory, State
    from planners.common.vehicle.vehicle import Vehicle
    from planners.fiss_planner import FissPlanner, FissPlannerSettings
    from planners.fiss_plus_planner import FissPlusPlanner, FissPlusPlannerSettings
    from planners.fop_plus_planner import FopPlusPlanner
    from planners.frenet_optimal_planner import FrenetOptimalPlanner, FrenetOptimalPlannerSettings, Stats
    from planners.common.geometry.cubic_spline import CubicSpline1D, CubicSpline2D
    from planners.common.geometry.polynomial import QuarticPolynomial, QuinticPolynomial
    from planners.common.cost.cost_function import CostFunction
    return types.SimpleNamespace(**{k: v for k, v in locals().items()})


# --------------------------------------------------------------------------- closed-loop driver (planning.py)
def install_driver_stubs():
    """Stubs that let the reference's OWN ``planners/benchmark/planning.py`` and
    ``planners/commonroad_interface/global_planner.py`` import and run unmodified:

    * ``commonroad.*`` names they use resolve to the stand-ins of
      ``fiss_plus_planner_b200/planners/commonroad_interface/commonroad_lite.py`` (scenario objects, ``CustomState``,
      ``Trajectory``) -- commonroad-io itself is absent;
    * ``commonroad_route_planner.route_planner.RoutePlanner`` resolves to the stand-in route search of our
      ``global_planner.py`` (the reference's own concatenation / de-dup / heading code then runs on its result);
    * matplotlib, PIL, omegaconf, commonroad_dc, SMP: inert modules (rendering / SMP are out of scope).
    """
    install_stubs()
    import enum
    from fiss_plus_planner_b200.planners.commonroad_interface import commonroad_lite as crl
    from fiss_plus_planner_b200.planners.commonroad_interface import global_planner as gpl
    from fiss_plus_planner_b200.planners.commonroad_interface import vehicle_parameters as vpar

    class _Inert(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return type(name, (), {})

    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None or not isinstance(m, types.ModuleType):
            m = _Inert(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        parent, _, child = name.rpartition(".")
        if parent:
            setattr(mod(parent), child, m)
        return m

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.collections", "PIL", "PIL.Image", "omegaconf",
                 "commonroad.common", "commonroad.common.solution", "commonroad.geometry", "commonroad.geometry.shape",
                 "commonroad.planning", "commonroad.planning.planning_problem", "commonroad.prediction",
                 "commonroad.prediction.prediction", "commonroad.visualization", "commonroad.visualization.mp_renderer",
                 "commonroad.scenario.traffic_sign", "commonroad.scenario.traffic_sign_interpreter",
                 "commonroad_dc", "commonroad_dc.feasibility", "commonroad_dc.feasibility.vehicle_dynamics",
                 "commonroad_route_planner", "commonroad_route_planner.utility",
                 "commonroad_route_planner.utility.visualization",
                 "SMP", "SMP.maneuver_automaton", "SMP.maneuver_automaton.maneuver_automaton", "SMP.motion_planner",
                 "SMP.motion_planner.motion_planner", "SMP.motion_planner.utility"):
        mod(name)
    mod("PIL", Image=mod("PIL.Image"))
    mod("commonroad.common.file_reader", CommonRoadFileReader=crl.CommonRoadFileReader)
    mod("commonroad.scenario.state", CustomState=crl.CustomState)
    mod("commonroad.scenario.trajectory", Trajectory=crl.Trajectory)
    mod("commonroad.scenario.obstacle", DynamicObstacle=crl.DynamicObstacle, ObstacleType=type("ObstacleType", (), {"CAR": "car"}))
    mod("commonroad.geometry.shape", Rectangle=crl.Rectangle)
    mod("commonroad.prediction.prediction", TrajectoryPrediction=crl.TrajectoryPrediction)
    mod("commonroad_dc.feasibility.vehicle_dynamics", VehicleParameterMapping=vpar.VehicleParameterMapping)
    mod("commonroad.common.solution", VehicleType=vpar.VehicleType)

    class RoutePlanner(gpl.RoutePlanner):
        class Backend(enum.Enum):
            NETWORKX = "networkx"
            NETWORKX_REVERSED = "networkx_reversed"
            PRIORITY_QUEUE = "priority_queue"

        def __init__(self, scenario, planning_problem, backend=None):
            super().__init__(scenario, planning_problem)

    mod("commonroad_route_planner.route_planner", RoutePlanner=RoutePlanner)


class RefRectangle:
    """What the reference reads of an obstacle shape (:189): ``shapely_object`` -- here the stand-in polygon."""

    def __init__(self, length, width):
        self.length, self.width = length, width
        hl, hw = length / 2.0, width / 2.0
        self.shapely_object = Polygon([(-hl, -hw), (-hl, hw), (hl, hw), (hl, -hw), (-hl, -hw)])


def load_reference_driver():
    """Import the reference's planning.py (unmodified) and return its ``frenet_optimal_planning``."""
    install_driver_stubs()
    import importlib
    planning = importlib.import_module("planners.benchmark.planning")
    return planning


def attach_shapely_shapes(scenario):
    """Give every obstacle shape of a commonroad_lite scenario the ``shapely_object`` the reference reads."""
    for ob in scenario.static_obstacles + scenario.dynamic_obstacles:
        sh = ob.obstacle_shape
        if not hasattr(sh, "shapely_object"):
            sh.shapely_object = RefRectangle(sh.length, sh.width).shapely_object
    return scenario


def load_reference_waymo():
    """Import the reference's ``planners/waymo_interface/waymo_interface.py`` UNMODIFIED and return the module
    (``convert_waymo_obstacle_to_cr``, :24-76, is what the f-3 goldens execute).  The script's imports resolve to:
    commonroad-io names -> the stand-ins of commonroad_lite (as for the closed-loop driver); ``fiss_planner.*`` -- a
    stale import path that does not exist in the reference tree (:22) -- and matplotlib -> inert modules; everything
    else (``common.vehicle.vehicle``, ``common.scenario.frenet``) is the reference's own code."""
    install_driver_stubs()
    import importlib
    from fiss_plus_planner_b200.planners.commonroad_interface import commonroad_lite as crl
    sys.modules["commonroad.scenario.state"].InitialState = crl.InitialState
    sys.modules["commonroad.scenario.obstacle"].Obstacle = type("Obstacle", (), {})
    stale = {}
    for name in ("fiss_planner", "fiss_planner.fiss_plus_planner"):
        m = types.ModuleType(name)
        m.FissPlusPlannerSettings = type("FissPlusPlannerSettings", (), {})
        m.FissPlusPlanner = type("FissPlusPlanner", (), {})
        stale[name] = sys.modules.get(name)
        sys.modules[name] = m
    sys.modules["fiss_planner"].fiss_plus_planner = sys.modules["fiss_planner.fiss_plus_planner"]
    try:
        return importlib.import_module("planners.waymo_interface.waymo_interface")
    finally:
        for name, old in stale.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old


def table_from_obstacle_objects(obstacles, t_obs):
    """Read a list of CommonRoad-style obstacle objects the way ``has_collision`` does (:173,185-189):
    ``(xyth [M, t_obs, 3], lw [M, 2], valid [M, t_obs], obstacles[0].prediction.final_time_step)``."""
    m = len(obstacles)
    xyth = np.zeros((m, t_obs, 3))
    valid = np.zeros((m, t_obs), dtype=bool)
    lw = np.zeros((m, 2))
    for j, ob in enumerate(obstacles):
        lw[j] = (ob.obstacle_shape.length, ob.obstacle_shape.width)
        for t in range(t_obs):
            st = ob.state_at_time(t)
            if st is not None:
                xyth[j, t] = (st.position[0], st.position[1], st.orientation)
                valid[j, t] = True
    final = int(obstacles[0].prediction.final_time_step) if m else 0
    return xyth, lw, valid, final
