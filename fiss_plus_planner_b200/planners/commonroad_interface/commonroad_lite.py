"""Minimal CommonRoad (2020a XML) reader and scenario objects -- just what the closed-loop driver reads.

The reference's driver (planners/benchmark/planning.py:35-208, :291-314) is written against the third-party
``commonroad-io`` package (2022.3, environment.yml), which is not in the reference tree and not installable
here.  This module restates the small slice of it that ``frenet_optimal_planning`` and the planners touch, with
the same attribute names, so that the driver reads like the reference's:

==============================================  ==================================================================
reference use (planning.py / planners)           here
==============================================  ==================================================================
``CommonRoadFileReader(path).open()`` (:303)     ``CommonRoadFileReader.open() -> (Scenario, PlanningProblemSet)``
``planning_problem_set.planning_problem_dict``   dict id -> ``PlanningProblem``
``scenario.static_obstacles/.dynamic_obstacles`` lists (:65-67)
``scenario.lanelet_network.find_lanelet_by_id``  ``LaneletNetwork`` / ``Lanelet`` (``center_vertices``, ``left_vertices``,
                                                 ``right_vertices``, ``successor``, ``adj_left`` ..., :55-59)
``obstacle.prediction.final_time_step``          ``TrajectoryPrediction`` (:70; frenet_optimal_planner.py:173)
``obstacle.state_at_time(t)`` -> None | state    ``DynamicObstacle.state_at_time`` (frenet_optimal_planner.py:187-189)
``obstacle.obstacle_shape``                      ``Rectangle(length, width)`` (+ ``vertices``)
``planning_problem.initial_state``               ``InitialState`` (position, orientation, velocity, acceleration=0.0)
``planning_problem.goal`` (GoalRegion)           ``state_list[0].has_value("velocity")``, ``lanelets_of_goal_position``,
                                                 ``is_reached(state)`` (:44-56,150)
``CustomState(**kw)`` / ``Trajectory`` (:139-146) ``CustomState`` / ``Trajectory``
==============================================  ==================================================================

Third-party behaviour restated from memory of commonroad-io 2022.3 (no source in the tree; flagged in DESIGN.md):
the centre line is the mean of the bounds; ``state_at_time`` is ``None`` outside ``[initial time, final time]``;
a missing ``<acceleration>`` in the planning problem's initial state reads as 0.0; ``GoalRegion.is_reached`` tests
every attribute the goal state carries (time interval, position inside the goal lanelets' polygons, orientation /
velocity intervals when present).  Host-side set-up code, run once per scenario; nothing here is on the hot path.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET

import numpy as np


# ---------------------------------------------------------------------------------------------- small value types
class Interval(object):
    def __init__(self, start, end):
        self.start = start
        self.end = end

    def contains(self, v) -> bool:
        return self.start <= v <= self.end

    def __repr__(self):
        return "Interval(%r, %r)" % (self.start, self.end)


class CustomState(object):
    """State with free attributes (``commonroad.scenario.state.CustomState``)."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def attributes(self):
        return [k for k in vars(self)]

    def has_value(self, name: str) -> bool:
        return getattr(self, name, None) is not None

    def __repr__(self):
        return "CustomState(%s)" % ", ".join("%s=%r" % kv for kv in vars(self).items())


InitialState = CustomState


class Trajectory(object):
    def __init__(self, initial_time_step: int, state_list: list):
        self.initial_time_step = int(initial_time_step)
        self.state_list = list(state_list)

    @property
    def final_time_step(self) -> int:
        return self.initial_time_step + len(self.state_list) - 1

    def state_at_time_step(self, t: int):
        i = t - self.initial_time_step
        return self.state_list[i] if 0 <= i < len(self.state_list) else None


class Rectangle(object):
    """Axis-aligned rectangle about the origin, ``length`` along x (``commonroad.geometry.shape.Rectangle``)."""

    def __init__(self, length: float, width: float):
        self.length = float(length)
        self.width = float(width)
        hl, hw = self.length / 2.0, self.width / 2.0
        self.vertices = np.array([(-hl, -hw), (-hl, hw), (hl, hw), (hl, -hw), (-hl, -hw)])


class TrajectoryPrediction(object):
    def __init__(self, trajectory: Trajectory, shape: Rectangle):
        self.trajectory = trajectory
        self.shape = shape

    @property
    def initial_time_step(self) -> int:
        return self.trajectory.initial_time_step

    @property
    def final_time_step(self) -> int:
        return self.trajectory.final_time_step


class DynamicObstacle(object):
    def __init__(self, obstacle_id: int, obstacle_type: str, obstacle_shape: Rectangle, initial_state: CustomState,
                 prediction: TrajectoryPrediction = None):
        self.obstacle_id = obstacle_id
        self.obstacle_type = obstacle_type
        self.obstacle_shape = obstacle_shape
        self.initial_state = initial_state
        self.prediction = prediction

    def state_at_time(self, time_step: int):
        if time_step == self.initial_state.time_step:
            return self.initial_state
        if self.prediction is None:
            return None
        return self.prediction.trajectory.state_at_time_step(time_step)

    def dense_table(self, t_obs: int):
        """(xyth [t_obs, 3], valid [t_obs]) -- the rows ``fiss_set_obstacles`` takes."""
        xyth = np.zeros((t_obs, 3))
        valid = np.zeros(t_obs, dtype=np.uint8)
        for t in range(t_obs):
            st = self.state_at_time(t)
            if st is not None:
                xyth[t] = (st.position[0], st.position[1], st.orientation)
                valid[t] = 1
        return xyth, valid


class StaticObstacle(object):
    """Present at every time step with its initial pose; ``prediction`` is None as in commonroad-io
    (a static obstacle in slot 0 makes the reference raise at frenet_optimal_planner.py:173 -- none in data/demo)."""

    def __init__(self, obstacle_id: int, obstacle_type: str, obstacle_shape: Rectangle, initial_state: CustomState):
        self.obstacle_id = obstacle_id
        self.obstacle_type = obstacle_type
        self.obstacle_shape = obstacle_shape
        self.initial_state = initial_state
        self.prediction = None

    def state_at_time(self, time_step: int):
        return self.initial_state


# ---------------------------------------------------------------------------------------------- road network
def _point_in_ring(pt, ring: np.ndarray) -> bool:
    """Even-odd rule with the boundary counted as inside (closed set)."""
    x, y = float(pt[0]), float(pt[1])
    x0, y0 = ring[:, 0], ring[:, 1]
    x1, y1 = np.roll(x0, -1), np.roll(y0, -1)
    # on an edge?
    cross = (x1 - x0) * (y - y0) - (y1 - y0) * (x - x0)
    dot = (x - x0) * (x - x1) + (y - y0) * (y - y1)
    if np.any((np.abs(cross) <= 1e-12 * (1.0 + np.hypot(x1 - x0, y1 - y0))) & (dot <= 0.0)):
        return True
    straddle = (y0 > y) != (y1 > y)
    with np.errstate(divide="ignore", invalid="ignore"):
        xi = x0 + (y - y0) * (x1 - x0) / (y1 - y0)
    return bool(np.count_nonzero(straddle & (x < xi)) % 2)


class Lanelet(object):
    def __init__(self, lanelet_id: int, left_vertices, right_vertices, predecessor, successor, adj_left=None,
                 adj_left_same_direction=None, adj_right=None, adj_right_same_direction=None, lanelet_type=()):
        self.lanelet_id = int(lanelet_id)
        self.left_vertices = np.asarray(left_vertices, dtype=np.float64)
        self.right_vertices = np.asarray(right_vertices, dtype=np.float64)
        self.center_vertices = 0.5 * (self.left_vertices + self.right_vertices)
        self.predecessor = list(predecessor)
        self.successor = list(successor)
        self.adj_left = adj_left
        self.adj_left_same_direction = adj_left_same_direction
        self.adj_right = adj_right
        self.adj_right_same_direction = adj_right_same_direction
        self.lanelet_type = set(lanelet_type)
        seg = np.hypot(*np.diff(self.center_vertices, axis=0).T)
        self.distance = np.concatenate(([0.0], np.cumsum(seg)))   # arc length along the centre line
        self._ring = np.vstack((self.right_vertices, self.left_vertices[::-1]))

    @property
    def polygon_vertices(self) -> np.ndarray:
        return self._ring

    def contains_point(self, pt) -> bool:
        return _point_in_ring(pt, self._ring)

    def orientation_at_start(self) -> float:
        d = self.center_vertices[1] - self.center_vertices[0]
        return float(np.arctan2(d[1], d[0]))


class LaneletNetwork(object):
    def __init__(self, lanelets):
        self.lanelets = list(lanelets)
        self._by_id = {l.lanelet_id: l for l in self.lanelets}

    def find_lanelet_by_id(self, lanelet_id: int) -> Lanelet:
        return self._by_id.get(int(lanelet_id))

    def find_lanelet_by_position(self, point_list):
        return [[l.lanelet_id for l in self.lanelets if l.contains_point(p)] for p in point_list]


class Scenario(object):
    def __init__(self, dt: float, scenario_id: str, lanelet_network: LaneletNetwork, static_obstacles, dynamic_obstacles):
        self.dt = float(dt)
        self.scenario_id = scenario_id
        self.lanelet_network = lanelet_network
        self.static_obstacles = list(static_obstacles)
        self.dynamic_obstacles = list(dynamic_obstacles)

    @property
    def obstacles(self):
        return self.static_obstacles + self.dynamic_obstacles


# ---------------------------------------------------------------------------------------------- planning problem
class GoalRegion(object):
    def __init__(self, state_list, lanelets_of_goal_position=None, lanelet_network: LaneletNetwork = None):
        self.state_list = list(state_list)
        self.lanelets_of_goal_position = lanelets_of_goal_position
        self._net = lanelet_network

    @staticmethod
    def _in(value, spec) -> bool:
        return spec.contains(value) if isinstance(spec, Interval) else value == spec

    def is_reached(self, state) -> bool:
        """True when ``state`` satisfies every attribute of at least one goal state."""
        for gi, goal in enumerate(self.state_list):
            ok = True
            if goal.has_value("time_step"):
                ok = ok and self._in(state.time_step, goal.time_step)
            if goal.has_value("position"):
                ok = ok and any(ring_owner.contains_point(state.position) for ring_owner in goal.position)
            if goal.has_value("orientation") and hasattr(state, "orientation"):
                a = (state.orientation - goal.orientation.start) % (2.0 * np.pi)
                ok = ok and a <= (goal.orientation.end - goal.orientation.start)
            if goal.has_value("velocity") and hasattr(state, "velocity"):
                ok = ok and self._in(state.velocity, goal.velocity)
            if ok:
                return True
        return False


class PlanningProblem(object):
    def __init__(self, planning_problem_id: int, initial_state: CustomState, goal_region: GoalRegion):
        self.planning_problem_id = planning_problem_id
        self.initial_state = initial_state
        self.goal = goal_region


class PlanningProblemSet(object):
    def __init__(self, problems):
        self.planning_problem_dict = {p.planning_problem_id: p for p in problems}


# ---------------------------------------------------------------------------------------------- XML
def _exact_or_interval(node, cast=float):
    if node is None:
        return None
    ex = node.find("exact")
    if ex is not None:
        return cast(ex.text)
    lo, hi = node.find("intervalStart"), node.find("intervalEnd")
    if lo is not None and hi is not None:
        return Interval(cast(lo.text), cast(hi.text))
    return None


def _points(node):
    return [(float(p.find("x").text), float(p.find("y").text)) for p in node.findall("point")]


def _state(node) -> CustomState:
    kw = {}
    pos = node.find("position")
    if pos is not None and pos.find("point") is not None:
        kw["position"] = np.array(_points(pos)[0])
    for tag, name, cast in (("orientation", "orientation", float), ("time", "time_step", int), ("velocity", "velocity", float),
                            ("acceleration", "acceleration", float), ("yawRate", "yaw_rate", float),
                            ("slipAngle", "slip_angle", float)):
        v = _exact_or_interval(node.find(tag), cast)
        if v is not None:
            kw[name] = v
    return CustomState(**kw)


class CommonRoadFileReader(object):
    """``CommonRoadFileReader(filename).open()`` (planning.py:303) for 2020a files with rectangle obstacles."""

    def __init__(self, filename: str):
        self._filename = filename

    def open(self, lanelet_assignment: bool = False):
        root = ET.parse(self._filename).getroot()
        lanelets = []
        for ln in root.findall("lanelet"):
            def ref(tag):
                n = ln.find(tag)
                return (None, None) if n is None else (int(n.get("ref")), n.get("drivingDir") == "same")
            al, als = ref("adjacentLeft")
            ar, ars = ref("adjacentRight")
            lanelets.append(Lanelet(
                int(ln.get("id")), _points(ln.find("leftBound")), _points(ln.find("rightBound")),
                [int(n.get("ref")) for n in ln.findall("predecessor")], [int(n.get("ref")) for n in ln.findall("successor")],
                al, als, ar, ars, [n.text for n in ln.findall("laneletType")]))
        net = LaneletNetwork(lanelets)

        def shape_of(node):
            sh = node.find("shape")
            rect = sh.find("rectangle") if sh is not None else None
            if rect is None:
                raise NotImplementedError("only <rectangle> obstacle shapes are supported (all of data/demo)")
            return Rectangle(float(rect.find("length").text), float(rect.find("width").text))

        dynamic = []
        for ob in root.findall("dynamicObstacle"):
            shape = shape_of(ob)
            init = _state(ob.find("initialState"))
            traj = ob.find("trajectory")
            pred = None
            if traj is not None:
                states = [_state(s) for s in traj.findall("state")]
                if states:
                    pred = TrajectoryPrediction(Trajectory(states[0].time_step, states), shape)
            dynamic.append(DynamicObstacle(int(ob.get("id")), ob.findtext("type"), shape, init, pred))
        static = [StaticObstacle(int(ob.get("id")), ob.findtext("type"), shape_of(ob), _state(ob.find("initialState")))
                  for ob in root.findall("staticObstacle")]
        scenario = Scenario(float(root.get("timeStepSize")), root.get("benchmarkID"), net, static, dynamic)

        problems = []
        for pp in root.findall("planningProblem"):
            init = _state(pp.find("initialState"))
            if not init.has_value("acceleration"):
                init.acceleration = 0.0
            goal_states, lanelets_of_goal = [], {}
            for gi, gs in enumerate(pp.findall("goalState")):
                st = _state(gs)
                pos = gs.find("position")
                if pos is not None:
                    ids = [int(n.get("ref")) for n in pos.findall("lanelet")]
                    if ids:
                        lanelets_of_goal[gi] = ids
                        st.position = [net.find_lanelet_by_id(i) for i in ids]
                    elif pos.find("rectangle") is not None or pos.find("circle") is not None or pos.find("polygon") is not None:
                        raise NotImplementedError("goal positions other than lanelet references are not supported")
                goal_states.append(st)
            problems.append(PlanningProblem(int(pp.get("id")), init,
                                            GoalRegion(goal_states, lanelets_of_goal or None, net)))
        return scenario, PlanningProblemSet(problems)


# ---------------------------------------------------------------------------------------------- writer (tests, fixtures)
def write_commonroad_xml(path: str, scenario: Scenario, problem: PlanningProblem, benchmark_id: str = None):
    """Write the subset this reader understands (round-trip fixture generator for the tests)."""
    def pt(parent, xy):
        p = ET.SubElement(parent, "point")
        ET.SubElement(p, "x").text = repr(float(xy[0]))
        ET.SubElement(p, "y").text = repr(float(xy[1]))

    def exact(parent, tag, v):
        ET.SubElement(ET.SubElement(parent, tag), "exact").text = repr(v)

    def state(parent, tag, st):
        n = ET.SubElement(parent, tag)
        pt(ET.SubElement(n, "position"), st.position)
        exact(n, "orientation", float(st.orientation))
        exact(n, "time", int(st.time_step))
        exact(n, "velocity", float(getattr(st, "velocity", 0.0)))
        if getattr(st, "acceleration", None) is not None:
            exact(n, "acceleration", float(st.acceleration))
        return n

    root = ET.Element("commonRoad", timeStepSize=repr(scenario.dt), commonRoadVersion="2020a",
                      benchmarkID=benchmark_id or scenario.scenario_id or "ZAM_Lite-1_1_T-1")
    for l in scenario.lanelet_network.lanelets:
        ln = ET.SubElement(root, "lanelet", id=str(l.lanelet_id))
        for tag, verts in (("leftBound", l.left_vertices), ("rightBound", l.right_vertices)):
            b = ET.SubElement(ln, tag)
            for xy in verts:
                pt(b, xy)
        for p in l.predecessor:
            ET.SubElement(ln, "predecessor", ref=str(p))
        for s in l.successor:
            ET.SubElement(ln, "successor", ref=str(s))
        if l.adj_left is not None:
            ET.SubElement(ln, "adjacentLeft", ref=str(l.adj_left), drivingDir="same" if l.adj_left_same_direction else "opposite")
        if l.adj_right is not None:
            ET.SubElement(ln, "adjacentRight", ref=str(l.adj_right), drivingDir="same" if l.adj_right_same_direction else "opposite")
    for ob in scenario.dynamic_obstacles:
        n = ET.SubElement(root, "dynamicObstacle", id=str(ob.obstacle_id))
        ET.SubElement(n, "type").text = ob.obstacle_type or "car"
        r = ET.SubElement(ET.SubElement(n, "shape"), "rectangle")
        ET.SubElement(r, "length").text = repr(ob.obstacle_shape.length)
        ET.SubElement(r, "width").text = repr(ob.obstacle_shape.width)
        state(n, "initialState", ob.initial_state)
        if ob.prediction is not None:
            tr = ET.SubElement(n, "trajectory")
            for st in ob.prediction.trajectory.state_list:
                state(tr, "state", st)
    pp = ET.SubElement(root, "planningProblem", id=str(problem.planning_problem_id))
    state(pp, "initialState", problem.initial_state)
    for gi, gs in enumerate(problem.goal.state_list):
        g = ET.SubElement(pp, "goalState")
        pos = ET.SubElement(g, "position")
        for lid in (problem.goal.lanelets_of_goal_position or {}).get(gi, []):
            ET.SubElement(pos, "lanelet", ref=str(lid))
        if gs.has_value("time_step"):
            t = ET.SubElement(g, "time")
            ET.SubElement(t, "intervalStart").text = str(gs.time_step.start)
            ET.SubElement(t, "intervalEnd").text = str(gs.time_step.end)
        if gs.has_value("velocity"):
            v = ET.SubElement(g, "velocity")
            ET.SubElement(v, "intervalStart").text = repr(gs.velocity.start)
            ET.SubElement(v, "intervalEnd").text = repr(gs.velocity.end)
    ET.ElementTree(root).write(path, encoding="UTF-8", xml_declaration=True)
