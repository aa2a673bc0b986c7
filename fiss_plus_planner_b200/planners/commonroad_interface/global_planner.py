"""Global route -> reference centre line (reference: planners/commonroad_interface/global_planner.py:9-106).

``GlobalPlanner().plan_global_route(scenario, planning_problem).concat_centerline`` is the ``[n, 4] = (x, y, yaw,
width)`` polyline that ``generate_frenet_frame`` fits the spline to (planning.py:37-39,101).  The reference
delegates the route search to the third-party ``commonroad_route_planner`` (2022.3; not in the tree, not
installable here).  Its published behaviour for the default ``NETWORKX_REVERSED`` backend is restated below as
a stand-in [memory-flagged, see DESIGN.md]: lanelets are nodes; a successor edge weighs the lanelet's centre-line
length, a same-direction lane-change edge weighs 1; the route is a shortest path searched backwards from the goal
lanelet to a start lanelet (so lane changes happen as late as possible), and among the candidate routes the one
whose first lanelet is best aligned with the initial orientation is taken.  Everything after the search -- the
extra successor lanelet, concatenation, first-occurrence de-duplication, headings, widths -- follows the
reference's own code (:68-96).  Set-up code, once per scenario, host NumPy.
"""
from __future__ import annotations

import heapq
import math

import numpy as np


class GlobalPlan(object):
    def __init__(self):
        self.lanelets = None
        self.lanelet_centerlines = None
        self.concat_centerline = None
        self.speed_limits = None
        self.required_speeds = None


class Route(object):
    def __init__(self, list_ids_lanelets):
        self.list_ids_lanelets = list(list_ids_lanelets)


def _angle_diff(a: float, b: float) -> float:
    d = (a - b + math.pi) % (2.0 * math.pi) - math.pi
    return abs(d)


class RoutePlanner(object):
    """Shortest lanelet sequence start -> goal on the reversed lanelet graph."""

    LANE_CHANGE_WEIGHT = 1.0

    def __init__(self, scenario, planning_problem):
        self.net = scenario.lanelet_network
        self.problem = planning_problem
        init = planning_problem.initial_state
        self.ids_start = self.net.find_lanelet_by_position([init.position])[0]
        goal = planning_problem.goal
        self.ids_goal = []
        if goal.lanelets_of_goal_position:
            for ids in goal.lanelets_of_goal_position.values():
                self.ids_goal.extend(ids)
        if not self.ids_start:
            raise RuntimeError("initial position is not on any lanelet")
        if not self.ids_goal:
            raise RuntimeError("goal carries no lanelet reference")
        # reversed graph: edge v -> u for every forward edge u -> v
        self.rev = {l.lanelet_id: [] for l in self.net.lanelets}
        for l in self.net.lanelets:
            for s in l.successor:
                if s in self.rev:
                    self.rev[s].append((l.lanelet_id, float(l.distance[-1])))
            for adj, same in ((l.adj_left, l.adj_left_same_direction), (l.adj_right, l.adj_right_same_direction)):
                if adj is not None and same and adj in self.rev:
                    self.rev[adj].append((l.lanelet_id, self.LANE_CHANGE_WEIGHT))

    def _shortest(self, id_goal: int, id_start: int):
        dist, prev, heap = {id_goal: 0.0}, {}, [(0.0, id_goal)]
        while heap:
            d, u = heapq.heappop(heap)
            if d > dist.get(u, math.inf):
                continue
            if u == id_start:
                break
            for v, w in self.rev[u]:
                nd = d + w
                if nd < dist.get(v, math.inf):
                    dist[v], prev[v] = nd, u
                    heapq.heappush(heap, (nd, v))
        if id_start not in dist:
            return None
        path = [id_start]
        while path[-1] != id_goal:
            path.append(prev[path[-1]])
        return path

    def plan_routes(self):
        routes = []
        for g in self.ids_goal:
            for s in self.ids_start:
                p = self._shortest(g, s)
                if p is not None:
                    routes.append(Route(p))
        if not routes:
            raise RuntimeError("no route from the initial lanelet to the goal lanelet")
        self.routes = routes
        return self

    def retrieve_first_route(self) -> Route:
        return self.routes[0]

    def retrieve_best_route_by_orientation(self) -> Route:
        yaw0 = float(self.problem.initial_state.orientation)
        return min(self.routes, key=lambda r: _angle_diff(
            self.net.find_lanelet_by_id(r.list_ids_lanelets[0]).orientation_at_start(), yaw0))


class GlobalPlanner(object):
    def plan_global_route(self, scenario, planning_problem, method: str = "NETWORKX_REVERSED",
                          plan_all_routes: bool = False, view_route: bool = False) -> GlobalPlan:
        holder = RoutePlanner(scenario, planning_problem).plan_routes()
        route = holder.retrieve_first_route() if plan_all_routes else holder.retrieve_best_route_by_orientation()

        net = scenario.lanelet_network
        plan = GlobalPlan()
        plan.lanelets = [net.find_lanelet_by_id(i) for i in route.list_ids_lanelets]
        # one lanelet beyond the goal when it has a successor (:68-75)
        tail = plan.lanelets[-1].successor
        if len(tail) != 0:
            plan.lanelets.append(net.find_lanelet_by_id(tail[0]))
        plan.lanelet_centerlines = np.array([l.center_vertices for l in plan.lanelets], dtype=object)

        # joined centre line, duplicates dropped keeping first occurrences in route order (:78-82)
        pts = np.array(np.concatenate(plan.lanelet_centerlines), dtype=float)
        _, first = np.unique(pts, return_index=True, axis=0)
        keep = np.sort(first)
        pts = pts[keep]
        # heading of each segment, the last one repeated (:85-88)
        yaws = np.arctan2(np.diff(pts[:, 1]), np.diff(pts[:, 0]))
        yaws = np.append(yaws, yaws[-1])
        # lane width at every vertex (:91-94)
        # (per-vertex 2-norms one at a time, like the reference: the vectorised axis=1 norm rounds differently by 1 ulp)
        widths = np.concatenate([np.array([np.linalg.norm(l.left_vertices[i] - l.right_vertices[i])
                                           for i in range(len(l.left_vertices))]) for l in plan.lanelets])[keep]
        plan.concat_centerline = np.column_stack((pts, yaws, widths))
        if view_route:
            print("Global Planning Results:")
            print("Passing through:", len(plan.lanelets), "lanelets")
            print("Contains:", plan.concat_centerline.shape, "lane points")
        return plan
