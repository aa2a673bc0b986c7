"""FopPlusPlanner on the B200 lattice engine (reference: planners/fop_plus_planner.py:11-40).

The reference samples and converts the whole lattice, pushes every candidate into a
``PriorityQueue`` ordered by ``cost_final`` and validates lazily in cost order.  Here the device
evaluates the whole lattice *including* both masks in one launch; the host then replays the
queue -- a ``heapq`` over cost-only-ordered keys pushed in the reference's enumeration order, so
ties break exactly as ``queue.PriorityQueue`` would -- to find the first feasible pop and to
reproduce the Stats counters (``num_iter = validated = collision_checks = #pops``, :30-35).
"""
from __future__ import annotations

import heapq

import numpy as np

from fiss_plus_planner_b200 import _shim
from fiss_plus_planner_b200.planners.common.scenario.frenet import FrenetState, FrenetTrajectory
from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
from fiss_plus_planner_b200.planners.frenet_optimal_planner import (CandidateBundle, FrenetOptimalPlanner,
                                                                    FrenetOptimalPlannerSettings, Stats)


class _CostKey(object):
    """Heap entry ordered by cost only, like ``FrenetTrajectory.__lt__`` (frenet.py:156)."""
    __slots__ = ("cost", "seq")

    def __init__(self, cost, seq):
        self.cost = cost
        self.seq = seq

    def __lt__(self, other):
        return self.cost < other.cost


class FopPlusPlanner(FrenetOptimalPlanner):
    def __init__(self, planner_settings: FrenetOptimalPlannerSettings, ego_vehicle: Vehicle, scenario=None,
                 device: int = 0, engine=None):
        super().__init__(planner_settings, ego_vehicle, scenario, device=device, engine=engine)
        self.candidate_trajs = []

    def plan(self, frenet_state: FrenetState, max_target_speed: float, obstacles: list, time_step_now: int = 0) -> FrenetTrajectory:
        self.stats = Stats()
        self.settings.highest_speed = max_target_speed
        self._upload_obstacles(obstacles)
        end = self._end_states()
        prm = self._params(time_step_now)
        ego6 = frenet_state.as_ego6()
        grid = self._lattice_grid()
        out = self.engine.plan_grid(ego6, grid, prm, want_records=True, want_volume=True, out=self._plan_outputs(grid))
        cost, flags = out["cost"][0], out["flags"][0]
        self.stats.num_trajs_generated = len(end)
        self.all_trajs.append(self._bundle(ego6, end, prm, cost, flags))

        heap = []
        for seq in range(len(end)):
            heapq.heappush(heap, _CostKey(cost[seq], seq))
        self.candidate_trajs = heap
        infeasible = (flags & _shim.FLAG_INFEASIBLE_MASK) != 0
        while heap:
            self.stats.num_iter += 1
            key = heapq.heappop(heap)
            self.stats.num_trajs_validated += 1
            self.stats.num_collison_checks += 1
            if not infeasible[key.seq]:
                if key.seq == int(out["best_idx"][0]):
                    rec, meta = out["records"][0], out["meta"][0]
                else:   # an exact cost tie popped in heap order: fetch that candidate's record instead
                    one = self.engine.eval_end_states(ego6, end[key.seq:key.seq + 1], prm, want_records=True)
                    rec = one["records"][0]
                    meta = (int(end[key.seq, 3]), int((one["flags"][0] >> _shim.FLAG_NCART_SHIFT) & _shim.FLAG_NCART_MASK))
                traj = self._trajectory_from_record(rec, meta, cost[key.seq])
                traj.lattice_index = key.seq
                self.best_traj = traj
                return self.best_traj
        return None
