"""FissPlusPlanner on the B200 lattice engine (reference: planners/fiss_plus_planner.py:15-326).

FISS+ = FISS with (i) a six-neighbour frontier search instead of the gradient walk and (ii) an
off-lattice refinement of the coarse winner: up to ``max_refine_iters`` rounds of six clipped
neighbours at +-resolution plus one normalised gradient step with the resolution halved.

Device use per cycle: one launch for the whole coarse grid (inherited from ``FissPlanner``), then
per refinement round one call for the six neighbours and one for the gradient step -- each call
returns cost AND both feasibility masks, so the final best-first validation of the refined
candidates is again pure bookkeeping -- and one call for the winner's arrays.
"""
from __future__ import annotations

import heapq
import time

import numpy as np

from fiss_plus_planner_b200 import _shim
from fiss_plus_planner_b200.engine import end_state_table
from fiss_plus_planner_b200.planners.common.scenario.frenet import FrenetState, FrenetTrajectory
from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
from fiss_plus_planner_b200.planners.fiss_planner import FissPlanner, FissPlannerSettings
from fiss_plus_planner_b200.planners.frenet_optimal_planner import Stats


class FissPlusPlannerSettings(FissPlannerSettings):
    def __init__(self, num_width: int = 5, num_speed: int = 5, num_t: int = 5, refine_iters: int = 3):
        super().__init__(num_width, num_speed, num_t)
        self.refine_trajectory = True
        self.max_refine_iters = refine_iters
        self.has_time_limit = False
        self.time_limit = 0.5
        self.decaying_factor = 0.5


class _Refined(object):
    """An off-lattice candidate in the ``refined_trajs`` queue: cost-ordered (frenet.py:156)."""
    __slots__ = ("cost_final", "row", "flags")

    def __init__(self, cost, row, flags):
        self.cost_final, self.row, self.flags = cost, row, flags

    def __lt__(self, other):
        return self.cost_final < other.cost_final


class FissPlusPlanner(FissPlanner):
    def __init__(self, planner_settings: FissPlusPlannerSettings, ego_vehicle: Vehicle, scenario=None,
                 device: int = 0, engine=None):
        super().__init__(planner_settings, ego_vehicle, scenario, device=device, engine=engine)
        self.frontier_idxs = []
        self.refined_trajs = []

    def explore_neighbors(self, idx: np.ndarray) -> tuple:
        """Touch the (up to) six axis neighbours; new ones that do not cost more than the centre go on
        the frontier; report the cheapest of centre + neighbours (fiss_plus_planner.py:30-59)."""
        _, centre = self.generate_trajectory(idx)
        lo = centre
        best = np.array(idx)
        is_min = True
        for dim in range(3):
            for step in (-1, +1):
                if (step < 0 and idx[dim] < 1) or (step > 0 and idx[dim] >= self.sizes[dim] - 1):
                    continue
                nb = np.array(idx)
                nb[dim] += step
                is_new, cost = self.generate_trajectory(nb)
                if is_new and cost <= centre:
                    heapq.heappush(self.frontier_idxs, (cost, nb))
                if cost <= lo:
                    lo, best, is_min = cost, nb, False
        return is_min, best

    def plan(self, frenet_state: FrenetState, max_target_speed: float, obstacles: list, time_step_now: int = 0) -> FrenetTrajectory:
        t_start = time.time()
        self.frontier_idxs = []
        self.refined_trajs = []
        self._begin_cycle(frenet_state, max_target_speed, obstacles, time_step_now)
        found = False
        while True:
            self.stats.num_iter += 1
            if not self.candidate_trajs:
                best_idx = self.find_initial_guess()
                if best_idx is None:
                    break
            else:
                best_idx = self.candidate_trajs[0][1]
            while True:
                _, best_idx = self.explore_neighbors(best_idx)
                if not self.frontier_idxs:
                    break
                _, best_idx = heapq.heappop(self.frontier_idxs)
            if not self.candidate_trajs:
                break
            _, idx = heapq.heappop(self.candidate_trajs)
            if self._validate_flags(int(self._flags[self._lin(idx)])):
                self._accept_lattice_winner(idx)
                found = True
                break

        if found and self.settings.refine_trajectory:
            time_left = self.settings.time_limit - (time.time() - t_start)
            if not self.settings.has_time_limit or time_left > 0.0:
                refined = self.refine_solution(self.best_traj, time_left, obstacles, time_step_now)
                if refined is not None:
                    self.best_traj = refined
        self._finish_cycle()
        return self.best_traj

    # -- refinement ------------------------------------------------------------------------------
    def _evaluate_end_states(self, xs: np.ndarray):
        """generate_trajectory_by_end_state (fiss_plus_planner.py:172-205) for a few (d, v, T) rows in
        one device call; every row counts as a generation and joins the refined queue."""
        rows = end_state_table(xs[:, 0], xs[:, 1], xs[:, 2], self.settings.tick_t)
        out = self.engine.eval_end_states(self._ego6, rows, self._prm, want_records=False)
        for r in range(len(rows)):
            self.stats.num_trajs_generated += 1
            heapq.heappush(self.refined_trajs, _Refined(out["cost"][r], rows[r], int(out["flags"][r])))
        return out["cost"]

    def generate_trajectory_by_end_state(self, end_state: FrenetState) -> float:
        return self._evaluate_end_states(np.array([[end_state.d, end_state.s_d, end_state.t]]))[0]

    def gradient_decent(self, J: float, x: np.ndarray, resolutions: np.ndarray, decaying_factor: float) -> tuple:
        """Central differences over six clipped neighbours, resolution decay, one normalised step
        (fiss_plus_planner.py:207-277).  ``resolutions`` is scaled IN PLACE, as in the reference,
        where it aliases ``self.sampling_res`` (:261,282)."""
        nbrs = np.empty((6, 3))
        for dim in range(3):
            lo = np.array(x)
            lo[dim] -= resolutions[dim]
            hi = np.array(x)
            hi[dim] += resolutions[dim]
            nbrs[2 * dim] = np.clip(lo, self.sampling_min, self.sampling_max)
            nbrs[2 * dim + 1] = np.clip(hi, self.sampling_min, self.sampling_max)
        J6 = self._evaluate_end_states(nbrs)            # same generation order: l, r per dimension
        d_J = np.array([J6[1] - J6[0], J6[3] - J6[2], J6[5] - J6[4]])
        d_x = np.array([nbrs[1, 0] - nbrs[0, 0], nbrs[3, 1] - nbrs[2, 1], nbrs[5, 2] - nbrs[4, 2]])
        grad = d_J / d_x
        resolutions *= decaying_factor
        x_new = x - resolutions * grad / np.linalg.norm(grad)
        x_new_clipped = np.clip(x_new, self.sampling_min, self.sampling_max)
        J_new = self._evaluate_end_states(x_new_clipped[None])[0]
        return True, J_new, x_new_clipped, resolutions

    def refine_solution(self, traj: FrenetTrajectory, time_limit: float, obstacles: list, time_step_now: int) -> FrenetTrajectory:
        """fiss_plus_planner.py:279-326, including its wall-clock break (checked even when
        ``has_time_limit`` is False; the budget is what the coarse search left of ``time_limit``)."""
        t_start = time.time()
        resolutions = self.sampling_res
        J_new = traj.cost_final
        x = np.array([traj.end_state.d, traj.end_state.s_d, traj.end_state.t])
        x0 = np.array(x)
        for _ in range(self.settings.max_refine_iters):
            valid, J_new, x, resolutions = self.gradient_decent(J_new, x, resolutions, self.settings.decaying_factor)
            if not valid:
                break
            if time.time() - t_start >= time_limit:
                break
        # Refined candidates are costed by the list kernel, the coarse winner by the lattice kernel: the two agree
        # to ~1e-16, but the reference's `>` sees an EXACT tie whenever a clipped neighbour coincides with the coarse
        # end state (x on the boundary of the sampling box) and then returns the refined copy (idx = [-1, -1, -1]).
        # Near a tie, compare against the coarse end state costed by the same kernel as the candidates.
        thr = traj.cost_final
        thr_same_kernel = None
        while self.refined_trajs:
            cand = heapq.heappop(self.refined_trajs)
            if abs(cand.cost_final - traj.cost_final) <= 1e-9 * abs(traj.cost_final):
                if thr_same_kernel is None:
                    row = end_state_table(x0[:1], x0[1:2], x0[2:3], self.settings.tick_t)
                    thr_same_kernel = float(self.engine.eval_end_states(self._ego6, row, self._prm)["cost"][0])
                thr = thr_same_kernel
            if cand.cost_final > thr:
                break
            if self._validate_flags(cand.flags):
                out = self._fetch_trajectory(cand.row, cand.cost_final)
                out.end_state = FrenetState(t=cand.row[2], s=0.0, s_d=cand.row[1], s_dd=0.0, s_ddd=0.0,
                                            d=cand.row[0], d_d=0.0, d_dd=0.0, d_ddd=0.0)
                return out
        return None
