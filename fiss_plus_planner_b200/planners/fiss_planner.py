"""FissPlanner on the B200 lattice engine (reference: planners/fiss_planner.py:13-269).

FISS explores the 3-D index grid ``[i_d][j_v][k_t]`` lazily: a cheap cost *estimate* picks the
start point, finite-difference "gradients" over neighbouring lattice points walk towards lower
cost, and only the points it touches are ever generated.  That search is sequential, data-dependent
and tiny, so it stays on the host -- but every number it consumes comes from the GPU: at the start
of ``plan()`` ONE launch evaluates the whole grid (cost and both feasibility masks; 270 candidates
are ~10 us of B200 time), and ``generate_trajectory`` becomes a lookup into that cost volume plus
the reference's lazy bookkeeping (``is_generated``, ``Stats.num_trajs_generated``, the
``(cost, idx)`` priority queue).  Indices, costs, queue order and counters therefore match the
reference step by step; only the winner's arrays are fetched (one more call, full record).
"""
from __future__ import annotations

import heapq

import numpy as np

from fiss_plus_planner_b200 import _shim
from fiss_plus_planner_b200.engine import fiss_grid, fiss_lattice
from fiss_plus_planner_b200.planners.common.scenario.frenet import FrenetState, FrenetTrajectory
from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
from fiss_plus_planner_b200.planners.frenet_optimal_planner import (CandidateBundle, FrenetOptimalPlanner,
                                                                    FrenetOptimalPlannerSettings, Stats)


class FissPlannerSettings(FrenetOptimalPlannerSettings):
    def __init__(self, num_width: int = 5, num_speed: int = 5, num_t: int = 5):
        super().__init__(num_width, num_speed, num_t)
        self.w_heuristic = 10.0
        self.vis_all_candidates = False


class _GridView(object):
    """``trajs_3d[i][j][k]`` -> placeholder / generated ``FrenetTrajectory`` built on demand."""

    def __init__(self, planner, prefix=()):
        self._p, self._prefix = planner, prefix

    def __len__(self):
        return int(self._p.sizes[len(self._prefix)])

    def __getitem__(self, i):
        idx = self._prefix + (int(i),)
        if len(idx) < 3:
            return _GridView(self._p, idx)
        return self._p._placeholder(idx)


class FissPlanner(FrenetOptimalPlanner):
    def __init__(self, planner_settings: FissPlannerSettings, ego_vehicle: Vehicle, scenario=None,
                 device: int = 0, engine=None):
        super().__init__(planner_settings, ego_vehicle, scenario, device=device, engine=engine)
        self.sampling_res = np.empty(3)
        self.sampling_min = np.empty(3)
        self.sampling_max = np.empty(3)
        self.candidate_trajs = []          # heap of (cost_final, idx) -- PriorityQueue in the reference
        self.trajs_3d = []
        self.sizes = None
        self.start_state = None
        self.prev_best_idx = None
        self.trajs_per_timestep = []
        # device-computed volumes for the current cycle
        self._table = None
        self._fgrid = None
        self._cost = None
        self._flags = None
        self._generated = None
        self._cost_est = None
        self._cost_heu = None
        self._ego6 = None
        self._prm = None

    # -- lattice ---------------------------------------------------------------------------------
    def sample_end_frenet_states(self):
        """Grid of end states + cost estimates (fiss_planner.py:33-99), as arrays.

        ``cost_est = d^2 / max(left^2, right^2) + (1 - (T - min_t)/(max_t - min_t))
        + (v_hi - v)^2 / (v_hi - v_lo)^2  [+ w_heuristic * |idx - prev_best_idx|^2 / (nd^2+nv^2+nt^2)]``
        evaluated with the reference's operation order so that the ``<=`` scan of
        ``find_initial_guess`` sees the same ties."""
        st = self.settings
        key = (st.max_road_width, self.vehicle.w, st.num_width, st.min_t, st.max_t, st.num_t, st.lowest_speed,
               st.highest_speed, st.num_speed, st.tick_t)
        if key != getattr(self, "_fiss_lattice_key", None):   # the lattice only moves when the settings do
            self._fiss_lattice = fiss_lattice(st, self.vehicle.w)
            self._fgrid = fiss_grid(st, self.vehicle.w)      # same points, as axes + [i_d][j_v][k_t] numbering
            self._fiss_lattice_key = key
        table, ds, vs, ts, res = self._fiss_lattice
        sw = st.max_road_width - self.vehicle.w + 0.3
        left, right = -sw / 2, sw / 2
        self.sampling_min[:] = (left, st.lowest_speed, st.min_t)
        self.sampling_max[:] = (right, st.highest_speed, st.max_t)
        self.sampling_res[:] = res
        lat_norm = max(np.power(left, 2), np.power(right, 2))
        est_lat = np.power(ds, 2) / lat_norm
        est_speed = np.power(st.highest_speed - vs, 2) / np.power(st.highest_speed - st.lowest_speed, 2)
        est_time = 1.0 - (ts - st.min_t) / (st.max_t - st.min_t)
        est = (est_lat[:, None, None] + est_time[None, None, :]) + est_speed[None, :, None]
        if self.prev_best_idx is not None:
            p = self.prev_best_idx
            max_sqr = np.power(st.num_width, 2) + np.power(st.num_speed, 2) + np.power(st.num_t, 2)
            ii, jj, kk = np.meshgrid(np.arange(len(ds)), np.arange(len(vs)), np.arange(len(ts)), indexing="ij")
            sqr = (np.power(ii - p[0], 2) + np.power(jj - p[1], 2)) + np.power(kk - p[2], 2)
            heu = st.w_heuristic * sqr / max_sqr
        else:
            heu = np.zeros_like(est)
        self._cost_heu = heu
        self._cost_est = est + heu
        self._table = table
        self.sizes = np.array([len(ds), len(vs), len(ts)])
        self._generated = np.zeros(tuple(self.sizes), dtype=bool)
        self.trajs_3d = _GridView(self)
        return self.trajs_3d

    def _lin(self, idx) -> int:
        return (int(idx[0]) * int(self.sizes[1]) + int(idx[1])) * int(self.sizes[2]) + int(idx[2])

    def _end_state_of(self, lin: int) -> FrenetState:
        d, v, t = self._table[lin, 0], self._table[lin, 1], self._table[lin, 2]
        return FrenetState(t=t, s=0.0, s_d=v, s_dd=0.0, s_ddd=0.0, d=d, d_d=0.0, d_dd=0.0, d_ddd=0.0)

    def _placeholder(self, idx) -> FrenetTrajectory:
        lin = self._lin(idx)
        tr = FrenetTrajectory()
        tr.idx = np.array(idx)
        tr.end_state = self._end_state_of(lin)
        tr.cost_heu = self._cost_heu[tuple(idx)]
        tr.cost_est = self._cost_est[tuple(idx)]
        tr.is_generated = bool(self._generated[tuple(idx)])
        if tr.is_generated:
            tr.cost_final = self._cost[lin]
        return tr

    def _evaluate_grid(self, time_step_now: int):
        """The one launch per cycle: cost + masks of every lattice point."""
        self._prm = self._params(time_step_now)
        out = self.engine.plan_grid(self._ego6, self._fgrid, self._prm, want_records=False, want_volume=True,
                                    out=self._plan_outputs(self._fgrid, want_records=False))
        self._cost, self._flags = out["cost"][0], out["flags"][0]

    # -- lazy generation (bookkeeping only; the numbers are already on the host) ---------------------
    def generate_trajectory(self, idx: np.ndarray) -> tuple:
        """fiss_planner.py:101-138: first touch of a lattice point counts as a generation and
        enqueues ``(cost_final, idx)``; later touches just return the cost."""
        key = (int(idx[0]), int(idx[1]), int(idx[2]))
        cost = self._cost[self._lin(idx)]
        if self._generated[key]:
            return False, cost
        self.stats.num_trajs_generated += 1
        self._generated[key] = True
        self._generated_order.append(self._lin(idx))
        heapq.heappush(self.candidate_trajs, (cost, idx))
        return True, cost

    def find_initial_guess(self) -> np.ndarray:
        """Lowest ``cost_est`` among the not-yet-generated points, LAST one on ties (`<=`, :140-150)."""
        if self._generated.all():
            return None
        est = np.where(self._generated, np.inf, self._cost_est).ravel()
        lo = est.min()
        lin = len(est) - 1 - int(np.argmax(est[::-1] == lo))
        return np.array(np.unravel_index(lin, tuple(self.sizes)))

    def find_gradients(self, idx: np.ndarray) -> np.ndarray:
        """One-sided cost differences towards +1 in each dimension (towards -1 at the upper edge),
        zeroed when they point outside the grid (fiss_planner.py:152-172)."""
        _, centre = self.generate_trajectory(idx)
        grad = np.empty(3)
        for dim in range(3):
            nb = np.array(idx)
            if idx[dim] < self.sizes[dim] - 1:
                nb[dim] += 1
                _, cost = self.generate_trajectory(nb)
                grad[dim] = cost - centre
                if grad[dim] >= 0 and idx[dim] == 0:
                    grad[dim] = 0.0
            else:
                nb[dim] -= 1
                _, cost = self.generate_trajectory(nb)
                grad[dim] = centre - cost
                if grad[dim] <= 0 and idx[dim] == self.sizes[dim] - 1:
                    grad[dim] = 0.0
        return grad

    def explore_next_sample(self, curr_idx: np.ndarray) -> tuple:
        """fiss_planner.py:174-188: converged when the point was already generated, else step one
        index per dimension against the gradient sign."""
        if self._generated[int(curr_idx[0]), int(curr_idx[1]), int(curr_idx[2])]:
            return True, curr_idx
        grad = self.find_gradients(curr_idx)
        step = np.where(grad > 0.0, -1, 1)
        return False, np.clip(np.array(curr_idx) + step, 0, self.sizes - 1)

    # -- validation ----------------------------------------------------------------------------------
    def _validate_flags(self, flags_word: int) -> bool:
        """check_constraints then check_collisions on one candidate, with the reference's counters
        (fiss_planner.py:237-260): a collision check is only counted when the constraints pass."""
        self.stats.num_trajs_validated += 1
        if flags_word & (_shim.FLAG_SPEED | _shim.FLAG_ACCEL | _shim.FLAG_CURVATURE):
            return False
        self.stats.num_collison_checks += 1
        return not (flags_word & _shim.FLAG_COLLISION)

    def _fetch_trajectory(self, end_row: np.ndarray, cost) -> FrenetTrajectory:
        one = self.engine.eval_end_states(self._ego6, end_row[None], self._prm, want_records=True)
        n_cart = int((one["flags"][0] >> _shim.FLAG_NCART_SHIFT) & _shim.FLAG_NCART_MASK)
        return FrenetTrajectory().fill_from_device_record(one["records"][0], int(end_row[3]), n_cart, cost)

    def _begin_cycle(self, frenet_state, max_target_speed, obstacles, time_step_now):
        self.stats = Stats()
        self.settings.highest_speed = max_target_speed
        self.start_state = frenet_state
        self.candidate_trajs = []
        self.best_traj = None
        self._generated_order = []
        self._upload_obstacles(obstacles)
        self._ego6 = frenet_state.as_ego6()
        self.sample_end_frenet_states()
        self._evaluate_grid(time_step_now)

    def _finish_cycle(self):
        order = np.array(self._generated_order, dtype=int)
        idx3 = np.array(np.unravel_index(order, tuple(self.sizes))).T if len(order) else None
        self.trajs_per_timestep = self._bundle(self._ego6, self._table[order], self._prm, self._cost[order],
                                               self._flags[order], idx3=idx3)
        self.all_trajs.append(self.trajs_per_timestep)
        self.trajs_per_timestep = []

    def _accept_lattice_winner(self, idx):
        lin = self._lin(idx)
        traj = self._fetch_trajectory(self._table[lin], self._cost[lin])
        traj.idx = idx
        traj.end_state = self._end_state_of(lin)
        traj.cost_heu = self._cost_heu[tuple(idx)]
        traj.cost_est = self._cost_est[tuple(idx)]
        self.best_traj = traj
        self.prev_best_idx = traj.idx

    def plan(self, frenet_state: FrenetState, max_target_speed: float, obstacles: list, time_step_now: int = 0) -> FrenetTrajectory:
        self._begin_cycle(frenet_state, max_target_speed, obstacles, time_step_now)
        while True:
            self.stats.num_iter += 1
            if not self.candidate_trajs:
                best_idx = self.find_initial_guess()
                if best_idx is None:
                    break                      # every lattice point searched, nothing feasible
            else:
                best_idx = self.candidate_trajs[0][1]
            converged = False
            while not converged:
                converged, best_idx = self.explore_next_sample(best_idx)
            if not self.candidate_trajs:
                break
            _, idx = heapq.heappop(self.candidate_trajs)
            if self._validate_flags(int(self._flags[self._lin(idx)])):
                self._accept_lattice_winner(idx)
                break
        self._finish_cycle()
        return self.best_traj
