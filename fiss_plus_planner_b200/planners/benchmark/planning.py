"""Closed-loop benchmark driver (reference: planners/benchmark/planning.py:35-208, :291-334).

``frenet_optimal_planning(scenario, planning_problem, vehicle_params, method, num_samples)`` keeps the
reference's signature, control flow and return tuple ``(goal_reached, Trajectory | None, avg_time, time_list,
stats, all_trajs)``: global route -> spline frame -> Cartesian->Frenet of the initial state -> one
``planner.plan()`` per 0.1 s simulation step (each on the B200 engine), the ego advancing to index 1 of the
winner every cycle, the three goal tests of :150-162.  ``planning(cfg, output_dir, input_dir, file)`` is the
per-scenario entry ``scripts/demo_cr.py`` calls (same cfg keys, including the reference's quirk that ``num_t`` is
taken from ``N_W_SAMPLE``, :294); the GIF rendering of :336-448 and ``informed_planning`` (:211-250) are out of
scope -- ``planning`` returns the result tuple instead of drawing it.

Scenario objects are duck-typed: real ``commonroad-io`` objects work, and so do the ones from
``planners/commonroad_interface/commonroad_lite.py`` (the stand-in reader used here, where commonroad-io is absent).
"""
from __future__ import annotations

import os
import time

import numpy as np

from fiss_plus_planner_b200.planners.common.scenario.frenet import FrenetState, State
from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
from fiss_plus_planner_b200.planners.commonroad_interface.commonroad_lite import (CommonRoadFileReader, CustomState,
                                                                                   Trajectory)
from fiss_plus_planner_b200.planners.commonroad_interface.global_planner import GlobalPlanner
from fiss_plus_planner_b200.planners.commonroad_interface.vehicle_parameters import VehicleParameterMapping, VehicleType
from fiss_plus_planner_b200.planners.fiss_planner import FissPlanner, FissPlannerSettings
from fiss_plus_planner_b200.planners.fiss_plus_planner import FissPlusPlanner, FissPlusPlannerSettings
from fiss_plus_planner_b200.planners.fop_plus_planner import FopPlusPlanner
from fiss_plus_planner_b200.planners.frenet_optimal_planner import (FrenetOptimalPlanner, FrenetOptimalPlannerSettings,
                                                                    Stats)

_PLANNERS = {
    "FOP": (FrenetOptimalPlanner, FrenetOptimalPlannerSettings),
    "FOP+": (FopPlusPlanner, FrenetOptimalPlannerSettings),
    "FISS": (FissPlanner, FissPlannerSettings),
    "FISS+": (FissPlusPlanner, FissPlusPlannerSettings),
}


def make_planner(method: str, num_samples: tuple, vehicle: Vehicle, scenario=None, **planner_kw):
    """The planner/settings pair the driver builds for ``method`` (planning.py:83-99)."""
    if method not in _PLANNERS:
        print("ERROR: Planning method entered is not recognized!")
        raise ValueError
    cls, settings_cls = _PLANNERS[method]
    return cls(settings_cls(*num_samples), vehicle, scenario, **planner_kw)


def frenet_optimal_planning(scenario, planning_problem, vehicle_params, method: str, num_samples: tuple,
                            verbose: bool = True, planner_hook=None, **planner_kw):
    """One closed-loop run.  ``planner_hook(planner)``, if given, is called once after construction (tests use
    it to pin ``time_limit`` of FISS+); ``planner_kw`` goes to the planner constructor (``device=``, ``engine=``)."""
    say = print if verbose else (lambda *a, **k: None)
    # global route -> centre line [n, 4]                                                          planning.py:37-39
    ego_lane_pts = GlobalPlanner().plan_global_route(scenario, planning_problem).concat_centerline

    # target speed: end of the goal's velocity interval, else 13.5 m/s                            :42-53
    goal_region = planning_problem.goal
    if goal_region.state_list[0].has_value("velocity"):
        speed_interval = goal_region.state_list[0].velocity
        min_speed, max_speed = speed_interval.start, speed_interval.end
        say(f"    Speed interval {min_speed}, {max_speed} m/s")
    else:
        min_speed, max_speed = 0.0, 13.5
        say(f"    Scenario has no speed interval, using {min_speed}, {max_speed} m/s")

    # centre of the goal lanelet (middle vertex)                                                  :55-60
    goal_lanelet = scenario.lanelet_network.find_lanelet_by_id(goal_region.lanelets_of_goal_position[0][0])
    centre_vertices = goal_lanelet.center_vertices
    goal_center = centre_vertices[int((centre_vertices.shape[0] - 1) / 2)]

    # obstacles: static first, then dynamic; horizon from the first dynamic obstacle             :63-70
    obstacles_all = scenario.static_obstacles + scenario.dynamic_obstacles
    final_time_step = scenario.dynamic_obstacles[0].prediction.final_time_step

    vehicle = Vehicle(vehicle_params)
    planner = make_planner(method, num_samples, vehicle, scenario, **planner_kw)
    if planner_hook is not None:
        planner_hook(planner)
    _, ref_ego_lane_pts = planner.generate_frenet_frame(ego_lane_pts)                           # :101

    init = planning_problem.initial_state                                                        # :104-108
    start_state = State(t=0.0, x=init.position[0], y=init.position[1], yaw=init.orientation, v=init.velocity,
                        a=init.acceleration)
    current_frenet_state = FrenetState()
    current_frenet_state.from_state(start_state, ref_ego_lane_pts)

    processing_time, num_cycles = 0.0, 0
    state_list, time_list = [], []
    stats = Stats()
    goal_reached = False
    for i in range(final_time_step):                                                             # :120-162
        num_cycles += 1
        t0 = time.time()
        best_traj_ego = planner.plan(current_frenet_state, max_speed, obstacles_all, i)
        t1 = time.time()
        processing_time += t1 - t0
        stats += planner.stats
        if best_traj_ego is None:
            break
        # the ego moves to index 1 of the winner                                                  :135-146
        current_state = best_traj_ego.state_at_time_step(1)
        current_frenet_state = best_traj_ego.frenet_state_at_time_step(1)
        state = CustomState(time_step=i, position=np.array([current_state.x, current_state.y]),
                            orientation=current_state.yaw, velocity=current_frenet_state.s_d,
                            velocity_y=current_frenet_state.d_d)
        state_list.append(state)
        time_list.append(t1 - t0)

        if goal_region.is_reached(state):
            say("Goal Reached")
            goal_reached = True
            break
        elif np.hypot(state.position[0] - goal_center[0], state.position[1] - goal_center[1]) <= vehicle.l / 2:
            say("    Goal Reached")
            goal_reached = True
            break
        elif np.hypot(state.position[0] - ref_ego_lane_pts[-1, 0], state.position[1] - ref_ego_lane_pts[-1, 1]) <= 3.0:
            say("    Reaching End of the Map, Stopping, Goal Not Reached")
            goal_reached = True   # sic (:161)
            break

    avg_processing_time = processing_time / num_cycles                                           # :193-194
    stats.average(num_cycles)
    ego_vehicle_traj = Trajectory(initial_time_step=0, state_list=state_list) if state_list else None
    return goal_reached, ego_vehicle_traj, avg_processing_time, time_list, stats, planner.all_trajs


def planning(cfg: dict, output_dir: str, input_dir: str, file: str, **planner_kw):
    """Per-scenario entry of scripts/demo_cr.py (planning.py:291-334), without the GIF stage.  Returns the
    ``frenet_optimal_planning`` tuple, or ``None`` when the scenario is infeasible ("<file> not feasible!")."""
    method = cfg["PLANNER"]
    num_samples = (cfg["N_W_SAMPLE"], cfg["N_S_SAMPLE"], cfg["N_W_SAMPLE"])   # sic: N_T_SAMPLE is never read (:294)
    vehicle_params = VehicleParameterMapping[VehicleType.VW_VANAGON.name].value

    scenario, planning_problem_set = CommonRoadFileReader(os.path.join(input_dir, file)).open()
    planning_problem = list(planning_problem_set.planning_problem_dict.values())[0]
    try:
        if method == "informed":
            raise NotImplementedError("informed_planning (SMP graph search) is out of scope: DESIGN.md section 7")
        result = frenet_optimal_planning(scenario, planning_problem, vehicle_params, method, num_samples, **planner_kw)
        if result[1] is None:
            print("No ego vehicle trajectory found")
            raise RuntimeError
    except RuntimeError:
        print("   ", f"{file} not feasible!")
        return None
    goal_reached, traj, avg_time, time_list, stats, _ = result
    print("   ", f"{file}: goal_reached={goal_reached} cycles={len(time_list)} avg plan() {1e3 * avg_time:.3f} ms "
          f"(p50 {1e3 * float(np.median(time_list)):.3f} ms) generated/cycle={stats.num_trajs_generated:.1f}")
    if output_dir and cfg.get("SAVE_TRAJECTORY", False):
        os.makedirs(output_dir, exist_ok=True)
        rows = np.array([[s.time_step, s.position[0], s.position[1], s.orientation, s.velocity, s.velocity_y]
                         for s in traj.state_list])
        np.savetxt(os.path.join(output_dir, os.path.splitext(file)[0] + f"_{method}.csv"), rows, delimiter=",",
                   header="time_step,x,y,orientation,velocity,velocity_y")
    return result
