"""Closed-loop goldens from the reference's OWN driver (build container only; needs /root/reference).

    python tests/golden/make_golden_driver.py

Runs the unmodified ``planners/benchmark/planning.py::frenet_optimal_planning`` of the reference (through
ref_harness.install_driver_stubs) on ``data/demo/DEU_Flensburg-1_1_T-1.xml`` (BASELINE config #1) and ``DEU_Flensburg-26_1_T-1.xml`` (obstacles cross the
ego lane: several candidates are rejected per cycle) for FOP / FOP+ /
FISS / FISS+ with the demo lattice (5, 5, 5) and freezes, per method, ``driver_<method>_<scenario>.npz``:
the centre line the reference's global_planner.py assembled, per-cycle plan() input (Frenet ego state), winner
(idx, end state, cost, n, n'), Stats, the recorded ego states, the averaged Stats, goal_reached and cycle count.
The scenario itself travels as ``scenario_<name>.xml.gz`` re-written by our own writer (full lanelet network,
all obstacles, planning problem; float repr round-trips exactly), so the GPU tests drive reader -> route ->
frame -> closed loop without /root/reference.
"""
from __future__ import annotations

import gzip
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import ref_harness as rh  # noqa: E402

planning = rh.load_reference_driver()
from fiss_plus_planner_b200.planners.commonroad_interface import commonroad_lite as crl  # noqa: E402
from fiss_plus_planner_b200.planners.commonroad_interface.vehicle_parameters import VehicleParameterMapping  # noqa: E402

SCENARIOS = ("DEU_Flensburg-1_1_T-1", "DEU_Flensburg-26_1_T-1")   # config #1; a scene where collisions decide
# the other three demo scenarios with the two cheap searches only (the reference's FOP costs ~1 s per cycle, 80-100 cycles)
SCENARIOS_SEARCH_ONLY = ("DEU_Lohmar-15_1_T-1", "DEU_Lohmar-54_1_T-1", "DEU_Lohmar-65_1_T-1")
NUM_SAMPLES = (5, 5, 5)


def record_plans(cls, log):
    """Wrap ``cls.plan`` (the unmodified reference method still does the work) to log inputs and outputs."""
    orig = cls.__dict__["plan"]

    def plan(self, frenet_state, max_target_speed, obstacles, time_step_now=0):
        fs = frenet_state
        best = orig(self, frenet_state, max_target_speed, obstacles, time_step_now)
        es = None if best is None else best.end_state
        log.append(dict(
            ego=[fs.s, fs.s_d, fs.s_dd, fs.d, fs.d_d, fs.d_dd], max_target_speed=max_target_speed, now=time_step_now,
            idx=[-9, -9, -9] if best is None else list(np.asarray(best.idx)),
            end=[np.nan] * 3 if es is None else [es.d, es.s_d, es.t],
            cost=np.nan if best is None else best.cost_final,
            n=-1 if best is None else len(best.t), n_cart=-1 if best is None else len(best.x),
            stats=[self.stats.num_iter, self.stats.num_trajs_generated, self.stats.num_trajs_validated,
                   self.stats.num_collison_checks]))
        return best

    cls.plan = plan
    return orig


def run(method, scenario, problem, vehicle_params, SCENARIO):
    cls = {"FOP": planning.FrenetOptimalPlanner, "FOP+": planning.FopPlusPlanner, "FISS": planning.FissPlanner,
           "FISS+": planning.FissPlusPlanner}[method]
    log = []
    orig = record_plans(cls, log)
    # FISS+ refinement stops on wall-clock time (fiss_plus_planner.py:153-156,296-299); a huge budget makes the
    # golden deterministic (always max_refine_iters rounds)
    settings_cls = planning.FissPlusPlannerSettings
    settings_init = settings_cls.__init__

    def patched_init(self, *a, **k):
        settings_init(self, *a, **k)
        self.time_limit = 1e9
    settings_cls.__init__ = patched_init
    centerline = []
    gen = cls.generate_frenet_frame if "generate_frenet_frame" in cls.__dict__ else None
    base = planning.FrenetOptimalPlanner
    base_gen = base.generate_frenet_frame

    def gen_spy(self, pts):
        centerline.append(np.array(pts))
        return base_gen(self, pts)
    base.generate_frenet_frame = gen_spy
    try:
        goal_reached, traj, avg_t, time_list, stats, _ = planning.frenet_optimal_planning(
            scenario, problem, vehicle_params, method, NUM_SAMPLES)
    finally:
        cls.plan = orig
        settings_cls.__init__ = settings_init
        base.generate_frenet_frame = base_gen
    del gen
    states = np.array([[s.time_step, s.position[0], s.position[1], s.orientation, s.velocity, s.velocity_y]
                       for s in traj.state_list])
    out = dict(centerline=centerline[0], goal_reached=goal_reached, cycles=len(log), states=states,
               avg_stats=np.array([stats.num_iter, stats.num_trajs_generated, stats.num_trajs_validated,
                                   stats.num_collison_checks], dtype=np.float64),
               num_samples=np.array(NUM_SAMPLES))
    for k in log[0]:
        out["cycle_" + k] = np.array([c[k] for c in log], dtype=np.float64)
    tag = method.replace("+", "plus")
    np.savez_compressed(os.path.join(HERE, f"driver_{tag}_{SCENARIO}.npz"), **out)
    print(f"driver_{tag}: cycles={len(log)} goal_reached={goal_reached} avg_stats={out['avg_stats']} "
          f"last state={states[-1]}")


def main():
    only = sys.argv[1:]
    for name in SCENARIOS:
        if not only or name in only:
            one_scenario(name)
    for name in SCENARIOS_SEARCH_ONLY:
        if not only or name in only:
            one_scenario(name, ("FISS", "FISS+"))


def one_scenario(SCENARIO, methods=("FOP", "FOP+", "FISS", "FISS+")):
    src = os.path.join(rh.REFERENCE_ROOT, "data", "demo", SCENARIO + ".xml")
    scenario, pps = crl.CommonRoadFileReader(src).open()
    problem = list(pps.planning_problem_dict.values())[0]
    tmp = os.path.join(HERE, f"scenario_{SCENARIO}.xml")
    crl.write_commonroad_xml(tmp, scenario, problem, SCENARIO)
    # the written file must read back to the same numbers
    sc2, pps2 = crl.CommonRoadFileReader(tmp).open()
    for a, b in zip(scenario.dynamic_obstacles, sc2.dynamic_obstacles):
        ta, va = a.dense_table(102)
        tb, vb = b.dense_table(102)
        assert np.array_equal(ta, tb) and np.array_equal(va, vb)
    for a, b in zip(scenario.lanelet_network.lanelets, sc2.lanelet_network.lanelets):
        assert np.array_equal(a.left_vertices, b.left_vertices) and a.successor == b.successor
    with open(tmp, "rb") as f, gzip.GzipFile(tmp + ".gz", "wb", mtime=0) as g:
        g.write(f.read())
    os.remove(tmp)
    rh.attach_shapely_shapes(scenario)
    vehicle_params = VehicleParameterMapping["VW_VANAGON"].value
    for method in methods:
        run(method, scenario, problem, vehicle_params, SCENARIO)


if __name__ == "__main__":
    main()
