"""cProfile of the FISS / FISS+ host search inside the closed loop (GPU box): where the per-cycle host time goes."""
import cProfile
import gzip
import os
import pstats
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fiss_plus_planner_b200.planners.benchmark.planning import frenet_optimal_planning  # noqa: E402
from fiss_plus_planner_b200.planners.commonroad_interface.commonroad_lite import CommonRoadFileReader  # noqa: E402
from fiss_plus_planner_b200.planners.commonroad_interface.vehicle_parameters import VehicleParameterMapping  # noqa: E402

name = "DEU_Flensburg-1_1_T-1"
with tempfile.TemporaryDirectory() as tmp:
    dst = os.path.join(tmp, name + ".xml")
    with gzip.open(os.path.join(ROOT, "tests", "golden", f"scenario_{name}.xml.gz"), "rb") as g, open(dst, "wb") as f:
        f.write(g.read())
    sc, pps = CommonRoadFileReader(dst).open()
pp = list(pps.planning_problem_dict.values())[0]
vp = VehicleParameterMapping["VW_VANAGON"].value
for method in sys.argv[1:] or ["FISS", "FISS+"]:
    frenet_optimal_planning(sc, pp, vp, method, (5, 5, 5), verbose=False)
    pr = cProfile.Profile()
    pr.enable()
    res = frenet_optimal_planning(sc, pp, vp, method, (5, 5, 5), verbose=False)
    pr.disable()
    print("=====", method, "cycles", len(res[3]), "p50 ms", 1e3 * sorted(res[3])[len(res[3]) // 2])
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
