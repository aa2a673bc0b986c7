#!/usr/bin/env python
"""bench.py -- candidate trajectories / s of the Frenet lattice hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic ego states: every candidate of a
9x6x5 (d, v, T) lattice x <=50 time steps against 32 predicted obstacles (BASELINE.json north_star /
configs[3] per-GPU shard: 512 ego states per GPU, weak scaling -> 4096 on 8 GPUs).

* ``value``   device-resident throughput: ego states already in HBM; the step is the fused lattice
              kernel in FULL-MATERIALISATION mode (x, y, yaw, v, kappa of every candidate written to
              HBM, FP64: 276 MB per step per GPU, larger than the 126 MB L2) + the record kernel (the pick
              fused in: argmin per ego state, then the winners' full records).  CUDA events on the launching
              stream, max over ranks.
* ``e2e``     the same metric through the host C-ABI with HOST buffers, H2D of the ego states and D2H of winners /
              records inside the timed region, every step: the streaming pair ``fiss_plan_grid_submit`` /
              ``fiss_plan_grid_wait`` (two lanes: the copy-back of step k under the kernels of step k+1);
              ``value_sync_call`` = one blocking ``fiss_plan_grid_host`` per step.  This path runs the WINNER-ONLY lattice
              kernel + the winners' records (what plan() returns), not the materialising kernel ``value`` times.
* ``roofline`` the lattice kernel alone: algorithmic bytes (SURVEY 8(d) formula) / its CUDA-event time,
              against MEASURED_PEAKS.json's HBM copy bandwidth.  The kernel is FP64-issue bound, not
              HBM bound: the fraction is reported as measured, not dressed up.
* ``cpu_baseline`` the reference's CPU path on all host cores: a Pool over independent ego states, one whole plan()
              per task -- the unmodified reference where /root/reference is mounted, else the oracle port.

``--impl reference`` times that CPU path alone (the reference is Python + commonroad/shapely with nothing to
install and cannot travel to the GPU box, where its hot path runs as restated in oracle/, pinned by tests/golden).
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LATTICE = (9, 6, 5)
MIN_T, MAX_T = 4.0, 5.0
NUM_OBS = 32
BATCH_PER_GPU = 512
SCENE = "cfg4_batch4096_32obs"
# --workload: the other BASELINE configs as side measurements (the default, and the line the driver reads, is cfg4)
WORKLOADS = {
    "cfg4": dict(scene="cfg4_batch4096_32obs", lattice=(9, 6, 5), t=(4.0, 5.0), obstacles=32, batch=512,
                 text="cfg4 shard: %d ego states/GPU x 9x6x5 lattice (270 candidates) x n<=50 steps x %d obstacles "
                      "(north_star target config; configs[3] = 4096 states on 8 GPUs)"),
    "cfg3": dict(scene="cfg3_64obs", lattice=(9, 6, 5), t=(4.0, 5.0), obstacles=64, batch=512,
                 text="cfg3 (collision-check bound): %d ego states/GPU x 9x6x5 lattice x n<=50 steps x %d obstacles"),
    "cfg5": dict(scene="cfg5_fine_lattice", lattice=(33, 17, 9), t=(8.0, 10.0), obstacles=32, batch=32,
                 text="cfg5 (fine lattice): %d ego states/GPU x 33x17x9 lattice (5049 candidates) x n<=100 steps x %d obstacles"),
}


def select_workload(name: str, batch_per_gpu):
    global LATTICE, MIN_T, MAX_T, NUM_OBS, BATCH_PER_GPU, SCENE, WORKLOAD_TEXT
    w = WORKLOADS[name]
    LATTICE, (MIN_T, MAX_T), NUM_OBS, SCENE, WORKLOAD_TEXT = w["lattice"], w["t"], w["obstacles"], w["scene"], w["text"]
    BATCH_PER_GPU = w["batch"]
    return BATCH_PER_GPU if batch_per_gpu is None else batch_per_gpu


WORKLOAD_TEXT = WORKLOADS["cfg4"]["text"]
METRIC = "candidate_trajectories_per_sec"
UNIT = "candidates/s"


# ------------------------------------------------------------------------------------------------ scene
def make_workload(n_gpus: int, batch_per_gpu: int):
    from fiss_plus_planner_b200 import synthetic as syn
    sc = syn.make_scene(SCENE, batch=batch_per_gpu * n_gpus, num_obstacles=NUM_OBS)
    return sc


def config_dict(n_gpus, batch_per_gpu):
    return {
        "workload": WORKLOAD_TEXT % (batch_per_gpu, NUM_OBS),
        "lattice": list(LATTICE), "steps_per_candidate": "%d..%d (T in [%g,%g] s, tick 0.1 s)" % (
            int(np.ceil(MIN_T / 0.1)), int(np.ceil(MAX_T / 0.1)), MIN_T, MAX_T),
        "obstacles": NUM_OBS, "batch_per_gpu": batch_per_gpu, "global_batch": batch_per_gpu * n_gpus,
        "parallelism": "problems sharded across %d GPU(s), no data-path collective" % n_gpus,
        "l2": "per-step output (%d MB/GPU in materialisation mode) exceeds the 126 MB L2" % round(
            5 * batch_per_gpu * LATTICE[0] * LATTICE[1] * LATTICE[2] * int(np.ceil(MAX_T / 0.1)) * 8 / 1e6),
    }


# ------------------------------------------------------------------------------------------------ CPU baseline
REFERENCE_ROOT = "/root/reference"
CPU_BLOCK = 6   # steps of one cycle of the CPU arm's sample (cpu_rate)
_CPU = {}


def _cpu_worker_init(kind, centerline, obs_arrays, lattice, min_t, max_t, max_target_speed, ego_l, ego_w, v_max, a_max):
    """Per-process set-up (spline fit, obstacle objects), outside the timed region: ``kind`` = "reference" builds the
    UNMODIFIED reference's FrenetOptimalPlanner from /root/reference (tests/golden/ref_harness.py: commonroad symbols
    stubbed, NumPy SAT in place of shapely/GEOS); "port" builds oracle/fop_oracle.py's restatement of it."""
    import warnings as _w
    _w.filterwarnings("ignore")
    if kind == "reference":
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import ref_harness as rh
        ref = rh.load_reference()
        veh = ref.Vehicle(rh.make_vehicle_params(l=ego_l, w=ego_w, v_max=v_max, a_max=a_max))
        st = ref.FrenetOptimalPlannerSettings(*lattice)
        st.min_t, st.max_t = min_t, max_t
        pl = ref.FrenetOptimalPlanner(st, veh)
        pl.generate_frenet_frame(centerline)
        obstacles = rh.make_ref_obstacles(*obs_arrays)
        _CPU.update(kind=kind, pl=pl, obstacles=obstacles, speed=max_target_speed,
                    state=lambda e: ref.FrenetState(0.0, e[0], e[1], e[2], 0.0, e[3], e[4], e[5], 0.0))
    else:
        from oracle import fop_oracle as fo
        st = fo.Settings(*lattice)
        st.min_t, st.max_t, st.highest_speed = min_t, max_t, max_target_speed
        pl = fo.FopOracle(st, ego_l, ego_w, v_max, a_max)
        pl.generate_frenet_frame(centerline)
        _CPU.update(kind=kind, pl=pl, obstacles=fo.ObstacleTable(*obs_arrays), speed=max_target_speed)


def _cpu_worker_plan(ego6):
    """One whole planning problem: FrenetOptimalPlanner.plan() of one ego state (frenet_optimal_planner.py:247-270)."""
    pl = _CPU["pl"]
    if _CPU["kind"] == "reference":
        pl.plan(_CPU["state"](ego6), _CPU["speed"], _CPU["obstacles"], 0)
        n = len(pl.all_trajs[-1])
        pl.all_trajs.clear()
        return n
    pl.plan(tuple(ego6), _CPU["speed"], _CPU["obstacles"], 0)
    return len(pl.last_all)


def cpu_kind():
    """The real reference when its tree is mounted (this container), else the oracle port (the GPU box)."""
    forced = os.environ.get("FISS_BENCH_CPU_KIND")      # "port" / "reference": for comparing the two in one place
    if forced in ("port", "reference"):
        return forced
    return "reference" if os.path.isdir(os.path.join(REFERENCE_ROOT, "planners")) else "port"


def cpu_rate(sc, steps: int, warmup: int, budget_s: float, min_s: float = 5.0):
    """Candidates/s of the reference's CPU path on all host cores: a ``multiprocessing.Pool`` over
    INDEPENDENT ego states, one whole plan() per task -- the reference itself is single-threaded, so this is how a user
    would fill the host.  One step = one ego state per worker.  At least ``min_s`` seconds are sampled (more steps than
    asked if need be) and at most ``budget_s``."""
    from fiss_plus_planner_b200 import synthetic as syn
    kind = cpu_kind()
    workers = os.cpu_count() or 1
    n_cand = LATTICE[0] * LATTICE[1] * LATTICE[2]
    init = (kind, sc.centerline, (sc.obs.xyth, sc.obs.lw, sc.obs.valid, sc.obs.final_time_step), LATTICE, MIN_T, MAX_T,
            sc.max_target_speed, syn.EGO_L, syn.EGO_W, syn.EGO_V_MAX, syn.EGO_A_MAX)
    ctx = mp.get_context("fork")
    times, cands = [], 0
    with ctx.Pool(workers, initializer=_cpu_worker_init, initargs=init) as pool:
        t_begin = time.perf_counter()
        i = 0
        while True:
            # the steps cycle through ONE fixed block of CPU_BLOCK x workers ego states, and whole cycles are timed: the
            # time of a plan() varies ~2x with the traffic around the ego state, and two runs that sampled different ego
            # states (this leg of the two arms, different --steps / --warmup) disagreed by 20 %
            egos = [tuple(float(v) for v in sc.ego[((i % CPU_BLOCK) * workers + w) % len(sc.ego)]) for w in range(workers)]
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker_plan, egos, chunksize=1)
            dt = time.perf_counter() - t0
            assert all(r == n_cand for r in res)
            if i >= warmup:
                times.append(dt)
                cands += sum(res)
            i += 1
            elapsed = time.perf_counter() - t_begin
            if elapsed > budget_s and times and len(times) % CPU_BLOCK == 0:
                break
            if len(times) >= steps and sum(times) >= min_s and len(times) % CPU_BLOCK == 0:
                break
    total = float(np.sum(times))
    what = ("the UNMODIFIED reference FrenetOptimalPlanner.plan() from /root/reference (commonroad symbols stubbed, NumPy SAT "
            "in place of shapely/GEOS)" if kind == "reference" else
            "oracle/fop_oracle.py, the line-by-line Python port of the reference's plan() (pinned bit-exact to it by "
            "tests/golden; NumPy SAT in place of shapely/GEOS)")
    return dict(value=cands / total, steps_done=len(times), ms_per_step=1e3 * total / len(times), cores=workers, kind=kind,
                sample="%d step(s) in %.1f s; each = %d independent ego states, one whole plan() per process (%d-candidate "
                       "lattice, n<=%d, %d obstacles); %s" % (len(times), total, workers, n_cand, int(np.ceil(MAX_T / 0.1)),
                                                             NUM_OBS, what))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(np.max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ closed loop
CLOSED_LOOP_SCENARIO = "DEU_Flensburg-1_1_T-1"   # BASELINE configs[0] (scripts/demo_cr.py, lattice 5x5x5)


def _load_fixture_scenario(tmpdir):
    import gzip
    from fiss_plus_planner_b200.planners.commonroad_interface.commonroad_lite import CommonRoadFileReader
    src = os.path.join(ROOT, "tests", "golden", "scenario_%s.xml.gz" % CLOSED_LOOP_SCENARIO)
    if not os.path.exists(src):
        return None
    dst = os.path.join(tmpdir, CLOSED_LOOP_SCENARIO + ".xml")
    with gzip.open(src, "rb") as g, open(dst, "wb") as f:
        f.write(g.read())
    scenario, pps = CommonRoadFileReader(dst).open()
    return scenario, list(pps.planning_problem_dict.values())[0]


def closed_loop_ours(device: int):
    """plan()-cycle latency of the four drop-in planners inside frenet_optimal_planning() on the config-1 scenario
    (reader -> route -> frame -> one CUDA plan() per 0.1 s step, ego advanced from the winner)."""
    import tempfile
    from fiss_plus_planner_b200.planners.benchmark.planning import frenet_optimal_planning
    from fiss_plus_planner_b200.planners.commonroad_interface.vehicle_parameters import VehicleParameterMapping
    out = {"scenario": CLOSED_LOOP_SCENARIO, "lattice": [5, 5, 5], "note": "wall time of planner.plan() per cycle, "
           "as planning.py:124-128 measures it (host marshalling + H2D + kernels + D2H + host search)"}
    with tempfile.TemporaryDirectory() as tmp:
        loaded = _load_fixture_scenario(tmp)
        if loaded is None:
            return None
        scenario, problem = loaded
        vp = VehicleParameterMapping["VW_VANAGON"].value
        for method in ("FOP", "FOP+", "FISS", "FISS+"):
            frenet_optimal_planning(scenario, problem, vp, method, (5, 5, 5), verbose=False, device=device)  # warm-up run
            reached, traj, avg_t, times, stats, _ = frenet_optimal_planning(scenario, problem, vp, method, (5, 5, 5),
                                                                            verbose=False, device=device)
            out[method] = {"p50_ms": 1e3 * float(np.median(times)), "mean_ms": 1e3 * float(np.mean(times)),
                           "cycles": len(times), "goal_reached": bool(reached),
                           "generated_per_cycle": float(stats.num_trajs_generated)}
    return out


def closed_loop_cpu_port(cycles: int = 3):
    """The same scenario through the oracle port of FrenetOptimalPlanner.plan() (1 core, like the reference)."""
    import tempfile
    from fiss_plus_planner_b200.planners.commonroad_interface.global_planner import GlobalPlanner
    from fiss_plus_planner_b200.planners.commonroad_interface.vehicle_parameters import VehicleParameterMapping
    from fiss_plus_planner_b200.planners.common.scenario.frenet import FrenetState, State
    from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import marshal_obstacles
    from oracle import fop_oracle as fo
    with tempfile.TemporaryDirectory() as tmp:
        loaded = _load_fixture_scenario(tmp)
        if loaded is None:
            return None
        scenario, problem = loaded
    veh = Vehicle(VehicleParameterMapping["VW_VANAGON"].value)
    pts = GlobalPlanner().plan_global_route(scenario, problem).concat_centerline
    opl = fo.FopOracle(fo.Settings(5, 5, 5), veh.l, veh.w, veh.max_speed, veh.max_accel)
    opl.generate_frenet_frame(pts)
    tab = marshal_obstacles(scenario.static_obstacles + scenario.dynamic_obstacles)
    obs = fo.ObstacleTable(tab.xyth, tab.lw, tab.valid.astype(bool), tab.final_time_step)
    # start state as the driver computes it
    from fiss_plus_planner_b200.planners.common.geometry.cubic_spline import CubicSpline2D
    csp = CubicSpline2D(pts[:, 0], pts[:, 1])
    ss = np.arange(0, csp.s[-1], 0.1)
    ref = np.column_stack(([csp.calc_position(v) for v in ss], [csp.calc_yaw(v) for v in ss], [csp.calc_curvature(v) for v in ss]))
    init = problem.initial_state
    fs = FrenetState()
    fs.from_state(State(t=0.0, x=init.position[0], y=init.position[1], yaw=init.orientation, v=init.velocity, a=init.acceleration), ref)
    ego6 = (fs.s, fs.s_d, fs.s_dd, fs.d, fs.d_d, fs.d_dd)
    times = []
    for i in range(cycles):
        t0 = time.perf_counter()
        tr = opl.plan(ego6, 13.5, obs, i)
        times.append(time.perf_counter() - t0)
        ego6 = (tr.s[1], tr.s_d[1], tr.s_dd[1], tr.d[1], tr.d_d[1], tr.d_dd[1])
    return {"FOP": {"p50_ms": 1e3 * float(np.median(times)), "cycles": cycles, "cores": 1,
                    "kind": "port (oracle FopOracle.plan, NumPy SAT in place of shapely/GEOS)"}}


# ------------------------------------------------------------------------------------------------ ours
def algorithmic_bytes(end, batch, num_obs, num_knots):
    """SURVEY 8(d), full-materialisation FP64: per candidate 5*n*8 + 16; per problem the tables read once."""
    n = end[:, 3]
    per_problem_out = float(np.sum(5 * n * 8 + 16))
    t_chk = np.ceil(n.max() / 2)
    per_problem_in = 24 * num_obs * t_chk + 16 * num_obs + 72 * num_knots + 48
    return batch * (per_problem_out + per_problem_in)


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created; stdout must carry exactly ONE
        # JSON line, so everything else goes to stderr until the line is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world
    bpg = args.batch_per_gpu
    sc = make_workload(n_gpus, bpg)

    # each rank stays on its own slice of the host cores (8 ranks + NCCL / sampler threads migrating over one NUMA node
    # is what a 2.7 ms timed region does not forgive)
    try:
        avail = sorted(os.sched_getaffinity(0))
        per = len(avail) // max(world, 1)
        if world > 1 and per >= 2:
            os.sched_setaffinity(0, set(avail[local_rank * per:(local_rank + 1) * per]))
    except (AttributeError, OSError):
        pass

    # CPU baseline first (rank 0, N=1 only), before CUDA is initialised in this process (fork-safe)
    cpu = None
    if n_gpus == 1 and not args.no_cpu_baseline:
        cpu = cpu_rate(sc, steps=args.cpu_steps, warmup=1, budget_s=40.0)

    from fiss_plus_planner_b200.engine import FissEngine, fop_grid, make_params
    from fiss_plus_planner_b200.planners.common.cost.cost_function import CostFunction
    from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import FrenetOptimalPlannerSettings
    from fiss_plus_planner_b200 import synthetic as syn

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # the clock sampler runs from here on: nvidia-smi's start-up (NVML initialisation over every GPU of the box) must not
    # fall into a timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    veh = Vehicle(syn.vehicle_params())
    st = FrenetOptimalPlannerSettings(*LATTICE)
    st.min_t, st.max_t, st.highest_speed = MIN_T, MAX_T, sc.max_target_speed
    eng = FissEngine(local_rank)
    eng.set_spline(sc.spline.device_table())
    eng.set_obstacles(sc.obs.xyth, sc.obs.lw, sc.obs.valid, sc.obs.final_time_step)
    grid = fop_grid(st, veh.w)
    end = grid.table()
    prm = make_params(st, veh, CostFunction("WX1").as_device_weights())
    ego = np.ascontiguousarray(sc.ego[rank * bpg:(rank + 1) * bpg])
    B, C = ego.shape[0], end.shape[0]
    n_stride = int(end[:, 3].max())

    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    f64 = torch.float64
    ego_t = torch.tensor(ego, dtype=f64, device=dev)
    cost_t = torch.empty(B * C, dtype=f64, device=dev)
    flags_t = torch.empty(B * C, dtype=torch.int32, device=dev)
    mat_t = torch.empty((5, B * C, n_stride), dtype=f64, device=dev)
    bidx_t = torch.empty(B, dtype=torch.int32, device=dev)
    bcost_t = torch.empty(B, dtype=f64, device=dev)
    rec_t = torch.empty((B, 16, n_stride), dtype=f64, device=dev)
    meta_t = torch.empty((B, 2), dtype=torch.int32, device=dev)

    def step_device(mat=mat_t):
        # ONE call = one plan step: lattice kernel + record kernel (pick fused in); one cudaGraphLaunch from the 2nd call on
        eng.plan_grid_dev(ego_t, grid, prm, cost_t, flags_t, mat, bidx_t, bcost_t, meta_t, rec_t, n_stride, stream=sptr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=f64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_steps(fn):
        """K steps between two events (the headline): nothing but the steps' own launches sits on the stream between them,
        so that consecutive steps chain (the next step's CTAs start on the SMs this step's last items leave idle --
        programmatic dependent launch, fiss_abi.cu eval_grid).  Then the same K steps once more with an event after every
        step and a synchronisation behind it: the median of that pass is the time of a step ON ITS OWN (issued onto an idle
        GPU, the library keeps the single-step settings: fiss_abi.cu eval_grid, `in_train`)."""
        for _ in range(args.warmup):
            fn()
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        for i in range(args.steps):
            fn()
        t1.record(stream)
        barrier()
        total = t0.elapsed_time(t1)
        per = []
        for i in range(args.steps):   # one step, then a synchronisation: the GPU is idle when the next step is issued
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            torch.cuda.synchronize()
            per.append(a.elapsed_time(b))
        barrier()
        return max_over_ranks(total), max_over_ranks(float(np.median(per)))

    # ---- device-resident timing (value): full materialisation + pick + winners' records
    time.sleep(0.5 if rank == 0 else 0.0)     # let the sampler's start-up pass
    launches0 = eng.launch_count
    total_ms, median_ms = timed_steps(step_device)
    launches = (eng.launch_count - launches0) * args.steps // (2 * args.steps + args.warmup)   # (warm-up + the two timed passes)

    # ---- the lattice kernel alone (roofline), same buffers: (a) K launches between two events -- its average launch
    # duration as it runs in the step loop, consecutive launches chained; (b) CUDA events around every launch -- a launch on
    # its own (the events serialise the launches)
    for _ in range(args.warmup):
        eng.eval_grid_dev(ego_t, grid, prm, cost_t, flags_t, mat_t, n_stride, stream=sptr)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for i in range(args.steps):
        eng.eval_grid_dev(ego_t, grid, prm, cost_t, flags_t, mat_t, n_stride, stream=sptr)
    t1.record(stream)
    barrier()
    kern_ms = t0.elapsed_time(t1) / args.steps
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        k_ev[i][0].record(stream)
        eng.eval_grid_dev(ego_t, grid, prm, cost_t, flags_t, mat_t, n_stride, stream=sptr)
        k_ev[i][1].record(stream)
    barrier()
    k_times = [a.elapsed_time(b) for a, b in k_ev]
    kern_alone_ms, kern_median_ms = float(np.mean(k_times)), float(np.median(k_times))

    # ---- winner-only mode (what plan() strictly needs; reported beside the headline)
    wo_ms, wo_median_ms = timed_steps(lambda: step_device(None))

    # ---- end to end through the host C-ABI (H2D + kernels + D2H inside the timed region, every step), page-locked host
    # buffers (engine.pinned_empty / alloc_plan_outputs).  (a) streaming: fiss_plan_grid_submit / _wait on two lanes, the
    # copy-back of step k under the kernels of step k + 1 -- how a caller with a sequence of batches drives it;
    # (b) one synchronous fiss_plan_grid_host call per step.
    ego_pin = eng.pinned_empty(ego.shape, np.float64)
    ego_pin[...] = ego
    n_lanes = 2   # two lanes hide the copy-back (a third changes nothing: the step is bound by what one stream carries)
    outs = [eng.alloc_plan_outputs(B, grid, want_records=True, want_volume=False, pinned=True) for _ in range(n_lanes)]

    def e2e_stream(k_steps):
        # steps k, k + 1, ... in flight on lanes k % n_lanes; a step's results are waited for n_lanes - 1 submits later
        for k in range(k_steps + n_lanes - 1):
            if k < k_steps:
                eng.plan_grid_submit(k % n_lanes, ego_pin, grid, prm, outs[k % n_lanes], stream=sptr)
            if k >= n_lanes - 1:
                eng.plan_grid_wait((k - (n_lanes - 1)) % n_lanes)

    e2e_stream(args.warmup)
    barrier()
    t0 = time.perf_counter()
    e2e_stream(args.steps)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    assert int(outs[(args.steps - 1) % n_lanes]["best_idx"][0]) == int(bidx_t[0].item())   # the device-resident path's winners
    for _ in range(args.warmup):
        eng.plan_grid(ego_pin, grid, prm, want_records=True, want_volume=False, stream=sptr, out=outs[0])
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = eng.plan_grid(ego_pin, grid, prm, want_records=True, want_volume=False, stream=sptr, out=outs[0])
    torch.cuda.synchronize()
    e2e_sync_s = max_over_ranks(time.perf_counter() - t0)
    assert int(out["best_idx"][0]) == int(bidx_t[0].item())
    # and once more with ordinary pageable NumPy buffers (staged through the handle's pinned bounce buffer)
    for _ in range(args.warmup):
        eng.plan_grid(ego, grid, prm, want_records=True, want_volume=False, stream=sptr)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.plan_grid(ego, grid, prm, want_records=True, want_volume=False, stream=sptr)
    torch.cuda.synchronize()
    e2e_pageable_s = max_over_ranks(time.perf_counter() - t0)
    # the clock record spans all timed regions (device-resident, kernel alone, winner-only, end-to-end)
    clocks = sampler.stop() if rank == 0 else None
    h2d = B * 48
    d2h = B * (8 + 4 + 8) + B * 16 * n_stride * 8

    # ---- plan-cycle latency, config 2: ONE ego state, 8 obstacles, through the same host call
    p50 = None
    if rank == 0:
        sc2 = syn.make_scene("cfg2_single_ego_8obs", batch=1)
        eng2 = FissEngine(local_rank)
        eng2.set_spline(sc2.spline.device_table())
        eng2.set_obstacles(sc2.obs.xyth, sc2.obs.lw, sc2.obs.valid, sc2.obs.final_time_step)
        out2 = eng2.alloc_plan_outputs(1, grid, want_records=True, want_volume=True, pinned=False)   # reused, as plan() does
        for _ in range(20):
            eng2.plan_grid(sc2.ego[:1], grid, prm, want_records=True, want_volume=True, stream=sptr, out=out2)
        lat = []
        for _ in range(400):
            t1 = time.perf_counter()
            eng2.plan_grid(sc2.ego[:1], grid, prm, want_records=True, want_volume=True, stream=sptr, out=out2)
            lat.append(time.perf_counter() - t1)
        p50 = 1e3 * float(np.median(lat))
        eng2.close()
    closed = None
    if rank == 0 and n_gpus == 1 and not args.no_closed_loop:
        closed = closed_loop_ours(local_rank)
        if closed is not None and not args.no_cpu_baseline:
            closed["cpu"] = closed_loop_cpu_port()

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        alg = algorithmic_bytes(end, B, NUM_OBS, len(sc.spline.s))
        # DRAM bytes per launch of the same kernel / workload from the committed `ncu --set full` capture
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and bpg == BATCH_PER_GPU and args.workload == "cfg4":
            tj = json.load(open(tpath))
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        achieved = alg / (kern_ms * 1e-3) / 1e9
        cand_total = B * C * n_gpus
        line = {
            "metric": METRIC, "value": cand_total * args.steps / (total_ms * 1e-3), "unit": UNIT,
            "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "ms_per_step_median": median_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(n_gpus, bpg),
            "mode": "full materialisation (x,y,yaw,v,kappa of every candidate to HBM, FP64) + pick + winner records",
            "value_winner_only": cand_total * args.steps / (wo_ms * 1e-3),
            "e2e": {"value": cand_total * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "mode": "winner-only lattice kernel + the winners' full records (NOT the materialising kernel that "
                            "`value` and `roofline` time): what plan() returns",
                    "call": "fiss_plan_grid_submit / fiss_plan_grid_wait, two lanes: pinned host ego states in, winners + "
                            "full records out to pinned host every step, the copy-back of step k under the kernels of "
                            "step k+1",
                    "value_sync_call": cand_total * args.steps / e2e_sync_s,
                    "sync_call": "one blocking fiss_plan_grid_host per step (no overlap between steps)",
                    "value_pageable_buffers": cand_total * args.steps / e2e_pageable_s},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "fiss_grid_kernel<yaw> (materialising)", "kernel_ms": kern_ms,
                         "kernel_ms_how": "average launch duration over a train of `steps` launches between two CUDA events "
                                          "(consecutive launches chained by programmatic dependent launch, as in the step loop)",
                         "kernel_ms_alone": kern_alone_ms, "kernel_ms_alone_median": kern_median_ms,
                         "kernel_ms_alone_how": "CUDA events around every single launch: the events keep the launches from overlapping (they are still issued back to back, so the library uses its train settings -- work drawn dynamically -- which cost a launch on its own ~3 %: with the single-step settings it measures 0.116-0.118 ms, profiles/r2y_chain_ab.txt)",
                         "frac_alone": alg / (kern_alone_ms * 1e-3) / 1e9 / peak,
                         "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                         "note": "FP64-issue bound by arithmetic (SURVEY 8(d)); see profiles/ for ncu fp64 pipe utilisation"},
            "plan_cycle_p50_ms": p50,
            "plan_cycle_config": "config 2: 1 ego state, 270 candidates, 8 obstacles, fiss_plan_grid_host incl. H2D/D2H of the "
                                 "winner, its full record and the whole cost / flags volume (the two kernels as one CUDA-graph launch)",
            "clocks": clocks,
        }
        if closed is not None:
            line["closed_loop"] = closed
        if cpu is not None:
            line["cpu_baseline"] = {"value": cpu["value"], "unit": UNIT, "cores": cpu["cores"], "kind": cpu["kind"],
                                    "sample": cpu["sample"]}
        if saved_stdout is not None:
            sys.stdout.flush()
            os.write(saved_stdout, (json.dumps(line) + "\n").encode())
        else:
            print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ split lattice
def run_split_lattice(args):
    """BASELINE config 5 with few problems: ONE fine lattice (33x17x9 = 5049 candidates x <=100 steps,
    32 obstacles) per problem, its lateral rows split across the ranks; every step = each rank's slab through the
    lattice kernel + pick, then ``fiss_allreduce_pick`` (NCCL from the C-ABI: one all-reduce for the pick, one for the
    winner's record).  STRONG scaling: the problems are the same on every rank, the work per GPU shrinks with N."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    saved_stdout = None
    if world > 1:
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from fiss_plus_planner_b200 import synthetic as syn
    from fiss_plus_planner_b200.batch import SplitLatticePlanner
    from fiss_plus_planner_b200.engine import FissEngine, fop_grid, make_params
    from fiss_plus_planner_b200.planners.common.cost.cost_function import CostFunction
    from fiss_plus_planner_b200.planners.common.vehicle.vehicle import Vehicle
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import FrenetOptimalPlannerSettings
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B = args.batch_per_gpu
    sc = syn.make_scene(SCENE, batch=B, num_obstacles=NUM_OBS)
    veh = Vehicle(syn.vehicle_params())
    st = FrenetOptimalPlannerSettings(*LATTICE)
    st.min_t, st.max_t, st.highest_speed = MIN_T, MAX_T, sc.max_target_speed
    eng = FissEngine(local_rank)
    eng.set_spline(sc.spline.device_table())
    eng.set_obstacles(sc.obs.xyth, sc.obs.lw, sc.obs.valid, sc.obs.final_time_step)
    grid = fop_grid(st, veh.w)
    prm = make_params(st, veh, CostFunction("WX1").as_device_weights())
    sp = SplitLatticePlanner(eng, grid, prm)
    if world > 1:
        eng.comm_init()
    ego = np.ascontiguousarray(sc.ego[:B])
    first = sp.plan(ego)                               # also allocates the device buffers
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    f64 = torch.float64

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=f64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if world == 1:
        sp._device_buffers(B)
        sp._dev["ego"].copy_(torch.from_numpy(ego))
    step = lambda: sp.plan_step_dev(stream=sptr)       # noqa: E731
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = eng.launch_count
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    marks[0].record(stream)
    for i in range(args.steps):
        step()
        marks[i + 1].record(stream)
    barrier()
    launches = eng.launch_count - launches0
    total_ms = max_over_ranks(marks[0].elapsed_time(marks[-1]))
    median_ms = max_over_ranks(float(np.median([marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps)])))
    # the device-resident path's winners are the synchronous path's
    assert np.array_equal(sp._dev["idx"].cpu().numpy().astype(np.int64), np.asarray(first["best_idx"]).astype(np.int64))
    for _ in range(args.warmup):
        sp.plan(ego)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = sp.plan(ego)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    C = grid.num_candidates
    n_stride = grid.n_stride
    if rank == 0:
        line = {
            "metric": METRIC, "value": B * C * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "ms_per_step_median": median_ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg5 split lattice: %d problem(s) x 33x17x9 lattice (5049 candidates) x n<=100 steps x %d "
                                   "obstacles; the lattice's 33 lateral rows split across %d GPU(s)" % (B, NUM_OBS, world),
                       "lattice": list(LATTICE), "obstacles": NUM_OBS, "problems": B,
                       "parallelism": "lateral rows of ONE lattice split across ranks; per step one ncclAllReduce(MIN, u64, "
                                      "%d B) for the pick + one ncclAllReduce(SUM, u64, %d B) moving the winners' records, "
                                      "both issued by libfissgpu.so on the kernels' stream" % (
                                          B * world * 16, B * (16 * n_stride + 1) * 8),
                       "l2": "winner-only mode: nothing is materialised; inputs stay resident"},
            "mode": "winner-only lattice kernel per slab + pick + winner records + cross-GPU pick",
            "e2e": {"value": B * C * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": B * 48,
                    "d2h_bytes_per_step": B * (8 + 4 + 8) + B * 16 * n_stride * 8, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "call": "SplitLatticePlanner.plan(): H2D of the ego states, slab kernels, fiss_allreduce_pick, D2H of the "
                            "global winners + records, synchronous per step"},
            "gpu_launches": int(launches), "winner_ids": [int(v) for v in np.asarray(out["best_idx"])[:8]],
        }
        if saved_stdout is not None:
            sys.stdout.flush()
            os.write(saved_stdout, (json.dumps(line) + "\n").encode())
        else:
            print(json.dumps(line))
    if world > 1:
        eng.comm_destroy()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path on the box's host cores, same config /
    metric / unit.  Where /root/reference is mounted (the build container) that is the UNMODIFIED
    ``FrenetOptimalPlanner.plan()`` (kind "reference"); on the GPU box -- the reference is pure Python with no installable
    package and cannot travel -- it is oracle/fop_oracle.py, the port pinned bit-exact to it (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_gpus = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    sc = make_workload(1, min(args.batch_per_gpu, 512))
    # keep the whole run within a few minutes whatever K and W are
    cpu = cpu_rate(sc, steps=args.steps, warmup=args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": n_gpus,
        "steps": cpu["steps_done"], "warmup": args.warmup, "ms_per_step": cpu["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(n_gpus, args.batch_per_gpu),
        "cpu_baseline": {"value": cpu["value"], "unit": UNIT, "cores": cpu["cores"], "kind": cpu["kind"],
                         "sample": cpu["sample"]},
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is pure Python + commonroad/shapely with no package to install (no setup.py / pyproject): "
                "baseline/_ref cannot exist.  kind=reference: its own plan() imported from /root/reference; kind=port: "
                "oracle/fop_oracle.py (pinned bit-exact to the reference by tests/golden) where that tree is absent",
    }
    print(json.dumps(line))


def main():
    warnings.filterwarnings("ignore")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=None)
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--split-lattice", action="store_true",
                    help="side workload (use with --workload cfg5): ONE lattice split across the ranks, cross-GPU pick by "
                         "fiss_allreduce_pick; --batch-per-gpu = number of problems (the same on every rank; default 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-closed-loop", action="store_true", help="skip the config-1 closed-loop latency section")
    ap.add_argument("--cpu-steps", type=int, default=6,
                    help="CPU-baseline sample: steps of one whole plan() per host core (6 x ~2 s = ~12 s of host time)")
    args = ap.parse_args()
    if args.split_lattice and args.batch_per_gpu is None:
        args.batch_per_gpu = 8
    args.batch_per_gpu = select_workload(args.workload, args.batch_per_gpu)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.split_lattice:
        run_split_lattice(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
