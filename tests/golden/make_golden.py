"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE (build container only).

    python tests/golden/make_golden.py            # needs /root/reference

Runs the unmodified reference modules (through ref_harness.py's commonroad / shapely stubs) on
seeded synthetic scenes and freezes inputs + outputs.  The committed .npz files are what travels:
``-m "not gpu"`` tests pin oracle/fop_oracle.py against them, ``-m gpu`` tests pin the CUDA path.
Collision masks in the goldens are reference control flow + oracle/sat_geometry.py's predicate
(GEOS is absent): see ref_harness.py.

Two families:
* ``dense_<name>.npz``   -- every candidate of one FrenetOptimalPlanner lattice: cost, n, n',
                            constraint mask, collision mask, winner, and (x, y, yaw, s_d, c)
                            for every 5th candidate + the winner.
* ``loop_<method>_<name>.npz`` -- closed-loop cycles of FOP / FOP+ / FISS / FISS+ plan():
                            per-cycle winner index, end state, cost, Stats, full winner arrays.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore", category=SyntaxWarning)

import ref_harness as rh  # noqa: E402
from fiss_plus_planner_b200 import synthetic as syn  # noqa: E402

ref = rh.load_reference()

TRAJ_FIELDS = ("t", "s", "s_d", "s_dd", "s_ddd", "d", "d_d", "d_dd", "d_ddd", "x", "y", "yaw", "ds", "c", "c_d", "c_dd")


def pad(rows, width):
    out = np.full((len(rows), width), np.nan)
    for i, r in enumerate(rows):
        r = np.asarray(r, dtype=np.float64)
        out[i, :len(r)] = r
    return out


def scene_inputs(sc, ego6, veh, now):
    return dict(centerline=sc.centerline, ego=np.asarray(ego6, dtype=np.float64),
                obs_xyth=sc.obs.xyth, obs_lw=sc.obs.lw, obs_valid=sc.obs.valid,
                final_time_step=sc.obs.final_time_step, num_samples=np.array(sc.num_samples),
                min_t=sc.min_t, max_t=sc.max_t, max_target_speed=sc.max_target_speed,
                time_step_now=now, ego_l=veh.l, ego_w=veh.w, max_speed=veh.max_speed, max_accel=veh.max_accel)


def frenet_state(e):
    return ref.FrenetState(0.0, e[0], e[1], e[2], 0.0, e[3], e[4], e[5], 0.0)


def dense_case(name, sc, ego6, veh_kw=None, now=0):
    veh = ref.Vehicle(rh.make_vehicle_params(**(veh_kw or {})))
    st = ref.FrenetOptimalPlannerSettings(*sc.num_samples)
    st.min_t, st.max_t = sc.min_t, sc.max_t
    pl = ref.FrenetOptimalPlanner(st, veh)
    pl.generate_frenet_frame(sc.centerline)
    obstacles = rh.make_ref_obstacles(sc.obs.xyth, sc.obs.lw, sc.obs.valid, sc.obs.final_time_step)
    st.highest_speed = sc.max_target_speed
    # stage by stage, so that masks exist for EVERY candidate (plan() only collision-checks survivors)
    fplist = pl.calc_global_paths(pl.calc_frenet_paths(frenet_state(ego6)))
    cost = np.array([fp.cost_final for fp in fplist])
    n = np.array([len(fp.t) for fp in fplist])
    n_cart = np.array([len(fp.x) for fp in fplist])
    ok = np.array([len(pl.check_constraints([fp])) == 1 for fp in fplist])
    coll = np.array([pl.has_collision(fp, obstacles, now, 2)[0] for fp in fplist])
    # and the real plan() for the winner
    pl2 = ref.FrenetOptimalPlanner(st, veh)
    pl2.generate_frenet_frame(sc.centerline)
    best = pl2.plan(frenet_state(ego6), sc.max_target_speed, obstacles, now)
    cands = pl2.all_trajs[-1]
    best_seq = -1 if best is None else [i for i, fp in enumerate(cands) if fp is best][0]
    keep = sorted(set(range(0, len(fplist), 5)) | ({best_seq} if best_seq >= 0 else set()))
    width = int(n.max())
    out = scene_inputs(sc, ego6, veh, now)
    out.update(cost=cost, n=n, n_cart=n_cart, constraint_ok=ok, collision=coll, best=best_seq,
               keep=np.array(keep), knots=np.array(pl.cubic_spline.s, dtype=np.float64))
    for f in ("x", "y", "yaw", "s_d", "c", "s", "d"):
        out["traj_" + f] = pad([getattr(fplist[i], f) for i in keep], width)
    np.savez_compressed(os.path.join(HERE, f"dense_{name}.npz"), **out)
    print(f"dense_{name}: C={len(fplist)} n={n.min()}..{n.max()} n'={n_cart.min()}..{n_cart.max()} "
          f"constraint_ok={ok.sum()} collision={coll.sum()} best={best_seq} cost_best={cost[best_seq] if best_seq >= 0 else None}")


def loop_case(method, name, sc, ego6, cycles, veh_kw=None):
    veh = ref.Vehicle(rh.make_vehicle_params(**(veh_kw or {})))
    cls, scls = {"FOP": (ref.FrenetOptimalPlanner, ref.FrenetOptimalPlannerSettings),
                 "FOP+": (ref.FopPlusPlanner, ref.FrenetOptimalPlannerSettings),
                 "FISS": (ref.FissPlanner, ref.FissPlannerSettings),
                 "FISS+": (ref.FissPlusPlanner, ref.FissPlusPlannerSettings)}[method]
    st = scls(*sc.num_samples)
    st.min_t, st.max_t = sc.min_t, sc.max_t
    if method == "FISS+":
        # refine_solution() stops on wall-clock time (fiss_plus_planner.py:153-156,296-299: the budget is
        # time_limit minus the time the coarse search took, checked even when has_time_limit is False).
        # A huge budget makes the golden deterministic: always max_refine_iters refinement rounds.
        st.time_limit = 1e9
    pl = cls(st, veh)
    pl.generate_frenet_frame(sc.centerline)
    obstacles = rh.make_ref_obstacles(sc.obs.xyth, sc.obs.lw, sc.obs.valid, sc.obs.final_time_step)
    fs = frenet_state(ego6)
    rec = dict(idx=[], end=[], cost=[], stats=[], n=[], n_cart=[], ego=[])
    arrays = {f: [] for f in TRAJ_FIELDS}
    for i in range(cycles):
        rec["ego"].append([fs.s, fs.s_d, fs.s_dd, fs.d, fs.d_d, fs.d_dd])
        best = pl.plan(fs, sc.max_target_speed, obstacles, i)
        assert best is not None, "golden loop expects a solution every cycle"
        rec["idx"].append(np.array(best.idx))
        es = best.end_state
        rec["end"].append([np.nan] * 3 if es is None else [es.d, es.s_d, es.t])
        rec["cost"].append(best.cost_final)
        rec["stats"].append([pl.stats.num_iter, pl.stats.num_trajs_generated, pl.stats.num_trajs_validated,
                             pl.stats.num_collison_checks])
        rec["n"].append(len(best.t))
        rec["n_cart"].append(len(best.x))
        for f in TRAJ_FIELDS:
            arrays[f].append(np.asarray(getattr(best, f), dtype=np.float64))
        fs = best.frenet_state_at_time_step(1)
    out = scene_inputs(sc, ego6, veh, 0)
    out.update({k: np.array(v) for k, v in rec.items()})
    width = max(rec["n"])
    for f in TRAJ_FIELDS:
        out["best_" + f] = pad(arrays[f], width)
    tag = method.replace("+", "plus")
    np.savez_compressed(os.path.join(HERE, f"loop_{tag}_{name}.npz"), **out)
    print(f"loop_{tag}_{name}: idx={[list(map(int, i)) for i in rec['idx']]} cost={np.round(rec['cost'], 4)} stats={rec['stats']}")


def blocked_scene():
    """cfg1-like scene with one slow obstacle in the ego lane ahead, so the cheapest candidates collide."""
    sc = syn.make_scene("cfg1_demo_substitute")
    rng = np.random.default_rng(77)
    # obstacle 0: s = 60 + 2.0 t, d = 1.3 (left half of the corridor: only right-swerving candidates pass)
    for t in range(sc.obs.xyth.shape[1]):
        s = 60.0 + 2.0 * 0.1 * t
        px, py = sc.spline.calc_position(s)
        yaw = sc.spline.calc_yaw(s)
        sc.obs.xyth[0, t] = (px - 1.3 * np.sin(yaw), py + 1.3 * np.cos(yaw), yaw)
    sc.obs.lw[0] = (4.5, 1.9)
    sc.obs.valid[0, :] = True
    sc.ego = np.array([[12.0, 9.0, 0.2, 0.15, 0.05, 0.0]])
    del rng
    return sc


def short_line_scene(length_knots, lattice=(5, 4, 3)):
    sc = syn.make_scene("cfg2_single_ego_8obs")
    sc.centerline = syn.reference_line(length_knots)
    from fiss_plus_planner_b200.planners.common.geometry.cubic_spline import CubicSpline2D
    sc.spline = CubicSpline2D(sc.centerline[:, 0], sc.centerline[:, 1])
    sc.num_samples = lattice
    return sc


def main():
    # ---- dense lattices (ego indices picked so that the masks are mixed)
    sc = syn.make_scene("cfg2_single_ego_8obs", batch=12)
    dense_case("cfg2_m8", sc, sc.ego[5])
    sc = syn.make_scene("cfg2_single_ego_8obs", batch=12)
    ego = sc.ego[5].copy()
    ego[3:] = 0.0           # d0 = d_d0 = d_dd0 = 0: the +-d halves of the lattice tie exactly
    dense_case("cfg2_m8_symmetric", sc, ego)
    sc = syn.make_scene("cfg3_64obs", batch=12)
    dense_case("cfg3_m64", sc, sc.ego[9])
    sc = syn.make_scene("cfg1_demo_substitute", batch=12)
    dense_case("cfg1_m27_tight_now30", sc, sc.ego[4], veh_kw=dict(v_max=12.9, a_max=1.0), now=30)
    sc = blocked_scene()
    dense_case("cfg1_blocked", sc, sc.ego[0])
    # truncation: a 13-knot (60 m) line; ego 30 m before its end, 0.4 m before, and beyond it
    sc = short_line_scene(13, (5, 6, 3))
    dense_case("short_line_partial", sc, np.array([30.0, 9.0, 0.3, 0.2, 0.0, 0.0]))
    sc = short_line_scene(13)
    dense_case("short_line_n1", sc, np.array([59.6, 9.0, 0.0, 0.1, 0.0, 0.0]))
    dense_case("short_line_n0", sc, np.array([61.0, 9.0, 0.0, 0.1, 0.0, 0.0]))
    # ---- closed loops, all four planners
    for method in ("FOP", "FOP+", "FISS", "FISS+"):
        sc = syn.make_scene("cfg1_demo_substitute", batch=12)
        loop_case(method, "cfg1", sc, sc.ego[4], cycles=4)
        sc = blocked_scene()
        loop_case(method, "blocked", sc, sc.ego[0], cycles=3)


if __name__ == "__main__":
    main()
