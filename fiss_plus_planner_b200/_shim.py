"""ctypes binding of libfissgpu.so (include/fiss_abi.h).  No fallback: if the CUDA library is not
built, importing a planner that needs it raises with the build command."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# FISSGPU_LIB: load another build of the same library (A/B runs of kernel variants on the GPU box)
LIB_PATH = os.environ.get("FISSGPU_LIB") or os.path.join(_HERE, "libfissgpu.so")

FISS_OK = 0
FLAG_SPEED, FLAG_ACCEL, FLAG_CURVATURE, FLAG_COLLISION = 1, 2, 4, 8
FLAG_INFEASIBLE_MASK = 15
FLAG_NCART_SHIFT, FLAG_NCART_MASK = 8, 0xFFFF
REC_ROWS = 16
MAT_ROWS = 5

EXPORTS = (
    "fiss_create", "fiss_destroy", "fiss_last_error", "fiss_abi_version", "fiss_arange_len",
    "fiss_set_spline", "fiss_fit_splines_host", "fiss_frame_samples_host", "fiss_set_obstacles", "fiss_set_obstacles_waymo",
    "fiss_eval_candidates_dev", "fiss_eval_grid_dev", "fiss_pick_winners_dev", "fiss_full_records_dev",
    "fiss_plan_lattice_host", "fiss_plan_grid_host", "fiss_eval_end_states_host", "fiss_launch_count",
    "fiss_plan_grid_dev", "fiss_plan_grid_submit", "fiss_plan_grid_wait",
    "fiss_comm_unique_id", "fiss_comm_init", "fiss_comm_destroy", "fiss_allreduce_pick",
)


class FissParams(C.Structure):
    """struct fiss_params (fiss_abi.h)."""
    _fields_ = [
        ("tick_t", C.c_double), ("target_speed", C.c_double), ("max_speed", C.c_double),
        ("max_accel", C.c_double), ("max_curvature", C.c_double), ("ego_length", C.c_double),
        ("ego_width", C.c_double), ("cost_time_offset", C.c_double), ("w_speed", C.c_double),
        ("w_accel", C.c_double), ("w_jerk", C.c_double), ("w_offset", C.c_double),
        ("time_step_now", C.c_int32), ("check_res", C.c_int32), ("check_curvature", C.c_int32),
        ("collide_all", C.c_int32),
    ]


class FissGrid(C.Structure):
    """struct fiss_grid (fiss_abi.h): the product lattice d_end x v_end x T with its candidate numbering."""
    _fields_ = [
        ("d_end", C.POINTER(C.c_double)), ("v_end", C.POINTER(C.c_double)), ("T", C.POINTER(C.c_double)),
        ("nd", C.c_int32), ("nv", C.c_int32), ("nt", C.c_int32),
        ("stride_d", C.c_int32), ("stride_v", C.c_int32), ("stride_t", C.c_int32),
    ]


GRID_AXIS_MAX = 64


class FissError(RuntimeError):
    pass


_lib = None


def load():
    """Load libfissgpu.so once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FissError(
            f"{LIB_PATH} is missing: the CUDA extension is not built.  Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or fiss_plus_planner_b200/build.py) first. "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, f64 = C.c_void_p, C.c_int32, C.c_double
    pp = C.POINTER(FissParams)
    gp = C.POINTER(FissGrid)
    sig = {
        "fiss_create": (i32, [i32, C.POINTER(vp)]),
        "fiss_destroy": (i32, [vp]),
        "fiss_last_error": (C.c_char_p, [vp]),
        "fiss_abi_version": (i32, []),
        "fiss_arange_len": (i32, [f64, f64]),
        "fiss_set_spline": (i32, [vp, vp, vp, i32]),
        "fiss_fit_splines_host": (i32, [vp, vp, vp, i32, i32, vp, i32]),
        "fiss_frame_samples_host": (i32, [vp, vp, f64, i32, vp]),
        "fiss_set_obstacles": (i32, [vp, vp, vp, vp, vp, i32, i32, i32]),
        "fiss_set_obstacles_waymo": (i32, [vp, vp, vp, vp, i32, i32, i32, C.POINTER(i32)]),
        "fiss_eval_candidates_dev": (i32, [vp, vp, vp, i32, vp, i32, pp, vp, vp, vp, i32]),
        "fiss_eval_grid_dev": (i32, [vp, vp, vp, i32, gp, pp, vp, vp, vp, i32]),
        "fiss_plan_grid_host": (i32, [vp, vp, vp, i32, gp, pp, vp, vp, vp, vp, i32, vp, vp]),
        "fiss_pick_winners_dev": (i32, [vp, vp, vp, i32, vp, i32, pp, vp, vp, vp, vp, vp, vp, i32]),
        "fiss_full_records_dev": (i32, [vp, vp, vp, vp, vp, i32, pp, vp, vp, vp, i32]),
        "fiss_plan_lattice_host": (i32, [vp, vp, vp, i32, vp, i32, pp, vp, vp, vp, vp, i32, vp, vp]),
        "fiss_eval_end_states_host": (i32, [vp, vp, vp, vp, i32, pp, vp, vp, vp, i32]),
        "fiss_launch_count": (C.c_int64, [vp]),
        "fiss_plan_grid_dev": (i32, [vp, vp, vp, i32, gp, pp, vp, vp, vp, vp, vp, vp, vp, i32]),
        "fiss_plan_grid_submit": (i32, [vp, vp, i32, vp, i32, gp, pp, vp, vp, vp, vp, i32]),
        "fiss_plan_grid_wait": (i32, [vp, i32]),
        "fiss_comm_unique_id": (i32, [vp, C.c_char_p]),
        "fiss_comm_init": (i32, [vp, vp, i32, i32, C.c_char_p]),
        "fiss_comm_destroy": (i32, [vp]),
        "fiss_allreduce_pick": (i32, [vp, vp, i32, i32, vp, i32, C.c_int64, C.c_int64, C.c_int64, vp, vp, vp, vp, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def arange_len(T: float, tick: float) -> int:
    return int(load().fiss_arange_len(float(T), float(tick)))


def ptr(a):
    """Raw address of a C-contiguous NumPy array (or None)."""
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"], "need a C-contiguous ndarray"
    return a.ctypes.data_as(C.c_void_p)
