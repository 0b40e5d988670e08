"""ctypes binding of ``libetp_b200.so`` (C ABI declared in ``include/etp_b200.h``).

There is no CPU fallback: if the CUDA library has not been built, importing the product
functions raises.  ``build_library()`` compiles it in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
# ETP_B200_LIB selects another build of the same sources (kernel tuning experiments); default: the in-tree library
LIB_PATH = os.environ.get("ETP_B200_LIB") or os.path.join(HERE, "libetp_b200.so")
SOURCES = [os.path.join(HERE, "csrc", f) for f in ("etp_kernels.cu", "etp_score.cu", "etp_select.cu", "etp_abi.cu")]
HEADERS = [os.path.join(HERE, "csrc", f) for f in ("etp_device.cuh", "etp_bvn.cuh", "etp_harm.cuh", "etp_select.cuh", "etp_launch.h")] + [
    os.path.join(os.path.dirname(HERE), "include", "etp_b200.h")]

NUM_WEIGHTS = 14
NUM_FIELDS = 13
NUM_RED = 4
NUM_COST = 14
MAX_ANGLE_BINS = 8
LEVEL_SKIPPED = 255
PRECISION_FP32, PRECISION_FP64 = 0, 1
OK, ERR_CUDA, ERR_INVALID, ERR_UNSUPPORTED, ERR_NO_CANDIDATE, ERR_STATE = 0, -1, -2, -3, -4, -5

FIELD_NAMES = ("d", "d_d", "d_dd", "d_ddd", "s", "s_d", "s_dd", "s_ddd", "x", "y", "yaw", "v", "curv")
RED_NAMES = ("ego_risk", "obst_risk", "ego_harm", "obst_harm")
COST_NAMES = ("bayes", "equality", "maximin", "responsibility", "ego", "risk_cost", "velocity",
              "dist_to_global_path", "lon_jerk", "lat_jerk", "travelled_dist", "factor", "total", "dist_to_goal_pos")
REASON_NAMES = ("", "velocity", "acceleration", "curvature (lvc)", "curvature (hvc)", "collision",
                "boundaries", "max_risk")

# array arguments cross as plain addresses (const double* / const int32_t* / const uint8_t* in the header): an integer
# from ndarray.ctypes.data is several times cheaper to produce than a typed ctypes pointer, and the closed loop passes a
# dozen of them per step
_dp = C.c_void_p
_ip = C.c_void_p
_bp = C.c_void_p


COLLISION_HOST, COLLISION_DISCRETE, COLLISION_SWEPT = 0, 1, 2
COLLISION_CODES = {"host": COLLISION_HOST, "discrete": COLLISION_DISCRETE, "swept": COLLISION_SWEPT}
# ETP_RELAX_* bits (include/etp_b200.h): opt-outs of reference quirks, strict reproduction is the default
RELAX_BITS = {"curvature_denominator": 1, "acceleration_check": 2, "maximin_gate": 4, "view_cone_wrap": 8}


class Params(C.Structure):
    _fields_ = [
        ("weights", C.c_double * NUM_WEIGHTS),
        ("planner_mode", C.c_int32),
        ("fast_prob_mahalanobis", C.c_int32),
        ("multiple_cost_functions", C.c_int32),
        ("n_angle_thresholds", C.c_int32),
        ("max_acceptable_risk", C.c_double),
        ("angle_thr", C.c_double * MAX_ANGLE_BINS),
        ("angle_coef_pos", C.c_double * (MAX_ANGLE_BINS + 1)),
        ("angle_coef_neg", C.c_double * (MAX_ANGLE_BINS + 1)),
        ("lr_const", C.c_double), ("lr_speed", C.c_double),
        ("lr1s_const", C.c_double), ("lr1s_speed", C.c_double),
        ("ped_const", C.c_double), ("ped_speed", C.c_double),
        ("veh_l", C.c_double), ("veh_w", C.c_double), ("veh_l_r", C.c_double), ("veh_m", C.c_double),
        ("veh_steer_max", C.c_double), ("veh_lat_a_max", C.c_double), ("veh_v_max", C.c_double),
        ("veh_a_max", C.c_double),
        ("collision_check", C.c_int32), ("relaxed_quirks", C.c_int32),
    ]


class Obstacles(C.Structure):
    _fields_ = [
        ("n_obstacles", C.c_int32), ("max_pred", C.c_int32),
        ("pred_len", _ip), ("pos", _dp), ("cov", _dp), ("yaw", _dp), ("vel", _dp),
        ("length", _dp), ("width", _dp), ("mass", _dp), ("is_protected", _ip), ("responsibility", _ip),
    ]


class RawPredictions(C.Structure):
    _fields_ = [
        ("n_obstacles", C.c_int32), ("max_pred", C.c_int32),
        ("pred_len", _ip), ("pos", _dp), ("cov", _dp), ("length", _dp), ("width", _dp),
        ("init_orientation", _dp), ("init_velocity", _dp), ("mass", _dp), ("is_protected", _ip),
        ("dt", C.c_double), ("safety_margin_length", C.c_double), ("safety_margin_width", C.c_double),
        ("ego_x", C.c_double), ("ego_y", C.c_double), ("ego_orientation", C.c_double),
    ]


class RoadBoundary(C.Structure):
    _fields_ = [("n_triangles", C.c_int32), ("triangles", _dp), ("check", C.c_int32)]


class Goal(C.Structure):
    _fields_ = [
        ("has_area", C.c_int32), ("n_polygons", C.c_int32),
        ("poly_start", _ip), ("vertices", _dp),
        ("n_circles", C.c_int32), ("circles", _dp),
        ("has_velocity", C.c_int32), ("velocity_start", C.c_double), ("velocity_end", C.c_double),
        ("has_orientation", C.c_int32), ("orientation_start", C.c_double), ("orientation_end", C.c_double),
        ("time_start", C.c_int32), ("time_end", C.c_int32), ("ego_time_step", C.c_int32),
        ("ego_x", C.c_double), ("ego_y", C.c_double),
        ("has_centers", C.c_int32), ("avg_goal_distance", C.c_double),
        ("n_goal_lines", C.c_int32), ("goal_line_start", _ip), ("goal_line_vertices", _dp),
        ("n_goal_points", C.c_int32), ("goal_points", _dp),
    ]


class Step(C.Structure):
    _fields_ = [
        ("c_s", C.c_double), ("c_s_d", C.c_double), ("c_s_dd", C.c_double),
        ("c_d", C.c_double), ("c_d_d", C.c_double), ("c_d_dd", C.c_double),
        ("v_list", _dp), ("nv", C.c_int32),
        ("t_list", _dp), ("nt", C.c_int32),
        ("n_samples", _ip),
        ("d_list", _dp), ("nd", C.c_int32),
        ("dt", C.c_double), ("v_thr", C.c_double),
        ("velocity_cost_mode", C.c_int32),
        ("velocity_target", C.c_double),
        ("cost_factor", C.c_double),
        ("ext_level", _bp),
        ("bd_harm", _dp),
        ("cost_factor_list", _dp),
        ("c_begin", C.c_int32), ("c_end", C.c_int32),
        ("shard_index", C.c_int32), ("shard_count", C.c_int32),
        ("goal", C.POINTER(Goal)),
    ]


class Out(C.Structure):
    _fields_ = [
        ("precision", C.c_int32),
        ("traj", C.c_void_p), ("prob", C.c_void_p), ("red", C.c_void_p), ("cost", C.c_void_p),
        ("level", C.c_void_p), ("reason", C.c_void_p),
        ("c_stride", C.c_int64),
        ("harm", C.c_void_p),
        ("bd_harm", C.c_void_p),
    ]


class Trajectories(C.Structure):
    _fields_ = [
        ("n_traj", C.c_int32), ("t_max", C.c_int32),
        ("n_samples", _ip), ("fields", _dp),
        ("velocity_cost_mode", C.c_int32),
        ("velocity_target", C.c_double), ("cost_factor", C.c_double),
        ("ext_level", _bp), ("bd_harm", _dp), ("cost_factor_list", _dp),
        ("goal", C.POINTER(Goal)), ("dt", C.c_double),
    ]


class Scenario(C.Structure):
    _fields_ = [
        ("step", C.POINTER(Step)), ("obstacles", C.POINTER(Obstacles)),
        ("knots", _dp), ("coef_x", _dp), ("coef_y", _dp), ("n_seg", C.c_int32),
    ]


class Cycle(C.Structure):
    _fields_ = [("raw", C.POINTER(RawPredictions)), ("obstacles", C.POINTER(Obstacles)), ("step", C.POINTER(Step))]


class Best(C.Structure):
    _fields_ = [
        ("key_hi", C.c_uint64), ("key_lo", C.c_uint64), ("cost", C.c_double),
        ("index", C.c_int32), ("level", C.c_int32), ("n_scored", C.c_int32), ("list_index", C.c_int32),
    ]


# every symbol include/etp_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "etp_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "etp_destroy": (C.c_int, [C.c_void_p]),
    "etp_last_error": (C.c_char_p, [C.c_void_p]),
    "etp_version": (C.c_char_p, []),
    "etp_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "etp_set_params": (C.c_int, [C.c_void_p, C.POINTER(Params)]),
    "etp_set_spline": (C.c_int, [C.c_void_p, _dp, _dp, _dp, C.c_int]),
    "etp_set_road_boundary": (C.c_int, [C.c_void_p, C.POINTER(RoadBoundary)]),
    "etp_set_obstacles": (C.c_int, [C.c_void_p, C.POINTER(Obstacles), C.c_int]),
    "etp_set_raw_predictions": (C.c_int, [C.c_void_p, C.POINTER(RawPredictions), C.c_int]),
    "etp_get_enriched": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp, _ip]),
    "etp_plan_step": (C.c_int, [C.c_void_p, C.POINTER(Step), C.POINTER(Out)]),
    "etp_score_trajectories": (C.c_int, [C.c_void_p, C.POINTER(Trajectories), C.POINTER(Out)]),
    "etp_principle_costs": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _ip, C.c_double, C.c_double, C.c_double, _dp]),
    "etp_plan_batch": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Scenario), C.c_int, C.POINTER(Best)]),
    "etp_plan_cycle": (C.c_int, [C.c_void_p, C.POINTER(Cycle), C.c_int, C.POINTER(Best), _dp, _ip]),
    "etp_tmax": (C.c_int, [C.POINTER(Step)]),
    "etp_get_best": (C.c_int, [C.c_void_p, C.POINTER(Best)]),
    "etp_get_trajectory": (C.c_int, [C.c_void_p, C.c_int, _dp, _ip]),
    "etp_get_best_trajectory": (C.c_int, [C.c_void_p, C.POINTER(Best), _dp, _ip]),
    "etp_best_device_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "etp_enable_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "etp_get_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "etp_get_stage_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int]),
    "etp_debug_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_int]),
    "etp_combine_best": (C.c_int, [C.POINTER(Best), C.c_int, C.c_int, C.POINTER(Best)]),
    "etp_combine_best_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
}


NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]


def nvcc_command(out_path=LIB_PATH, extra=()):
    """One-shot form of the build (what build_library does file by file, in parallel)."""
    return ["nvcc", *NVCC_FLAGS, "-shared", "-o", out_path, *SOURCES, *extra]


def library_is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build_library(force=False, verbose=False, extra=(), out_path=None):
    """Compile the kernels + C ABI in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    if not force and not library_is_stale():
        return LIB_PATH
    # every translation unit on its own nvcc process (the fused kernel alone takes a minute), then one link
    build_dir = os.path.join(HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    extra = list(extra)
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(build_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = ["nvcc", *NVCC_FLAGS, *extra, "-c", "-o", obj, src]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = ""
    for cmd, pr in procs:
        out, _ = pr.communicate()
        log += out
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out)
    cmd = ["nvcc", *NVCC_FLAGS, "-shared", "-o", out_path or LIB_PATH, *objs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(log + res.stdout + res.stderr)
    return out_path or LIB_PATH


_LIB = None


def load_library():
    """dlopen the CUDA library; raises (no fallback) when it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the scoring path has no CPU fallback. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib
