"""Host side of the fused scoring path: packs the reference's data contracts into the C ABI,
owns the device tensors (PyTorch) and returns the selection.

PyTorch is used for device memory and streams only; all arithmetic of the path runs in the
hand-written kernels of ``csrc/`` reached through ``libetp_b200.so``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _abi
from .configs import WEIGHT_KEYS
from .spline import CubicSpline2D
from .synthetic import PROTECTED_TYPES, UNPROTECTED_TYPES, PlanningStep, obstacle_mass

MODE_CODES = {"risk": 0, "WaleNet": 1, "ground_truth": 2}


class EtpError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"etp_b200 error {code}: {message}")
        self.code = code


class NoCandidateError(ValueError):
    """Every candidate was skipped: the reference's ``max([])`` ValueError (frenet_functions.py:562-564)."""


def _struct_dtype(ctype):
    """numpy record dtype with the memory layout of a ctypes Structure (pointers as unsigned 64-bit addresses)."""
    kinds = {C.c_double: "f8", C.c_float: "f4", C.c_int32: "i4", C.c_int64: "i8", C.c_uint64: "u8", C.c_void_p: "u8"}
    names, formats, offsets = [], [], []
    for name, typ in ctype._fields_:
        names.append(name)
        formats.append(kinds.get(typ, "u8"))  # anything else in these structs is a pointer
        offsets.append(getattr(ctype, name).offset)
    return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": C.sizeof(ctype)})


def _dptr(a):
    return a.ctypes.data  # address of a float64 array the caller keeps alive


def _iptr(a):
    return a.ctypes.data  # address of an int32 array the caller keeps alive


def harm_model_tables(modes, coeffs):
    """Thresholds on |angle| and per-band coefficients of the protected-occupant model selected by
    risk.json (harm_estimation.py:358-400; logistic_regression_symmetrical.py / _asymmetrical.py)."""
    if modes["harm_mode"] != "log_reg":
        if modes["harm_mode"] in ("ref_speed", "gidas"):
            raise NotImplementedError(
                f"harm_mode {modes['harm_mode']!r} is not part of the scored path: the reference's "
                "own implementations raise on array input (SURVEY.md Q10, Q11)")
        raise ValueError("Please select a valid mode for harm estimation (log_reg, ref_speed, gidas)")
    if modes["crash_angle_simplified"] is False:
        raise NotImplementedError("comprehensive crash-angle estimation needs pycrcc and is out of scope")
    lr = coeffs["log_reg"]
    if modes["ignore_angle"]:
        p = lr["ignore_angle"]
        return [], [0.0], [0.0], p["const"], p["speed"]
    sym, red = modes["sym_angle"], modes["reduced_angle_areas"]
    if red and sym:
        p = lr["reduced_sym_angle_areas"]
        t_a = 45 / 180 * np.pi
        thr = [t_a, 3 * t_a]
        pos = neg = [0.0, p["side"], p["rear"]]
    elif red:
        p = lr["reduced_angle_areas"]
        thr = [45 / 180 * np.pi, 135 / 180 * np.pi]
        pos = [0.0, p["driver_side"], p["rear"]]
        neg = [0.0, p["right_side"], p["rear"]]
    elif sym:
        p = lr["complete_sym_angle_areas"]
        thr = [k / 180 * np.pi for k in (15, 45, 75, 105, 135, 165)]
        pos = neg = [0.0, p["Imp_1_11"], p["Imp_2_10"], p["Imp_3_9"], p["Imp_4_8"], p["Imp_5_7"], p["Imp_6"]]
    else:
        p = lr["complete_angle_areas"]
        thr = [k / 180 * np.pi for k in (15, 45, 75, 105, 135, 165)]
        pos = [0.0, p["Imp_11"], p["Imp_10"], p["Imp_9"], p["Imp_8"], p["Imp_7"], p["Imp_6"]]
        neg = [0.0, p["Imp_1"], p["Imp_2"], p["Imp_3"], p["Imp_4"], p["Imp_5"], p["Imp_6"]]
    return thr, list(pos), list(neg), p["const"], p["speed"]


def make_params(params, vehicle, mode="risk") -> _abi.Params:
    """Flatten params_dict (frenet_planner.py:149-153) + vehicle constants into the ABI struct."""
    weights, modes, coeffs = params["weights"], params["modes"], params["harm"]
    thr, pos, neg, const, speed = harm_model_tables(modes, coeffs)
    p = _abi.Params()
    for i, k in enumerate(WEIGHT_KEYS):
        p.weights[i] = float(weights.get(k, 0.0))
    if mode not in MODE_CODES:
        raise ValueError("mode must be ground_truth, WaleNet, or risk")
    p.planner_mode = MODE_CODES[mode]
    p.fast_prob_mahalanobis = int(bool(modes["fast_prob_mahalanobis"]))
    p.multiple_cost_functions = int(bool(modes["multiple_cost_functions"]))
    p.n_angle_thresholds = len(thr)
    p.max_acceptable_risk = float(modes["max_acceptable_risk"])
    for i, v in enumerate(thr):
        p.angle_thr[i] = v
    for i in range(_abi.MAX_ANGLE_BINS + 1):
        p.angle_coef_pos[i] = pos[min(i, len(pos) - 1)]
        p.angle_coef_neg[i] = neg[min(i, len(neg) - 1)]
    p.lr_const, p.lr_speed = const, speed
    ia = coeffs["log_reg"]["ignore_angle"]
    p.lr1s_const, p.lr1s_speed = ia["const"], ia["speed"]
    p.ped_const, p.ped_speed = coeffs["pedestrian"]["const"], coeffs["pedestrian"]["speed"]
    p.veh_l, p.veh_w, p.veh_l_r, p.veh_m = vehicle.l, vehicle.w, vehicle.l_r, vehicle.m
    p.veh_steer_max = vehicle.steering.max
    p.veh_lat_a_max = vehicle.lateral_a_max
    p.veh_v_max = vehicle.longitudinal.v_max
    p.veh_a_max = vehicle.longitudinal.a_max
    # extension key (not in the reference's risk.json): test the candidates against the predicted boxes on the device
    # instead of taking pycrcc's answer through ext_level ("host", the default) -- ETP_COLLISION_* in etp_b200.h
    check = modes.get("collision_check_device", "host")
    if check not in _abi.COLLISION_CODES:
        raise ValueError("modes['collision_check_device'] must be host, discrete or swept")
    p.collision_check = _abi.COLLISION_CODES[check]
    # extension key: reference quirks to opt out of; none by default = strict reference behaviour
    for name in modes.get("relaxed_quirks", ()):
        if name not in _abi.RELAX_BITS:
            raise ValueError(f"modes['relaxed_quirks']: unknown quirk {name!r} (known: {sorted(_abi.RELAX_BITS)})")
        p.relaxed_quirks |= _abi.RELAX_BITS[name]
    return p


def make_goal(goal, keep):
    """GoalContext -> etp_goal; the arrays it points to are appended to ``keep``."""
    g = _abi.Goal()
    polys = [np.ascontiguousarray(p, dtype=np.float64).reshape(-1, 2) for p in (goal.polygons or [])]
    circles = np.ascontiguousarray(goal.circles if goal.circles else np.zeros((0, 3)), dtype=np.float64).reshape(-1, 3)
    g.has_area = int(goal.has_area)
    g.n_polygons, g.n_circles = len(polys), len(circles)
    start = np.ascontiguousarray(np.concatenate([[0], np.cumsum([len(p) for p in polys])]), dtype=np.int32)
    verts = np.ascontiguousarray(np.concatenate(polys) if polys else np.zeros((0, 2)), dtype=np.float64)
    g.poly_start, g.vertices, g.circles = _iptr(start), _dptr(verts), _dptr(circles)
    keep += [start, verts, circles]
    if goal.velocity is not None:
        g.has_velocity, g.velocity_start, g.velocity_end = 1, float(goal.velocity[0]), float(goal.velocity[1])
    if goal.orientation is not None:
        g.has_orientation, g.orientation_start, g.orientation_end = 1, float(goal.orientation[0]), float(goal.orientation[1])
    g.time_start, g.time_end = int(goal.time_step[0]), int(goal.time_step[1])
    g.ego_time_step = int(goal.ego_time_step)
    g.ego_x, g.ego_y = float(goal.ego_position[0]), float(goal.ego_position[1])
    if goal.goal_centers is not None and len(goal.goal_centers) > 0:
        pos = np.asarray(goal.ego_position, dtype=np.float64)
        # distance() of helper_functions.py:114 is the Euclidean norm; np.mean over the list (:492-497)
        g.has_centers = 1
        g.avg_goal_distance = float(np.mean([np.linalg.norm(np.asarray(c, dtype=np.float64) - pos) for c in goal.goal_centers]))
    elif goal.goal_centers is not None:
        g.has_centers = -1  # a goal position without centre points: NaN in the reference (:492-497), refused by the library
    # goal positions of calc_dist_to_goal_pos (calc_trajectory_cost.py:760-803)
    lines = [np.ascontiguousarray(l, dtype=np.float64).reshape(-1, 2) for l in (getattr(goal, "goal_lines", None) or [])]
    if lines:
        ls = np.ascontiguousarray(np.concatenate([[0], np.cumsum([len(l) for l in lines])]), dtype=np.int32)
        lv = np.ascontiguousarray(np.concatenate(lines), dtype=np.float64)
        g.n_goal_lines, g.goal_line_start, g.goal_line_vertices = len(lines), _iptr(ls), _dptr(lv)
        keep += [ls, lv]
    pts = getattr(goal, "goal_points", None)
    if pts is not None and len(pts) > 0:
        pa = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 2)
        g.n_goal_points, g.goal_points = len(pa), _dptr(pa)
        keep.append(pa)
    return g


def pack_predictions(predictions, obstacle_types=None, masses=None, protections=None):
    """`predictions` dict (SURVEY.md 8b data contract 1) -> SoA host arrays, dict order kept.

    Obstacle class comes from ``obstacle_types[id]`` (an ObstacleType name such as "CAR") like the
    reference's ``scenario.obstacle_by_id(id).obstacle_type`` lookups (harm_estimation.py:258-260,356),
    or from explicit ``masses`` / ``protections``.
    """
    oids = list(predictions.keys())
    O = len(oids)
    lens = np.array([len(predictions[o]["pos_list"]) for o in oids], dtype=np.int32)
    Tp = int(lens.max()) if O else 1
    pos = np.zeros((O, Tp, 2))
    cov = np.zeros((O, Tp, 4))
    yaw = np.zeros((O, Tp))
    vel = np.zeros((O, Tp))
    length = np.zeros(O)
    width = np.zeros(O)
    mass = np.zeros(O)
    prot = np.zeros(O, dtype=np.int32)
    resp = np.zeros(O, dtype=np.int32)
    preds = [predictions[oid] for oid in oids]
    uniform = O > 0 and int(lens.min()) == Tp
    if uniform:  # equal prediction lengths (the usual case): four stacked copies instead of 4 * O sliced ones
        pos[:] = np.stack([np.asarray(p["pos_list"], dtype=np.float64) for p in preds]).reshape(O, Tp, 2)
        cov[:] = np.stack([np.asarray(p["cov_list"], dtype=np.float64) for p in preds]).reshape(O, Tp, 4)
        yaw[:] = np.stack([np.asarray(p["orientation_list"], dtype=np.float64)[:Tp] for p in preds])
        vel[:] = np.stack([np.asarray(p["v_list"], dtype=np.float64)[:Tp] for p in preds])
    for j, oid in enumerate(oids):
        p = preds[j]
        n = lens[j]
        if not uniform:
            pos[j, :n] = np.asarray(p["pos_list"], dtype=np.float64).reshape(n, 2)
            cov[j, :n] = np.asarray(p["cov_list"], dtype=np.float64).reshape(n, 4)
            yaw[j, :n] = np.asarray(p["orientation_list"], dtype=np.float64)[:n]
            vel[j, :n] = np.asarray(p["v_list"], dtype=np.float64)[:n]
        length[j] = p["shape"]["length"]
        width[j] = p["shape"]["width"]
        resp[j] = int(p.get("responsibility", 0))
        if masses is not None:
            mass[j] = masses[oid]
            prot[j] = int(protections[oid])
        else:
            t = obstacle_types[oid]
            t = getattr(t, "name", t)
            if t in PROTECTED_TYPES:
                prot[j] = 1
            elif t in UNPROTECTED_TYPES:
                prot[j] = 0
            else:
                raise ValueError(f"obstacle type {t!r} has no protection class (harm_estimation.py:52-56)")
            mass[j] = obstacle_mass(t, length[j] * width[j])
    return dict(ids=oids, O=O, Tp=Tp, pred_len=lens, pos=pos, cov=cov, yaw=yaw, vel=vel, length=length,
                width=width, mass=mass, prot=prot, resp=resp)


@dataclass
class ScoreResult:
    """Outputs of one planning step.  Tensors live on the device; ``numpy()`` copies them."""

    n_dense: int
    c_begin: int
    c_end: int
    Tmax: int
    n_samples: np.ndarray
    obstacle_ids: list
    best: Optional[dict]
    traj: object = None
    prob: object = None
    red: object = None
    cost: object = None
    level: object = None
    reason: object = None
    bd_harm: object = None
    _ctx: object = field(default=None, repr=False)

    def numpy(self):
        """Host copies in the logical (candidate-major) layout: traj [13][C][T], prob [C][T-1][O],
        red [4][C][O], cost [13][C], level [C], reason [C].  On the device every real-valued tensor is
        candidate-minor (include/etp_b200.h, etp_out); the transposes below are views."""
        out = {}
        self._ctx.stream.synchronize()
        for k in ("traj", "prob", "red", "cost", "level", "reason", "bd_harm"):
            v = getattr(self, k)
            out[k] = None if v is None else v.detach().cpu().numpy()
        if out["traj"] is not None:
            out["traj"] = out["traj"].transpose(0, 2, 1)
        if out["prob"] is not None:
            out["prob"] = out["prob"].transpose(2, 0, 1)
        if out["red"] is not None:
            out["red"] = out["red"].transpose(0, 2, 1)
        return out

    def best_trajectory(self):
        """The 11 arrays FrenetPlanner stores in ``self._trajectory`` (frenet_planner.py:564-576),
        recomputed in fp64 on the device.  Note ``v_mps`` is ``s_d`` (quirk Q7)."""
        f, T = self._ctx.trajectory(self.best["index"])
        names = _abi.FIELD_NAMES
        g = {n: f[i, :T] for i, n in enumerate(names)}
        dt = self._ctx.last_dt
        return {
            "s_loc_m": g["s"], "d_loc_m": g["d"], "d_d_loc_mps": g["d_d"], "d_dd_loc_mps2": g["d_dd"],
            "x_m": g["x"], "y_m": g["y"], "psi_rad": g["yaw"], "kappa_radpm": g["curv"],
            "v_mps": g["s_d"], "ax_mps2": g["s_dd"], "time_s": np.arange(T) * dt,
        }


class ScoringContext:
    """One ``etp_ctx``: one device, one stream, reusable across planning steps."""

    def __init__(self, device=0, precision="fp32"):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("the scoring path runs on CUDA only (no CPU fallback)")
        self.torch = torch
        self.lib = _abi.load_library()
        self.device = torch.device("cuda", device)
        self.precision = {"fp32": _abi.PRECISION_FP32, "fp64": _abi.PRECISION_FP64}[precision]
        self.dtype = torch.float32 if self.precision == _abi.PRECISION_FP32 else torch.float64
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)  # make sure the primary context exists
        h = C.c_void_p()
        rc = self.lib.etp_create(C.byref(h), device)
        if rc != 0:
            raise EtpError(rc, "etp_create failed")
        self.h = h
        # a dedicated (non-default) stream: every launch, copy and event of this context is ordered on it
        self.stream = torch.cuda.Stream(device=self.device)
        self._check(self.lib.etp_set_stream(self.h, C.c_void_p(self.stream.cuda_stream)))
        self.obstacle_ids = []
        self.O = 0
        self.has_pred = False
        self.last_dt = 0.1
        self._keep = []

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.etp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc == _abi.OK:
            return
        msg = self.lib.etp_last_error(self.h).decode()
        if rc == _abi.ERR_NO_CANDIDATE:
            raise NoCandidateError(msg)
        if rc == _abi.ERR_UNSUPPORTED:
            raise NotImplementedError(msg)
        raise EtpError(rc, msg)

    # -- inputs -----------------------------------------------------------------------------
    def set_params(self, params, vehicle, mode="risk"):
        self._params_key = None
        p = make_params(params, vehicle, mode)
        self._check(self.lib.etp_set_params(self.h, C.byref(p)))

    def set_spline(self, csp):
        self._spline_key = None
        csp = CubicSpline2D.from_foreign(csp)
        k = np.ascontiguousarray(csp.s, dtype=np.float64)
        cx = np.ascontiguousarray(csp.coef_x, dtype=np.float64)
        cy = np.ascontiguousarray(csp.coef_y, dtype=np.float64)
        self._check(self.lib.etp_set_spline(self.h, _dptr(k), _dptr(cx), _dptr(cy), len(k) - 1))

    def set_road_boundary(self, triangles=None, check="swept"):
        """Road boundary of the scenario as triangles [n, 3, 2] (the region outside the road, e.g.
        ``boundary.create_road_boundary_obstacle(method="aligned_triangulation")``, frenet_planner.py:226-236); every
        later step tests the candidates against it on the device (level 2 and boundary harm).  ``None`` removes it."""
        b = _abi.RoadBoundary()
        tri = np.zeros((0, 3, 2)) if triangles is None else np.ascontiguousarray(triangles, dtype=np.float64).reshape(-1, 3, 2)
        b.n_triangles, b.triangles, b.check = len(tri), _dptr(tri), _abi.COLLISION_CODES[check]
        self._check(self.lib.etp_set_road_boundary(self.h, C.byref(b)))
        # what is resident: the caller's object (kept alive, so its identity cannot be recycled) and its content
        self._boundary_key = None if triangles is None else (triangles, tri.tobytes(), check)

    def ensure_road_boundary(self, rb):
        """``rb`` = (triangles, check) or None: upload unless exactly this boundary is resident (same object and same
        bytes, or equal bytes of another object)."""
        cur = getattr(self, "_boundary_key", None)
        if rb is None:
            if cur is not None:
                self.set_road_boundary(None)
            return
        tri, check = rb
        if cur is not None and cur[2] == check:
            raw = np.ascontiguousarray(tri, dtype=np.float64).tobytes()
            if raw == cur[1]:
                return
        self.set_road_boundary(tri, check)

    def set_predictions(self, predictions, obstacle_types=None, masses=None, protections=None):
        o = _abi.Obstacles()
        if predictions is None:
            o.n_obstacles = -1
            self.obstacle_ids, self.O, self.has_pred = [], 0, False
            self._check(self.lib.etp_set_obstacles(self.h, C.byref(o), self.precision))
            return
        pk = pack_predictions(predictions, obstacle_types, masses, protections)
        o.n_obstacles, o.max_pred = pk["O"], pk["Tp"]
        o.pred_len = _iptr(pk["pred_len"])
        o.pos, o.cov, o.yaw, o.vel = _dptr(pk["pos"]), _dptr(pk["cov"]), _dptr(pk["yaw"]), _dptr(pk["vel"])
        o.length, o.width, o.mass = _dptr(pk["length"]), _dptr(pk["width"]), _dptr(pk["mass"])
        o.is_protected, o.responsibility = _iptr(pk["prot"]), _iptr(pk["resp"])
        self._check(self.lib.etp_set_obstacles(self.h, C.byref(o), self.precision))
        self.obstacle_ids, self.O, self.has_pred = pk["ids"], pk["O"], True

    def set_raw_predictions(self, predictions, shapes, init_orientation, init_velocity, obstacle_types, dt, ego_position,
                            ego_orientation, safety_margin_length=1.0, safety_margin_width=0.5, masses=None, protections=None,
                            upload=True):
        """Raw predictions ``{id: {'pos_list', 'cov_list'}}`` straight from the predictor: orientation, velocity, shape
        margins and the action-space responsibility flag are derived on the device (prediction_helpers.py:75-135,
        responsibility.py:57-89) and the obstacle tile is packed from them.  ``shapes[id] = (length, width)`` of the raw
        obstacle; empty predictions are dropped like the reference does (:98-100).

        ``upload=False`` only prepares the ``etp_raw_predictions`` record and returns it (``plan_cycle`` hands it to the
        library together with the step)."""
        # closed loop: the same obstacles with the same static description (the caller hands over the very same dict objects,
        # kept alive here) and equally long predictions as in the previous cycle -> only positions, covariances and the ego
        # pose change; the struct and the static arrays of the previous cycle are used again
        fast = getattr(self, "_raw_fast", None)
        if (fast is not None and masses is None and fast["refs"][0] is shapes and fast["refs"][1] is init_orientation
                and fast["refs"][2] is init_velocity and fast["refs"][3] is obstacle_types and fast["dt"] == dt
                and fast["margins"] == (safety_margin_length, safety_margin_width) and tuple(predictions) == fast["oids"]):
            vals = list(predictions.values())
            Tp = fast["Tp"]
            if all(len(p["pos_list"]) == Tp for p in vals):
                try:  # arrays of one shape (the usual predictor output): one concatenate each into the resident arrays
                    np.concatenate([p["pos_list"] for p in vals], out=fast["pos2"])
                    np.concatenate([p["cov_list"] for p in vals], out=fast["cov2"])
                except (ValueError, TypeError):  # lists, or another layout of the same data
                    np.stack([p["pos_list"] for p in vals], out=fast["pos"].reshape(fast["pos_shape"]))
                    np.stack([p["cov_list"] for p in vals], out=fast["cov"].reshape(fast["cov_shape"]))
                r = fast["struct"]
                r.ego_x, r.ego_y, r.ego_orientation = float(ego_position[0]), float(ego_position[1]), float(ego_orientation)
                if upload:
                    self._check(self.lib.etp_set_raw_predictions(self.h, C.byref(r), self.precision))
                return r
        self._raw_fast = None
        oids = [oid for oid in predictions if len(predictions[oid]["pos_list"]) > 0]
        O = len(oids)
        lens = np.array([len(predictions[o]["pos_list"]) for o in oids], dtype=np.int32)
        Tp = int(lens.max()) if O else 1
        if O and int(lens.min()) == Tp:  # equally long predictions (the usual predictor output): one stack each
            pos = np.ascontiguousarray(np.stack([predictions[o]["pos_list"] for o in oids]), dtype=np.float64).reshape(O, Tp, 2)
            cov = np.ascontiguousarray(np.stack([predictions[o]["cov_list"] for o in oids]), dtype=np.float64).reshape(O, Tp, 4)
        else:
            pos, cov = np.zeros((O, Tp, 2)), np.zeros((O, Tp, 4))
            for j, oid in enumerate(oids):
                n = lens[j]
                pos[j, :n] = np.asarray(predictions[oid]["pos_list"], dtype=np.float64).reshape(n, 2)
                cov[j, :n] = np.asarray(predictions[oid]["cov_list"], dtype=np.float64).reshape(n, 4)
        # what does not change from cycle to cycle (shape, initial state, class, mass) is kept per obstacle
        cache = self.__dict__.setdefault("_raw_static", {})
        static = np.empty((6, O))
        for j, oid in enumerate(oids):
            if masses is not None:
                key = (shapes[oid], init_orientation[oid], init_velocity[oid], masses[oid], int(protections[oid]))
            else:
                key = (shapes[oid], init_orientation[oid], init_velocity[oid], obstacle_types[oid], safety_margin_length, safety_margin_width)
            hit = cache.get(oid)
            if hit is None or hit[0] != key:
                ln, wd = shapes[oid]
                if masses is not None:
                    m, pr = masses[oid], int(protections[oid])
                else:
                    t = getattr(obstacle_types[oid], "name", obstacle_types[oid])
                    if t in PROTECTED_TYPES:
                        pr = 1
                    elif t in UNPROTECTED_TYPES:
                        pr = 0
                    else:
                        raise ValueError(f"obstacle type {t!r} has no protection class (harm_estimation.py:52-56)")
                    m = obstacle_mass(t, (ln + safety_margin_length) * (wd + safety_margin_width))
                hit = (key, (ln, wd, init_orientation[oid], init_velocity[oid], m, pr))
                cache[oid] = hit
            static[:, j] = hit[1]
        length, width, ori0, vel0, mass = (np.ascontiguousarray(static[k]) for k in range(5))
        prot = static[5].astype(np.int32)
        r = _abi.RawPredictions()
        r.n_obstacles, r.max_pred = O, Tp
        r.pred_len, r.pos, r.cov = _iptr(lens), _dptr(pos), _dptr(cov)
        r.length, r.width, r.init_orientation, r.init_velocity = _dptr(length), _dptr(width), _dptr(ori0), _dptr(vel0)
        r.mass, r.is_protected = _dptr(mass), _iptr(prot)
        r.dt, r.safety_margin_length, r.safety_margin_width = float(dt), float(safety_margin_length), float(safety_margin_width)
        r.ego_x, r.ego_y, r.ego_orientation = float(ego_position[0]), float(ego_position[1]), float(ego_orientation)
        if upload:
            self._check(self.lib.etp_set_raw_predictions(self.h, C.byref(r), self.precision))
        self.obstacle_ids, self.O, self.has_pred = oids, O, True
        self._raw_shape = (O, Tp, lens)
        self._raw_keep = (r, lens, pos, cov, length, width, ori0, vel0, mass, prot)
        if O and masses is None and int(lens.min()) == Tp and len(oids) == len(predictions):
            p0 = predictions[oids[0]]
            self._raw_fast = dict(refs=(shapes, init_orientation, init_velocity, obstacle_types), dt=dt,
                                  margins=(safety_margin_length, safety_margin_width), oids=tuple(oids), Tp=Tp, pos=pos, cov=cov,
                                  pos_shape=(O,) + np.shape(p0["pos_list"]), cov_shape=(O,) + np.shape(p0["cov_list"]), struct=r,
                                  pos2=pos.reshape((O * Tp,) + np.shape(p0["pos_list"])[1:]),
                                  cov2=cov.reshape((O * Tp,) + np.shape(p0["cov_list"])[1:]),
                                  keep=(lens, length, width, ori0, vel0, mass, prot))
        return r

    def enriched(self, predictions):
        """Read the derived arrays of the last ``set_raw_predictions`` back into the prediction dict (what the reference's
        log line prints): 'orientation_list', 'v_list', 'shape', 'responsibility'.  Empty predictions are deleted."""
        O, Tp, lens = self._raw_shape
        yaw, vel = np.zeros((O, Tp)), np.zeros((O, Tp))
        length, width = np.zeros(O), np.zeros(O)
        resp = np.zeros(O, dtype=np.int32)
        self._check(self.lib.etp_get_enriched(self.h, _dptr(yaw), _dptr(vel), _dptr(length), _dptr(width), _iptr(resp)))
        for oid in [o for o in predictions if len(predictions[o]["pos_list"]) == 0]:
            del predictions[oid]
        for j, oid in enumerate(self.obstacle_ids):
            n = int(lens[j])
            predictions[oid]["orientation_list"] = yaw[j, :n].copy()
            predictions[oid]["v_list"] = vel[j, :n].copy()
            predictions[oid]["shape"] = {"length": float(length[j]), "width": float(width[j])}
            predictions[oid]["responsibility"] = int(resp[j])
        return predictions

    def ensure_params(self, params, vehicle, mode="risk"):
        """set_params unless the same parameter objects with the same content are already on the device."""
        # identity of the objects plus their scalar content (dict order is stable): cheap enough for every cycle of a
        # closed loop, and an in-place edit of a weight or a mode still triggers the upload
        # every value make_params reads, as plain lists (compared with ==, never hashed: nested dicts are fine)
        harm = params["harm"]
        pkey = (mode, list(params["weights"].values()), list(params["modes"].values()),
                [list(t.values()) for t in harm["log_reg"].values()], list(harm["pedestrian"].values()),
                vehicle.l, vehicle.w, vehicle.l_r, vehicle.m, vehicle.steering.max, vehicle.lateral_a_max,
                vehicle.longitudinal.v_max, vehicle.longitudinal.a_max)
        if getattr(self, "_params_key", None) != pkey:
            self.set_params(params, vehicle, mode)
            self._params_key = pkey

    def configure(self, step: PlanningStep):
        """params + spline + predictions of a PlanningStep.  Parameters and spline are uploaded again only when they
        changed (same objects, same content fingerprint); predictions change every step and always are."""
        self.ensure_params(step.params, step.vehicle, step.mode)
        # the resident spline is recognised by the caller's object (kept alive here, so its identity cannot be recycled by
        # another spline) plus a content fingerprint that catches in-place edits
        cur = getattr(self, "_spline_key", None)
        if cur is None or cur[0] is not step.csp or cur[1] != self._spline_print(step.csp):
            csp = CubicSpline2D.from_foreign(step.csp)
            self.set_spline(csp)
            self._spline_key = (step.csp, self._spline_print(step.csp))
        if not getattr(step, "predictions_resident", False):
            self.set_predictions(step.predictions, step.obstacle_types)
        self.ensure_road_boundary(getattr(step, "road_boundary", None))  # (triangles, check) or None

    @staticmethod
    def _spline_print(csp):
        c = CubicSpline2D.from_foreign(csp)
        cx, cy = np.asarray(c.coef_x), np.asarray(c.coef_y)
        return (len(c.s), float(c.s[0]), float(c.s[-1]), float(cx[0][0]), float(cx[-1][3]), float(cy[0][1]), float(cy[-1][1]))

    # -- one planning step --------------------------------------------------------------------
    def make_step_struct(self, step, c_begin=None, c_end=None, shard_index=0, shard_count=1):
        v = np.ascontiguousarray(step.v_list, dtype=np.float64)
        t = np.ascontiguousarray(step.t_list, dtype=np.float64)
        d = np.ascontiguousarray(step.d_list, dtype=np.float64)
        # samples per horizon, len(np.arange(0, tT, dt)) (frenet_functions.py:267): remembered per (horizons, dt)
        nkey = (t.tobytes(), step.dt)
        cache = self.__dict__.setdefault("_nsamp_cache", {}) if isinstance(self, ScoringContext) else {}
        n = cache.get(nkey)
        if n is None:
            n = np.ascontiguousarray([len(np.arange(0.0, tT, step.dt)) for tT in t], dtype=np.int32)
            if len(cache) < 64:
                cache[nkey] = n
        s = _abi.Step()
        s.c_s, s.c_s_d, s.c_s_dd = step.c_s, step.c_s_d, step.c_s_dd
        s.c_d, s.c_d_d, s.c_d_dd = step.c_d, step.c_d_d, step.c_d_dd
        s.v_list, s.nv = _dptr(v), len(v)
        s.t_list, s.nt = _dptr(t), len(t)
        s.n_samples = _iptr(n)
        s.d_list, s.nd = _dptr(d), len(d)
        s.dt, s.v_thr = step.dt, step.v_thr
        s.velocity_cost_mode = int(step.velocity_cost_mode)
        s.velocity_target = float(step.velocity_target)
        s.cost_factor = float(step.cost_factor)
        keep = [v, t, d, n]
        if step.ext_level is not None:
            e = np.ascontiguousarray(step.ext_level, dtype=np.uint8)
            s.ext_level = e.ctypes.data
            keep.append(e)
        if step.bd_harm is not None:
            b = np.ascontiguousarray(step.bd_harm, dtype=np.float64)
            s.bd_harm = _dptr(b)
            keep.append(b)
        if getattr(step, "cost_factor_list", None) is not None:
            fl = np.ascontiguousarray(step.cost_factor_list, dtype=np.float64)
            s.cost_factor_list = _dptr(fl)
            keep.append(fl)
        if getattr(step, "goal", None) is not None:
            g = make_goal(step.goal, keep)
            s.goal = C.pointer(g)
            keep.append(g)
        dense = len(v) * len(t) * len(d)
        s.c_begin = 0 if c_begin is None else int(c_begin)
        s.c_end = dense if c_end is None else int(c_end)
        s.shard_index, s.shard_count = int(shard_index), int(shard_count)
        return s, keep, n

    def allocate_outputs(self, Cs, Tmax, O, traj=True, prob=True, red=True, cost=True):
        """Device tensors of one call, candidate-minor with the candidate stride padded to a multiple of 32
        (etp_out in include/etp_b200.h).  The returned tensors are views of the first ``Cs`` columns."""
        torch = self.torch
        kw = dict(device=self.device, dtype=self.dtype)
        Cp = max(32, (Cs + 31) // 32 * 32)
        bufs = dict(
            traj=torch.empty((_abi.NUM_FIELDS, Tmax, Cp), **kw)[..., :Cs] if traj else None,
            prob=torch.empty((Tmax - 1, O, Cp), **kw)[..., :Cs] if prob else None,
            red=torch.empty((_abi.NUM_RED, O, Cp), **kw)[..., :Cs] if red else None,
            cost=torch.empty((_abi.NUM_COST, Cp), **kw)[..., :Cs] if cost else None,
            level=torch.empty((Cs,), device=self.device, dtype=torch.uint8),
            reason=torch.empty((Cs,), device=self.device, dtype=torch.uint8),
            bd_harm=torch.empty((Cs,), device=self.device, dtype=torch.float64),
        )
        for t in bufs.values():
            if t is not None:
                t.record_stream(self.stream)
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        bufs["c_stride"] = Cp
        return bufs

    def launch(self, s, bufs=None):
        """Asynchronous etp_plan_step; ``bufs`` is a dict from allocate_outputs or None (argmin only)."""
        if bufs is None:
            self._check(self.lib.etp_plan_step(self.h, C.byref(s), None))
            return
        o = _abi.Out()
        o.precision = self.precision
        o.c_stride = int(bufs.get("c_stride", 0))
        for k in ("traj", "prob", "red", "cost", "level", "reason", "harm", "bd_harm"):
            t = bufs.get(k)
            setattr(o, k, None if t is None else C.c_void_p(t.data_ptr()))
        self._check(self.lib.etp_plan_step(self.h, C.byref(s), C.byref(o)))

    def best(self):
        b = _abi.Best()
        self._check(self.lib.etp_get_best(self.h, C.byref(b)))
        return dict(index=b.index, level=b.level, cost=b.cost, n_scored=b.n_scored, list_index=b.list_index,
                    key_hi=b.key_hi, key_lo=b.key_lo)

    def best_and_trajectory(self):
        """(best dict, fields [13][Tmax] fp64, T) of the last step with one launch, one copy, one synchronisation."""
        b = _abi.Best()
        f = np.zeros((_abi.NUM_FIELDS, self._last_tmax), dtype=np.float64)
        n = C.c_int32(0)
        self._check(self.lib.etp_get_best_trajectory(self.h, C.byref(b), _dptr(f), C.byref(n)))
        best = dict(index=b.index, level=b.level, cost=b.cost, n_scored=b.n_scored, list_index=b.list_index,
                    key_hi=b.key_hi, key_lo=b.key_lo)
        return best, f, int(n.value)

    def best_raw(self):
        """Best record without raising on 'no candidate' (used by the sharded arg-min)."""
        b = _abi.Best()
        rc = self.lib.etp_get_best(self.h, C.byref(b))
        if rc not in (_abi.OK, _abi.ERR_NO_CANDIDATE):
            self._check(rc)
        return b

    def enable_timing(self, enable=True):
        self._check(self.lib.etp_enable_timing(self.h, int(enable)))

    def timing(self):
        """(profile stage ms, fused scoring kernel ms) of the last step, from CUDA events."""
        a, b = C.c_float(0), C.c_float(0)
        self._check(self.lib.etp_get_timing(self.h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def stage_timing(self):
        """(profile stage, fused scoring kernel, selection) milliseconds of the last step, from CUDA events."""
        v = (C.c_float * 3)()
        self._check(self.lib.etp_get_stage_timing(self.h, v, 3))
        return float(v[0]), float(v[1]), float(v[2])

    def debug_counters(self):
        v = (C.c_uint64 * 8)()
        self._check(self.lib.etp_debug_counters(self.h, v, 8))
        import struct

        return {"fp64_fallback_evals": int(v[0]), "rescored": int(v[1]), "bound_violations": int(v[2]),
                "worst_error_over_bound": struct.unpack("<f", struct.pack("<I", int(v[3]) & 0xffffffff))[0],
                "h2d_bytes": int(v[4]), "d2h_bytes": int(v[5]), "launches": int(v[6]), "graph_launches": int(v[7])}

    def best_device_tensor(self):
        """uint8 view (torch) of the device-resident etp_best record, for collectives."""
        p = C.c_void_p()
        self._check(self.lib.etp_best_device_ptr(self.h, C.byref(p)))
        n = C.sizeof(_abi.Best)

        class _Raw:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (p.value, False), "version": 2}

        return self.torch.as_tensor(_Raw(), device=self.device)

    def trajectory(self, index):
        f = np.zeros((_abi.NUM_FIELDS, self._last_tmax), dtype=np.float64)
        n = C.c_int32(0)
        self._check(self.lib.etp_get_trajectory(self.h, int(index), _dptr(f), C.byref(n)))
        return f, int(n.value)

    # -- scenario sweep ---------------------------------------------------------------------------
    def pack_scenarios(self, steps):
        """Host-side packing of independent planning steps for ``plan_batch`` (structs + the arrays they point to).
        All steps must share params / vehicle / mode (the context's parameters).

        A sweep over one planning grid (equal grid sizes, equally long predictions: the usual case) is packed array-wise --
        one stacked copy per input field for the WHOLE batch, scenario structs pointing into it; anything else goes step
        by step."""
        fast = self._pack_scenarios_uniform(steps)
        if fast is not None:
            return fast
        n = len(steps)
        arr = (_abi.Scenario * n)()
        keep = []
        for i, step in enumerate(steps):
            s, k, _ = self.make_step_struct(step)
            o = _abi.Obstacles()
            if step.predictions is None:
                o.n_obstacles = -1
                pk = None
            else:
                pk = pack_predictions(step.predictions, step.obstacle_types)
                o.n_obstacles, o.max_pred = pk["O"], pk["Tp"]
                o.pred_len = _iptr(pk["pred_len"])
                o.pos, o.cov, o.yaw, o.vel = _dptr(pk["pos"]), _dptr(pk["cov"]), _dptr(pk["yaw"]), _dptr(pk["vel"])
                o.length, o.width, o.mass = _dptr(pk["length"]), _dptr(pk["width"]), _dptr(pk["mass"])
                o.is_protected, o.responsibility = _iptr(pk["prot"]), _iptr(pk["resp"])
            csp = CubicSpline2D.from_foreign(step.csp)
            kn = np.ascontiguousarray(csp.s, dtype=np.float64)
            cx = np.ascontiguousarray(csp.coef_x, dtype=np.float64)
            cy = np.ascontiguousarray(csp.coef_y, dtype=np.float64)
            arr[i].step, arr[i].obstacles = C.pointer(s), C.pointer(o)
            arr[i].knots, arr[i].coef_x, arr[i].coef_y, arr[i].n_seg = _dptr(kn), _dptr(cx), _dptr(cy), len(kn) - 1
            keep.append((s, k, o, pk, kn, cx, cy))
        return arr, keep

    @staticmethod
    def _pack_scenarios_uniform(steps):
        n = len(steps)
        if n == 0:
            return None
        s0 = steps[0]
        nv, nt, nd = len(s0.v_list), len(s0.t_list), len(s0.d_list)
        preds, flat = [], []
        for st in steps:
            p = st.predictions
            if (p is None or not p or len(st.v_list) != nv or len(st.t_list) != nt or len(st.d_list) != nd or st.ext_level is not None
                    or st.bd_harm is not None or st.cost_factor_list is not None or st.goal is not None):
                return None
            vals = list(p.values())
            preds.append((p, vals))
            flat += vals
        Tp = len(flat[0]["pos_list"])
        try:
            pos = np.array([p["pos_list"] for p in flat], dtype=np.float64)
            cov = np.array([p["cov_list"] for p in flat], dtype=np.float64)
            yaw = np.array([p["orientation_list"] for p in flat], dtype=np.float64)
            vel = np.array([p["v_list"] for p in flat], dtype=np.float64)
        except ValueError:  # ragged prediction lengths
            return None
        No = len(flat)
        if pos.shape != (No, Tp, 2) or cov.shape != (No, Tp, 2, 2) or yaw.shape != (No, Tp) or vel.shape != (No, Tp):
            return None
        length = np.array([p["shape"]["length"] for p in flat], dtype=np.float64)
        width = np.array([p["shape"]["width"] for p in flat], dtype=np.float64)
        resp = np.array([p.get("responsibility", 0) for p in flat], dtype=np.int32)
        # class of every obstacle -> (protected, mass or None for the area formula), looked up once per class name
        classes = {}
        prot_l, mass_l = [], []
        for st, (p, _) in zip(steps, preds):
            ot = st.obstacle_types
            for oid in p:
                t = ot[oid]
                c = classes.get(t)
                if c is None:
                    name = getattr(t, "name", t)
                    if name not in PROTECTED_TYPES and name not in UNPROTECTED_TYPES:
                        raise ValueError(f"obstacle type {name!r} has no protection class (harm_estimation.py:52-56)")
                    by_area = name in ("CAR", "PRIORITY_VEHICLE", "PARKED_VEHICLE", "TAXI")
                    c = classes[t] = (1 if name in PROTECTED_TYPES else 0, np.nan if by_area else obstacle_mass(name, 0.0))
                prot_l.append(c[0])
                mass_l.append(c[1])
        prot = np.array(prot_l, dtype=np.int32)
        mass = np.array(mass_l, dtype=np.float64)
        car = np.isnan(mass)
        mass[car] = -1333.5 + 526.9 * np.power((length * width)[car], 0.8)  # properties.py:37, the same ufunc as obstacle_mass
        plen = np.full(No, Tp, dtype=np.int32)
        v = np.array([st.v_list for st in steps], dtype=np.float64)
        t = np.array([st.t_list for st in steps], dtype=np.float64)
        d = np.array([st.d_list for st in steps], dtype=np.float64)
        # samples per horizon, len(np.arange(0, tT, dt)) (frenet_functions.py:267): computed once per distinct (tT, dt)
        nsamp = {}
        ns = np.empty((n, nt), dtype=np.int32)
        for i, st in enumerate(steps):
            for k, tT in enumerate(st.t_list):
                key = (float(tT), st.dt)
                c = nsamp.get(key)
                if c is None:
                    c = nsamp[key] = len(np.arange(0.0, tT, st.dt))
                ns[i, k] = c
        # the struct arrays are filled column by column through numpy views of the ctypes layouts
        counts = np.array([len(vals) for _, vals in preds], dtype=np.int64)
        o0 = np.concatenate([[0], np.cumsum(counts)[:-1]])
        idx = np.arange(n, dtype=np.int64)
        st_arr = np.zeros(n, dtype=_struct_dtype(_abi.Step))
        for name in ("c_s", "c_s_d", "c_s_dd", "c_d", "c_d_d", "c_d_dd", "dt", "v_thr", "velocity_target", "cost_factor"):
            st_arr[name] = [getattr(st, name) for st in steps]
        st_arr["velocity_cost_mode"] = [int(st.velocity_cost_mode) for st in steps]
        st_arr["nv"], st_arr["nt"], st_arr["nd"], st_arr["c_end"], st_arr["shard_count"] = nv, nt, nd, nv * nt * nd, 1
        st_arr["v_list"] = v.ctypes.data + 8 * nv * idx
        st_arr["t_list"] = t.ctypes.data + 8 * nt * idx
        st_arr["d_list"] = d.ctypes.data + 8 * nd * idx
        st_arr["n_samples"] = ns.ctypes.data + 4 * nt * idx
        ob_arr = np.zeros(n, dtype=_struct_dtype(_abi.Obstacles))
        ob_arr["n_obstacles"], ob_arr["max_pred"] = counts, Tp
        for name, a, size in (("pred_len", plen, 4), ("pos", pos, 16 * Tp), ("cov", cov, 32 * Tp), ("yaw", yaw, 8 * Tp), ("vel", vel, 8 * Tp),
                              ("length", length, 8), ("width", width, 8), ("mass", mass, 8), ("is_protected", prot, 4),
                              ("responsibility", resp, 4)):
            ob_arr[name] = a.ctypes.data + size * o0
        # splines: a table is shared by consecutive scenarios that bring the same spline object or equal tables
        tables, knp, cxp, cyp, nsg = [], [], [], [], []
        f64 = lambda a: a if (type(a) is np.ndarray and a.dtype == np.float64 and a.flags.c_contiguous) \
            else np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        for st in steps:
            if not tables or tables[-1][0] is not st.csp:
                csp = CubicSpline2D.from_foreign(st.csp)
                kn, cx, cy = f64(csp.s), f64(csp.coef_x), f64(csp.coef_y)
                sig = (kn.tobytes(), cx.tobytes(), cy.tobytes())
                if tables and sig == tables[-1][4]:
                    tables[-1] = (st.csp,) + tables[-1][1:]
                else:
                    tables.append((st.csp, kn, cx, cy, sig, (kn.ctypes.data, cx.ctypes.data, cy.ctypes.data, len(kn) - 1)))
            ptr = tables[-1][5]
            knp.append(ptr[0]); cxp.append(ptr[1]); cyp.append(ptr[2]); nsg.append(ptr[3])
        sc_arr = np.zeros(n, dtype=_struct_dtype(_abi.Scenario))
        sc_arr["step"] = st_arr.ctypes.data + st_arr.itemsize * idx
        sc_arr["obstacles"] = ob_arr.ctypes.data + ob_arr.itemsize * idx
        sc_arr["knots"], sc_arr["coef_x"], sc_arr["coef_y"], sc_arr["n_seg"] = knp, cxp, cyp, nsg
        arr = (_abi.Scenario * n).from_buffer(sc_arr)
        keep = [sc_arr, st_arr, ob_arr, pos, cov, yaw, vel, length, width, mass, prot, resp, plen, v, t, d, ns, tables,
                [tb[0] for tb in tables]]
        return arr, keep

    def plan_batch(self, steps=None, packed=None):
        """One planning step per scenario, arg-min only (etp_plan_batch): list of best dicts, ``None`` where every
        candidate of the scenario was skipped."""
        if packed is None:
            if not steps:
                return []
            self.set_params(steps[0].params, steps[0].vehicle, steps[0].mode)
            packed = self.pack_scenarios(steps)
        arr, keep = packed
        n = len(arr)
        bests = (_abi.Best * n)()
        self._check(self.lib.etp_plan_batch(self.h, n, arr, self.precision, bests))
        out = []
        for b in bests:
            out.append(None if b.index < 0 else dict(index=b.index, level=b.level, cost=b.cost, n_scored=b.n_scored,
                                                     list_index=b.list_index, key_hi=b.key_hi, key_lo=b.key_lo))
        return out

    def plan_batch_pipelined(self, steps, chunk=512):
        """``plan_batch`` over ``steps`` in chunks of ``chunk`` scenarios with the host work overlapped: the prediction dicts
        of chunk k + 1 are packed on the calling thread while the library call of chunk k (ctypes releases the GIL for its
        duration) runs on a worker thread.  Same results, in order."""
        if not steps:
            return []
        from concurrent.futures import ThreadPoolExecutor

        self.set_params(steps[0].params, steps[0].vehicle, steps[0].mode)
        out, pending = [], None
        with ThreadPoolExecutor(max_workers=1) as pool:
            for i in range(0, len(steps), chunk):
                packed = self.pack_scenarios(steps[i:i + chunk])
                if pending is not None:
                    out += pending.result()
                pending = pool.submit(self.plan_batch, None, packed)
            out += pending.result()
        return out

    @staticmethod
    def local_candidates(s):
        """Rows of the output tensors for a step struct: (profiles of this shard) * nd."""
        n_prof = (s.c_end - s.c_begin) // s.nd
        cnt = max(1, s.shard_count)
        idx = s.shard_index if s.shard_count > 0 else 0
        n_loc = (n_prof - idx + cnt - 1) // cnt if idx < n_prof else 0
        return n_loc * s.nd

    def local_to_dense(self, s):
        """Dense candidate index of every local output row."""
        cnt = max(1, s.shard_count)
        idx = s.shard_index if s.shard_count > 0 else 0
        rows = np.arange(self.local_candidates(s))
        k, i = rows // s.nd, rows % s.nd
        return (s.c_begin // s.nd + idx + k * cnt) * s.nd + i

    def plan_winner(self, step: PlanningStep):
        """Arg-min only planning step returning ``(best, trajectory dict)``: the 11 arrays FrenetPlanner stores in
        ``self._trajectory`` (frenet_planner.py:564-576), one read-back."""
        self.configure(step)
        s, keep, n = self.make_step_struct(step)
        self.launch(s, None)
        self._last_tmax = int(n.max())
        self.last_dt = step.dt
        best, f, T = self.best_and_trajectory()
        g = {name: f[i, :T] for i, name in enumerate(_abi.FIELD_NAMES)}
        traj = {"s_loc_m": g["s"], "d_loc_m": g["d"], "d_d_loc_mps": g["d_d"], "d_dd_loc_mps2": g["d_dd"], "x_m": g["x"],
                "y_m": g["y"], "psi_rad": g["yaw"], "kappa_radpm": g["curv"], "v_mps": g["s_d"], "ax_mps2": g["s_dd"],
                "time_s": np.arange(T) * step.dt}
        return best, traj

    def plan_cycle(self, step: PlanningStep, raw=None):
        """One closed-loop cycle through ``etp_plan_cycle``: this cycle's predictions (``raw``: the record of
        ``set_raw_predictions(..., upload=False)``; else the enriched dict of ``step.predictions``), the step, and back the
        winner with its trajectory -- one library call, one upload, a replayed CUDA graph of the cycle's launches, one
        read-back.  Returns ``(best, trajectory dict)`` like ``plan_winner``."""
        self.ensure_params(step.params, step.vehicle, step.mode)
        cur = getattr(self, "_spline_key", None)
        if cur is None or cur[0] is not step.csp or cur[1] != self._spline_print(step.csp):
            self.set_spline(CubicSpline2D.from_foreign(step.csp))
            self._spline_key = (step.csp, self._spline_print(step.csp))
        self.ensure_road_boundary(getattr(step, "road_boundary", None))
        if raw is not None and step.ext_level is None and step.bd_harm is None and getattr(step, "cost_factor_list", None) is None:
            return self._plan_cycle_resident(step, raw)
        s, keep, n = self.make_step_struct(step)
        cy = _abi.Cycle()
        cy.step = C.pointer(s)
        if raw is not None:
            cy.raw = C.pointer(raw)
        else:
            o = _abi.Obstacles()
            if step.predictions is None:
                # `predictions is None` (no risk assessment at all) is not a cycle of the closed loop: the regular path
                self.set_predictions(None)
                self.launch(s, None)
                self._last_tmax, self.last_dt = int(n.max()), step.dt
                best, f, T = self.best_and_trajectory()
                return best, self._trajectory_dict(f, T, step.dt)
            pk = pack_predictions(step.predictions, step.obstacle_types)
            o.n_obstacles, o.max_pred = pk["O"], pk["Tp"]
            o.pred_len = _iptr(pk["pred_len"])
            o.pos, o.cov, o.yaw, o.vel = _dptr(pk["pos"]), _dptr(pk["cov"]), _dptr(pk["yaw"]), _dptr(pk["vel"])
            o.length, o.width, o.mass = _dptr(pk["length"]), _dptr(pk["width"]), _dptr(pk["mass"])
            o.is_protected, o.responsibility = _iptr(pk["prot"]), _iptr(pk["resp"])
            cy.obstacles = C.pointer(o)
            self.obstacle_ids, self.O, self.has_pred = pk["ids"], pk["O"], True
            keep.append((o, pk))
        Tmax = int(n.max())
        self._last_tmax, self.last_dt = Tmax, step.dt
        b = _abi.Best()
        f = np.zeros((_abi.NUM_FIELDS, Tmax), dtype=np.float64)
        T = C.c_int32(0)
        self._check(self.lib.etp_plan_cycle(self.h, C.byref(cy), self.precision, C.byref(b), _dptr(f), C.byref(T)))
        best = dict(index=b.index, level=b.level, cost=b.cost, n_scored=b.n_scored, list_index=b.list_index,
                    key_hi=b.key_hi, key_lo=b.key_lo)
        return best, self._trajectory_dict(f, int(T.value), step.dt)

    def _plan_cycle_resident(self, step, raw):
        """``plan_cycle`` of the closed loop proper (raw predictions, nothing pre-evaluated on the host): the records handed to
        the library live as long as the grid keeps its shape -- the sample lists are copied into their arrays, the six state
        scalars are set, nothing is allocated but the result rows."""
        nv, nt, nd = len(step.v_list), len(step.t_list), len(step.d_list)
        cs = self.__dict__.get("_cycle_state")
        if cs is None or cs["shape"] != (nv, nt, nd):
            v, t, d = np.empty(nv), np.empty(nt), np.empty(nd)
            n = np.empty(nt, dtype=np.int32)
            st = _abi.Step()
            st.v_list, st.nv, st.t_list, st.nt, st.d_list, st.nd, st.n_samples = _dptr(v), nv, _dptr(t), nt, _dptr(d), nd, _iptr(n)
            st.c_begin, st.c_end, st.shard_index, st.shard_count = 0, nv * nt * nd, 0, 1
            cy = _abi.Cycle()
            cy.step = C.pointer(st)
            cs = self._cycle_state = dict(shape=(nv, nt, nd), v=v, t=t, d=d, n=n, st=st, cy=cy, best=_abi.Best(), T=C.c_int32(0),
                                          tkey=None, Tmax=0)
        cs["v"][:] = step.v_list
        cs["t"][:] = step.t_list
        cs["d"][:] = step.d_list
        tkey = (cs["t"].tobytes(), step.dt)
        if tkey != cs["tkey"]:  # samples per horizon, len(np.arange(0, tT, dt)) (frenet_functions.py:267)
            cs["n"][:] = [len(np.arange(0.0, tT, step.dt)) for tT in cs["t"]]
            cs["tkey"], cs["Tmax"] = tkey, int(cs["n"].max())
        st, cy, b, T, Tmax = cs["st"], cs["cy"], cs["best"], cs["T"], cs["Tmax"]
        st.c_s, st.c_s_d, st.c_s_dd, st.c_d, st.c_d_d, st.c_d_dd = step.c_s, step.c_s_d, step.c_s_dd, step.c_d, step.c_d_d, step.c_d_dd
        st.dt, st.v_thr = step.dt, step.v_thr
        st.velocity_cost_mode, st.velocity_target, st.cost_factor = int(step.velocity_cost_mode), float(step.velocity_target), float(step.cost_factor)
        if cs.get("raw") is not raw:
            cy.raw = C.pointer(raw)
            cs["raw"] = raw
        goal = getattr(step, "goal", None)
        if goal is not None:  # ego state and time step of this cycle: a new record per cycle, the geometry arrays with it
            keep = []
            g = make_goal(goal, keep)
            st.goal = C.pointer(g)
            cs["goal"] = (g, keep)
        elif cs.get("goal") is not None:
            st.goal = None
            cs["goal"] = None
        self._last_tmax, self.last_dt = Tmax, step.dt
        f = np.empty((_abi.NUM_FIELDS, Tmax), dtype=np.float64)
        self._check(self.lib.etp_plan_cycle(self.h, C.byref(cy), self.precision, C.byref(b), _dptr(f), C.byref(T)))
        best = dict(index=b.index, level=b.level, cost=b.cost, n_scored=b.n_scored, list_index=b.list_index,
                    key_hi=b.key_hi, key_lo=b.key_lo)
        return best, self._trajectory_dict(f, T.value, step.dt)

    @staticmethod
    def _trajectory_dict(f, T, dt):
        """The 11 arrays FrenetPlanner stores in ``self._trajectory`` (frenet_planner.py:564-576; ``v_mps`` is ``s_d``, Q7)."""
        g = {name: f[i, :T] for i, name in enumerate(_abi.FIELD_NAMES)}
        return {"s_loc_m": g["s"], "d_loc_m": g["d"], "d_d_loc_mps": g["d_d"], "d_dd_loc_mps2": g["d_dd"], "x_m": g["x"],
                "y_m": g["y"], "psi_rad": g["yaw"], "kappa_radpm": g["curv"], "v_mps": g["s_d"], "ax_mps2": g["s_dd"],
                "time_s": np.arange(T) * dt}

    def plan(self, step: PlanningStep, materialize=True, c_begin=None, c_end=None, bufs=None,
             configure=True, shard_index=0, shard_count=1) -> ScoreResult:
        """calc_frenet_trajectories + sort_frenet_trajectories + selection of one step
        (frenet_planner.py:326-340, 433-457, 555-558)."""
        if configure:
            self.configure(step)
        s, keep, n = self.make_step_struct(step, c_begin, c_end, shard_index, shard_count)
        Tmax = int(n.max())
        Cs = self.local_candidates(s)
        if materialize and bufs is None:
            bufs = self.allocate_outputs(Cs, Tmax, self.O)
        self.launch(s, bufs if materialize else None)
        self._last_tmax = Tmax
        self.last_dt = step.dt
        best = self.best()
        res = ScoreResult(n_dense=step.n_dense, c_begin=s.c_begin, c_end=s.c_end, Tmax=Tmax, n_samples=n,
                          obstacle_ids=list(self.obstacle_ids), best=best, _ctx=self)
        if materialize:
            for k in ("traj", "prob", "red", "cost", "level", "reason", "bd_harm"):
                setattr(res, k, bufs.get(k))
        return res
