"""Synthetic planning steps of the shapes named in BASELINE.json / SURVEY.md section 8d.

Everything is drawn from ``numpy.random.default_rng(seed)`` in float64.  The generator
lives in the product package because ``bench.py`` and the drop-in planner need it; the
oracle only ever sees its plain-array output.
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np

from .configs import default_params
from .spline import CubicSpline2D
from .vehicleparams import VehicleParameters

# obstacle classes of the harm model (harm_estimation.py:41-58, properties.py:14-46)
PROTECTED_TYPES = ("CAR", "TRUCK", "BUS", "PRIORITY_VEHICLE", "PARKED_VEHICLE", "TRAIN", "TAXI")
UNPROTECTED_TYPES = ("BICYCLE", "PEDESTRIAN", "MOTORCYCLE", "UNKNOWN")
_FIXED_MASS = {"TRUCK": 25000.0, "BUS": 13000.0, "BICYCLE": 90.0, "PEDESTRIAN": 75.0,
               "TRAIN": 118800.0, "MOTORCYCLE": 250.0}


def obstacle_mass(type_name: str, size: float) -> float:
    """Mass table of risk_assessment/helpers/properties.py:14-46 (size = length*width incl. margins)."""
    if type_name in ("CAR", "PRIORITY_VEHICLE", "PARKED_VEHICLE", "TAXI"):
        return float(-1333.5 + 526.9 * np.power(size, 0.8))  # np.power like the reference (properties.py:37)
    return _FIXED_MASS.get(type_name, 0.0)


_RAMPS = {}


def get_v_list(v_min, v_max, v_cur, v_goal_min=None, v_goal_max=None, n_samples=3, mode="linspace"):
    """End-velocity samples, ``linspace`` mode of frenet_functions.py:760-807.

    The current velocity is appended after the linspace, unsorted (quirk Q8).  The
    ``deterministic`` / ``random`` modes need ``scipy.stats.beta.fit`` and stay on the
    reference host code; they are outside the scored path.
    """
    if n_samples <= 0:
        raise ValueError("Number of samples must be at least 1")
    if (v_goal_min is None or v_goal_max is None) and v_goal_max != v_goal_min:
        raise AttributeError("Both goal velocities must be None or both must be filled")
    if mode != "linspace":
        raise ValueError("only the linspace v-list mode is part of the scored path")
    if v_goal_min is None:
        lo, hi = v_min, v_max
    else:
        lo, hi = max(min(v_goal_min, v_min), 0.001), max(v_goal_max, v_max)
    # np.append(np.linspace(lo, hi, n_samples - 1), v_cur) without the two calls' overhead (a planning cycle pays for it every
    # 0.1 s): numpy's own arithmetic, arange(n) * step + start with the last sample set to the stop value
    n = n_samples - 1
    out = np.empty(n_samples, dtype=np.float64)
    if n == 1:
        out[0] = lo
    elif n > 1:
        lo, hi = float(lo), float(hi)
        step = (hi - lo) / (n - 1)
        ramp = _RAMPS.get(n)
        if ramp is None:
            ramp = _RAMPS[n] = np.arange(0, n, dtype=np.float64)
        if step == 0.0:
            np.multiply(ramp, hi - lo, out=out[:n])
            out[:n] /= n - 1
        else:
            np.multiply(ramp, step, out=out[:n])
        out[:n] += lo
        out[n - 1] = hi
    out[n] = v_cur
    return out


def planner_v_list(v_cur, t_list, vehicle, n_samples, v_goal_min=None, v_goal_max=None):
    """Velocity range rule of FrenetPlanner._step_planner (frenet_planner.py:301-319)."""
    a_max = vehicle.longitudinal.a_max
    max_v = min(v_cur + (a_max / 2.0) * max(t_list), vehicle.longitudinal.v_max)
    min_v = max(0.01, v_cur - a_max * min(t_list))
    return get_v_list(min_v, max_v, v_cur, v_goal_min, v_goal_max, n_samples, "linspace")


def enrich_predictions(predictions, shapes, init_orientation, init_velocity, dt,
                       safety_margin_length=1.0, safety_margin_width=0.5):
    """orientation / velocity / shape enrichment of prediction_helpers.py:75-135 (host side).

    ``shapes[id] = (length, width)`` of the raw obstacle; margins are added here like the
    reference does.  Mutates and returns ``predictions``.
    """
    for oid in list(predictions.keys()):
        pos = np.asarray(predictions[oid]["pos_list"], dtype=np.float64)
        n = len(pos)
        if n == 0:
            del predictions[oid]
            continue
        if n == 1:
            ori = [init_orientation[oid]]
            vel = [init_velocity[oid]]
        else:
            t = [0.0 + i * dt for i in range(n)]
            dx = np.gradient(pos[:, 0], t)
            dy = np.gradient(pos[:, 1], t)
            if all(dxi < 0.0001 for dxi in dx) and all(dyi < 0.0001 for dyi in dy):
                ori = np.full((1, n), init_orientation[oid])[0]
            else:
                ori = np.arctan2(dy, dx)
            vel = np.sqrt(np.power(dx, 2) + np.power(dy, 2))
        predictions[oid]["orientation_list"] = ori
        predictions[oid]["v_list"] = vel
        predictions[oid]["shape"] = {
            "length": shapes[oid][0] + safety_margin_length,
            "width": shapes[oid][1] + safety_margin_width,
        }
    return predictions


def assign_responsibility_by_action_space(ego_position, ego_orientation, predictions, wrap=False):
    """View-cone flag of planner/utils/responsibility.py:57-89 (no angle wrap, quirk Q9; ``wrap`` = the opt-out
    "view_cone_wrap": bearing and orientation are compared modulo 2 pi)."""
    for oid in predictions:
        pos0 = np.asarray(predictions[oid]["pos_list"])[0]
        ang = np.arctan2(pos0[1] - ego_position[1], pos0[0] - ego_position[0])
        if wrap:
            inside = abs(np.remainder(ang - ego_orientation + np.pi, 2 * np.pi) - np.pi) <= np.pi / 4
        else:
            inside = ego_orientation - (np.pi / 4) <= ang <= ego_orientation + (np.pi / 4)
        predictions[oid]["responsibility"] = 0 if inside else 1
    return predictions


@dataclasses.dataclass
class GoalContext:
    """What ``get_cost_factor`` (calc_trajectory_cost.py:349-415) and ``velocity_costs`` (:438-519) read from
    ``planning_problem.goal``, ``goal_area`` and ``ego_state`` (etp_goal in include/etp_b200.h).

    ``polygons``: vertex arrays [n, 2] of the goal area (``lanelet.convert_to_polygon().vertices`` or the goal states'
    position shapes, helper_functions.py:226-271), ``circles``: (x, y, r); both ``None`` = ``goal_area is None``
    (survival scenario).  Intervals are (start, end) or ``None`` when the goal state has no such attribute.
    ``goal_centers``: the points velocity_costs measures the remaining distance to (:473-489), ``None`` if it finds none."""

    polygons: Optional[list] = None
    circles: Optional[list] = None
    velocity: Optional[tuple] = None
    orientation: Optional[tuple] = None
    time_step: tuple = (0, 0)
    ego_time_step: int = 0
    ego_position: tuple = (0.0, 0.0)
    goal_centers: Optional[list] = None
    # goal positions calc_dist_to_goal_pos measures the trajectory's end against (:760-803): the goal lanelets' centre lines
    # ([n, 2] polylines) if the goal is given in lanelets, else the centres of the goal states' position shapes
    goal_lines: Optional[list] = None
    goal_points: Optional[list] = None

    @property
    def has_area(self):
        return self.polygons is not None or self.circles is not None


@dataclasses.dataclass
class PlanningStep:
    """All inputs of one planning step of the scored path (SURVEY.md 8b data contracts)."""

    c_s: float
    c_s_d: float
    c_s_dd: float
    c_d: float
    c_d_d: float
    c_d_dd: float
    v_list: np.ndarray
    t_list: np.ndarray
    d_list: np.ndarray
    dt: float
    v_thr: float
    csp: CubicSpline2D
    predictions: Optional[dict]
    obstacle_types: dict
    ego_position: np.ndarray
    ego_orientation: float
    vehicle: VehicleParameters
    params: dict
    mode: str = "risk"
    # host-only cost context (pycrcc / goal logic stays on the host)
    velocity_cost_mode: int = 0      # 0 |target-mean v|, 1 30-mean v, 2 mean v, 3 zero
    velocity_target: float = 0.0
    cost_factor: float = 1.0
    ext_level: Optional[np.ndarray] = None   # per dense candidate: 1 collision, 2 boundary, else 10
    bd_harm: Optional[np.ndarray] = None     # per dense candidate boundary harm
    cost_factor_list: Optional[np.ndarray] = None  # per dense candidate get_cost_factor result (overrides cost_factor)
    predictions_resident: bool = False  # the context already holds these predictions (ScoringContext.set_raw_predictions)
    goal: Optional[GoalContext] = None  # velocity cost and cost factor derived per candidate on the device (overrides the scalars)
    road_boundary: Optional[tuple] = None  # (triangles [n, 3, 2] outside the road, "discrete" | "swept"): level 2 and boundary harm on the device

    @property
    def n_dense(self):
        return len(self.v_list) * len(self.t_list) * len(self.d_list)

    @property
    def n_samples(self):
        """T per t_list entry: len(np.arange(0, tT, dt)) (frenet_functions.py:267)."""
        return np.array([len(np.arange(0.0, tT, self.dt)) for tT in self.t_list], dtype=np.int64)


def synthetic_path(n_knots=121, spacing=5.0, amp=3.0, wavelength=30.0):
    x = spacing * np.arange(n_knots, dtype=np.float64)
    y = amp * np.sin(x / wavelength)
    return CubicSpline2D(x, y)


GRIDS = {
    # name: (nv, nd, d_max_abs, t_list, n_obstacles)
    "cfg1": (5, 5, 2.5, [2.0], 5),
    "planning": (15, 15, 3.5, [2.0], 8),
    "cfg2": (100, 100, 3.5, [3.0], 8),
    "cfg3": (1000, 1000, 3.5, [5.0], 32),
}


def make_obstacles(csp, rng, n_obstacles, n_pred, dt, v0, t_max, lateral_spread=4.0,
                   zero_cov_index=1, static_index=2):
    """Constant-velocity Gaussian predictions around the path (SURVEY.md 8d 'Obstacles')."""
    predictions, types, shapes, ori0, vel0 = {}, {}, {}, {}, {}
    k = np.arange(n_pred, dtype=np.float64)
    for i in range(n_obstacles):
        oid = 100 + i
        s0 = rng.uniform(10.0, 10.0 + v0 * t_max + 30.0)
        lat = rng.uniform(-lateral_spread, lateral_spread)
        speed = rng.uniform(0.0, 15.0)
        heading_rel = rng.choice([0.0, np.pi]) + rng.normal(0.0, 0.1)
        px, py = csp.calc_position(s0)
        yaw_path = float(csp.calc_yaw(s0))
        p0 = np.array([px - lat * np.sin(yaw_path), py + lat * np.cos(yaw_path)])
        heading = yaw_path + heading_rel
        pedestrian = (i % 4) == 3
        if i == static_index:
            speed = 0.0
        pos = p0[None, :] + (speed * dt * k)[:, None] * np.array([np.cos(heading), np.sin(heading)])[None, :]
        sx = 0.3 + 0.05 * k
        sy = 0.8 * sx
        rho = 0.3
        cov = np.empty((n_pred, 2, 2))
        cov[:, 0, 0] = sx * sx
        cov[:, 1, 1] = sy * sy
        cov[:, 0, 1] = cov[:, 1, 0] = rho * sx * sy
        if i == zero_cov_index:
            cov[:] = 0.0
        if i == static_index:
            cov[:] = 0.0
            cov[:, 0, 0] = cov[:, 1, 1] = 0.02
        predictions[oid] = {"pos_list": pos, "cov_list": cov}
        if pedestrian:
            types[oid] = "PEDESTRIAN"
            shapes[oid] = (0.5, 0.5)
        else:
            types[oid] = "PARKED_VEHICLE" if i == static_index else "CAR"
            shapes[oid] = (4.5, 1.8)
        ori0[oid] = heading
        vel0[oid] = speed
    enrich_predictions(predictions, shapes, ori0, vel0, dt)
    return predictions, types


def road_boundary_triangles(csp, half_width=3.6, outer=12.0, ds=4.0, s_begin=None, s_end=None):
    """Synthetic stand-in of ``boundary.create_road_boundary_obstacle`` (frenet_planner.py:226-236): the region outside
    a road of constant half width around the reference path, two triangles per ``ds`` metres and side, [n, 3, 2]."""
    s0 = float(csp.s[0]) if s_begin is None else s_begin
    s1 = float(csp.s[-1]) if s_end is None else s_end
    ss = np.arange(s0, s1, ds)
    pts = {}
    for side in (1.0, -1.0):
        for name, off in (("in", half_width), ("out", outer)):
            p = []
            for s in ss:
                x, y = csp.calc_position(s)
                yaw = csp.calc_yaw(s)
                p.append((x - side * off * np.sin(yaw), y + side * off * np.cos(yaw)))
            pts[(side, name)] = np.array(p)
    tris = []
    for side in (1.0, -1.0):
        a, b = pts[(side, "in")], pts[(side, "out")]
        for k in range(len(ss) - 1):
            tris.append([a[k], a[k + 1], b[k]])
            tris.append([a[k + 1], b[k + 1], b[k]])
    return np.ascontiguousarray(tris, dtype=np.float64)


def make_step(name="cfg2", seed=1234, v0=12.0, weights="weights_ethical.json", c_d=0.2, c_d_d=0.0,
              c_d_dd=0.0, c_s=5.0, c_s_dd=0.0, lateral_spread=4.0, n_obstacles=None, nv=None, nd=None,
              t_list=None, n_pred=None, vehicle="bmw_320i", obstacle_seed=None) -> PlanningStep:
    """One synthetic planning step of a named grid shape."""
    gnv, gnd, dmax, gt, gno = GRIDS[name]
    nv = gnv if nv is None else nv
    nd = gnd if nd is None else nd
    t_list = np.asarray(gt if t_list is None else t_list, dtype=np.float64)
    n_obstacles = gno if n_obstacles is None else n_obstacles
    veh = VehicleParameters(vehicle)
    dt, v_thr = 0.1, 3.0
    csp = synthetic_path()
    rng = np.random.default_rng(seed if obstacle_seed is None else obstacle_seed)
    T = int(max(len(np.arange(0.0, tT, dt)) for tT in t_list))
    n_pred = T if n_pred is None else n_pred
    if n_obstacles > 0:
        predictions, types = make_obstacles(csp, rng, n_obstacles, n_pred, dt, v0, float(max(t_list)),
                                            lateral_spread)
    else:
        predictions, types = {}, {}
    px, py = csp.calc_position(c_s)
    yaw = float(csp.calc_yaw(c_s))
    ego_pos = np.array([px - c_d * np.sin(yaw), py + c_d * np.cos(yaw)])
    assign_responsibility_by_action_space(ego_pos, yaw, predictions)
    return PlanningStep(
        c_s=c_s, c_s_d=v0, c_s_dd=c_s_dd, c_d=c_d, c_d_d=c_d_d, c_d_dd=c_d_dd,
        v_list=planner_v_list(v0, t_list, veh, nv), t_list=t_list,
        d_list=np.linspace(-dmax, dmax, nd), dt=dt, v_thr=v_thr, csp=csp,
        predictions=predictions, obstacle_types=types, ego_position=ego_pos, ego_orientation=yaw,
        vehicle=veh, params=default_params(weights), mode="risk",
        velocity_cost_mode=0, velocity_target=v0, cost_factor=1.0,
    )
