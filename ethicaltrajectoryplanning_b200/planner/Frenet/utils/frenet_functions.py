"""Drop-in for ``planner/Frenet/utils/frenet_functions.py``: ``FrenetTrajectory`` (:32-102),
``calc_frenet_trajectories`` (:207-474), ``sort_frenet_trajectories`` (:477-595), ``get_v_list`` (:760-896).

Same signatures, same returned objects; the work is one fused CUDA launch per call (``etp_plan_step`` for lists this
module generated itself, ``etp_score_trajectories`` for arbitrary ``fp_list``s).
"""
from __future__ import annotations

import numpy as np

from .... import _abi, dropin
from ....synthetic import PlanningStep, get_v_list  # noqa: F401  (get_v_list re-exported like the reference)
from ...utils.timers import device_stages, timer_or_noop
from .calc_trajectory_cost import build_cost_dict, cost_context, goal_context
from .validity_checks import VALIDITY_LEVELS, device_road_boundary, host_validity_inputs


class FrenetTrajectory:
    """Trajectory in frenet Coordinates with longitudinal and lateral position and up to 3rd derivative.
    It also includes the global pose and curvature (frenet_functions.py:32-102: same attributes, same order)."""

    def __init__(self, t=None, d=None, d_d=None, d_dd=None, d_ddd=None, s=None, s_d=None, s_dd=None, s_ddd=None,
                 x=None, y=None, yaw=None, v=None, curv=None):
        self.t = t
        self.d = d
        self.d_d = d_d
        self.d_dd = d_dd
        self.d_ddd = d_ddd
        self.s = s
        self.s_d = s_d
        self.s_dd = s_dd
        self.s_ddd = s_ddd
        self.x = x
        self.y = y
        self.yaw = yaw
        self.v = v
        self.curv = curv
        self.valid_level = 0
        self.reason_invalid = None
        self.cost = 0
        self.ego_risk_dict = []
        self.obst_risk_dict = []


class FrenetTrajectoryList(list):
    """``list`` of FrenetTrajectory that remembers the sampling grid it was generated from, so that
    ``sort_frenet_trajectories`` can regenerate the candidates on the device instead of uploading them."""

    generation = None    # dict of calc_frenet_trajectories arguments
    dense_index = None   # dense candidate index (v -> t -> d order incl. skipped slots) of every list entry


def calc_frenet_trajectories(c_s, c_s_d, c_s_dd, c_d, c_d_d, c_d_dd, d_list, t_list, v_list, dt, csp, v_thr=3.0,
                             exec_timer=None):
    """Calculate all possible frenet trajectories from a given starting point (frenet_functions.py:207-474).

    Returns the candidates in generation order v -> t -> d; targets with ``ds <= |dT|`` are left out (:333-334)."""
    ctx = dropin.default_context()
    step = PlanningStep(
        c_s=float(c_s), c_s_d=float(c_s_d), c_s_dd=float(c_s_dd), c_d=float(c_d), c_d_d=float(c_d_d),
        c_d_dd=float(c_d_dd), v_list=np.asarray(v_list, dtype=np.float64), t_list=np.asarray(t_list, dtype=np.float64),
        d_list=np.asarray(d_list, dtype=np.float64), dt=float(dt), v_thr=float(v_thr), csp=csp, predictions=None,
        obstacle_types={}, ego_position=np.zeros(2), ego_orientation=0.0, vehicle=dropin.default_vehicle(),
        params=dropin.params_with(), mode="risk", velocity_cost_mode=3)
    ctx.set_params(step.params, step.vehicle, step.mode)
    ctx.set_spline(csp)
    ctx.set_predictions(None)
    s, keep, n = ctx.make_step_struct(step)
    Tmax, Cs = int(n.max()), ctx.local_candidates(s)
    bufs = ctx.allocate_outputs(Cs, Tmax, 0, prob=False, red=False, cost=False)
    # the reference times the sub-steps of its triple loop under "simulation/calculate trajectories/..." (frenet_functions.py:
    # 255-452); here they are one launch, timed on the device
    with device_stages(ctx, timer_or_noop(exec_timer), "simulation/calculate trajectories/"):
        ctx.launch(s, bufs)
        ctx.stream.synchronize()
    traj = bufs["traj"].detach().cpu().numpy().astype(np.float64)  # [13][Tmax][C]
    level = bufs["level"].detach().cpu().numpy()
    out = FrenetTrajectoryList()
    dense = []
    nt, nd = len(step.t_list), len(step.d_list)
    for c in np.flatnonzero(level != _abi.LEVEL_SKIPPED):
        T = int(n[(c // nd) % nt])
        f = traj[:, :T, c]
        out.append(FrenetTrajectory(t=np.arange(0.0, step.t_list[(c // nd) % nt], step.dt)[:T], d=f[0], d_d=f[1], d_dd=f[2],
                                    d_ddd=f[3], s=f[4], s_d=f[5], s_dd=f[6], s_ddd=f[7], x=f[8], y=f[9], yaw=f[10],
                                    v=f[11], curv=f[12]))
        dense.append(int(c))
    out.generation = dict(c_s=c_s, c_s_d=c_s_d, c_s_dd=c_s_dd, c_d=c_d, c_d_d=c_d_d, c_d_dd=c_d_dd, d_list=step.d_list,
                          t_list=step.t_list, v_list=step.v_list, dt=dt, csp=csp, v_thr=v_thr)
    out.dense_index = np.asarray(dense, dtype=np.int64)
    return out


def _ego_pose(ego_state):
    if ego_state is None:
        return np.zeros(2), 0.0
    return np.asarray(ego_state.position, dtype=np.float64), float(ego_state.orientation)


def sort_frenet_trajectories(ego_state, fp_list, global_path, predictions, mode, params, planning_problem, scenario,
                             vehicle_params, ego_id, dt, sensor_radius, road_boundary, collision_checker, goal_area,
                             exec_timer=None, reach_set=None, obstacle_types=None):
    """Check validity of all frenet trajectories in fp_list and cost the ones of the highest non-empty level
    (frenet_functions.py:477-595).  Returns ``(ft_list_highest_validity, ft_list_invalid, validity_dict)``; like the
    reference the caller sorts the first list by ``fp.cost`` (frenet_planner.py:457)."""
    if mode not in ("ground_truth", "WaleNet", "risk"):
        raise ValueError("mode must be ground_truth, WaleNet, or risk")
    if reach_set is not None or params["weights"].get("responsibility", 0.0) > 0.0:
        raise NotImplementedError("reach-set responsibility (planner/utils/reachable_set.py) is outside the scored path")
    n_list = len(fp_list)
    validity_dict = {key: [] for key in VALIDITY_LEVELS}
    if n_list == 0:
        max([])  # the reference raises ValueError here (frenet_functions.py:562-564)
    types = dropin.obstacle_type_names(scenario, predictions, obstacle_types) if predictions else {}
    ext_level, bd_harm = host_validity_inputs(fp_list, ego_state, scenario, vehicle_params, predictions, road_boundary,
                                              collision_checker, params)
    vel_mode, vel_target, factor = cost_context(fp_list, ego_state, planning_problem, scenario, dt, goal_area)
    goal = goal_context(ego_state, planning_problem, scenario, goal_area)  # goal logic on the device when it can
    generated = (isinstance(fp_list, FrenetTrajectoryList) and fp_list.generation is not None
                 and fp_list.dense_index is not None and len(fp_list.dense_index) == n_list)
    ctx = dropin.default_context()
    if generated:
        g = fp_list.generation
        pos, ori = _ego_pose(ego_state)
        step = PlanningStep(
            c_s=float(g["c_s"]), c_s_d=float(g["c_s_d"]), c_s_dd=float(g["c_s_dd"]), c_d=float(g["c_d"]),
            c_d_d=float(g["c_d_d"]), c_d_dd=float(g["c_d_dd"]), v_list=g["v_list"], t_list=g["t_list"], d_list=g["d_list"],
            dt=float(g["dt"]), v_thr=float(g["v_thr"]), csp=g["csp"], predictions=predictions, obstacle_types=types,
            ego_position=pos, ego_orientation=ori, vehicle=vehicle_params, params=params, mode=mode,
            velocity_cost_mode=vel_mode, velocity_target=vel_target, cost_factor=1.0)
        dense = fp_list.dense_index
        n_dense = step.n_dense
        if ext_level is not None:
            e = np.full(n_dense, 10, dtype=np.uint8)
            e[dense] = ext_level
            step.ext_level = e
        if bd_harm is not None:
            b = np.zeros(n_dense)
            b[dense] = bd_harm
            step.bd_harm = b
        fl = np.ones(n_dense)
        fl[dense] = factor
        step.cost_factor_list = fl
        step.goal = goal
        step.road_boundary = device_road_boundary(road_boundary)
        # "simulation/sort trajectories/check validity/total" and ".../calculate costs/total" (frenet_functions.py:545,
        # calc_trajectory_cost.py:79-344) are one fused launch here
        with device_stages(ctx, timer_or_noop(exec_timer), "simulation/sort trajectories/"):
            res = ctx.plan(step)
        out = res.numpy()
        red = out["red"][:, dense].astype(np.float64)
        cost = out["cost"][:, dense].astype(np.float64)
        level = out["level"][dense].astype(np.int64)
        reason = out["reason"][dense].astype(np.int64)
        bd_dev = out["bd_harm"][dense]
        oids = res.obstacle_ids
    else:
        r = dropin.score_trajectories(list(fp_list), predictions, types, vehicle_params, params, mode=mode,
                                      ext_level=ext_level, bd_harm=bd_harm, velocity_cost_mode=vel_mode,
                                      velocity_target=vel_target, cost_factor=np.broadcast_to(factor, (n_list,)), ctx=ctx,
                                      goal=goal, dt=dt, road_boundary=device_road_boundary(road_boundary))
        red, cost, level, reason, oids, bd_dev = r.red, r.cost, r.level, r.reason, r.obstacle_ids, r.bd_harm
    if device_road_boundary(road_boundary) is not None:
        bd_harm = bd_dev  # the device's own boundary harm (etp_out.bd_harm)
    return fill_trajectories(fp_list, red, cost, level, reason, oids, predictions, bd_harm, params, mode)


def fill_trajectories(fp_list, red, cost, level, reason, oids, predictions, bd_harm, params, mode):
    """Per-candidate result arrays -> the attributes sort_frenet_trajectories leaves on the trajectories
    (frenet_functions.py:529-593) and its return value.  Shared with the columnar log reader (logging.py)."""
    validity_dict = {key: [] for key in VALIDITY_LEVELS}
    for i, fp in enumerate(fp_list):
        if predictions is not None:  # frenet_functions.py:529-542
            T = len(fp.x)
            have = [j for j, oid in enumerate(oids) if min(T - 1, len(predictions[oid]["pos_list"])) > 0]
            fp.ego_risk_dict = {oids[j]: float(red[0, i, j]) for j in have}
            fp.obst_risk_dict = {oids[j]: float(red[1, i, j]) for j in have}
            fp.ego_harm_dict = {oids[j]: float(red[2, i, j]) for j in have}
            fp.obst_harm_dict = {oids[j]: float(red[3, i, j]) for j in have}
            fp.bd_harm = 0 if bd_harm is None else bd_harm[i]
        fp.valid_level = int(level[i])
        fp.reason_invalid = _abi.REASON_NAMES[int(reason[i])]
        validity_dict[fp.valid_level].append(fp)

    validity_level = max([lvl for lvl in VALIDITY_LEVELS if len(validity_dict[lvl])])
    ft_list_highest_validity = validity_dict[validity_level]
    ft_list_invalid = [fp for inv in VALIDITY_LEVELS if inv < validity_level for fp in validity_dict[inv]]
    cost_predictions = None if mode in ("ground_truth", "WaleNet") else predictions
    for i, fp in enumerate(fp_list):
        if fp.valid_level == validity_level:
            fp.cost, fp.cost_dict = build_cost_dict(fp, cost[:, i], params, validity_level, cost_predictions)
    return ft_list_highest_validity, ft_list_invalid, validity_dict
