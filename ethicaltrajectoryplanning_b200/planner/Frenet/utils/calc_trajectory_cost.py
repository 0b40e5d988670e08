"""Drop-in for ``planner/Frenet/utils/calc_trajectory_cost.py``: ``calc_trajectory_costs`` (:33-346).

The cost terms of the scored path (risk principles, velocity, distance to the global path, jerks, travelled
distance) come from the fused kernel.  The goal logic behind ``get_cost_factor`` (:349-415) and the target of
``velocity_costs`` (:438-519) is CommonRoad / pycrcc geometry and stays on the host: ``cost_context`` evaluates the
parts that need no geometry and takes the rest from the planning problem object when it provides them.

With a goal area given as plain geometry (``GoalArea``: polygons / circles instead of a pycrcc ShapeGroup) the whole
goal logic runs on the device: ``goal_context`` turns planning problem + goal area + ego state into the
``GoalContext`` that crosses the ABI as ``etp_goal``.
"""
from __future__ import annotations

import numpy as np

from .... import _abi, dropin
from ....configs import RISK_COSTS_KEY
from ....synthetic import GoalContext

COST = {name: i for i, name in enumerate(_abi.COST_NAMES)}


def cost_context(fp_list, ego_state, planning_problem, scenario, dt, goal_area):
    """(velocity_cost_mode, velocity_target, cost_factor) of one planning step.

    * a planning problem that carries ``velocity_cost_mode / velocity_target / cost_factor`` (a pre-evaluated
      context, e.g. from the caller's own CommonRoad code) is used as is; ``cost_factor`` may be per trajectory;
    * otherwise the branch structure of velocity_costs (:438-519) is followed as far as it needs no geometry:
      goal states without position -> 0.0 (mode 3); remaining time <= 0 -> ``30 - mean v`` (mode 1); else
      ``|avg_dist / remaining_time - mean v|`` (mode 0).  Goal-area containment tests need pycrcc and default to
      "not reached", cost factor 1.0.
    """
    if planning_problem is None:
        return 3, 0.0, 1.0
    if hasattr(planning_problem, "velocity_cost_mode"):
        return (int(planning_problem.velocity_cost_mode), float(getattr(planning_problem, "velocity_target", 0.0)),
                getattr(planning_problem, "cost_factor", 1.0))
    goal = planning_problem.goal
    state0 = goal.state_list[0]
    centers = []
    if getattr(goal, "lanelets_of_goal_position", None) is not None:
        for lanelet_id in goal.lanelets_of_goal_position[0]:
            lanelet = scenario.lanelet_network.find_lanelet_by_id(lanelet_id)
            centers.append(lanelet.center_vertices[int(len(lanelet.center_vertices) / 2.0)])
    elif hasattr(state0, "position"):
        if hasattr(state0.position, "center"):
            centers.append(state0.position.center)
    else:
        return 3, 0.0, 1.0
    pos = np.asarray(ego_state.position, dtype=np.float64)
    avg_dist = float(np.mean([np.linalg.norm(np.asarray(c, dtype=np.float64) - pos) for c in centers]))
    remaining = (state0.time_step.end - ego_state.time_step) * dt  # calc_remaining_time_steps with t = 0
    if remaining > 0.0:
        return 0, avg_dist / remaining, 1.0
    return 1, 0.0, 1.0


class GoalArea:
    """Goal area as plain geometry: the union of simple polygons ([n, 2] vertex arrays) and circles ((x, y, r)).
    Stands where the reference holds a pycrcc ShapeGroup (helper_functions.py:226-271)."""

    def __init__(self, polygons=(), circles=()):
        self.polygons = [np.asarray(p, dtype=np.float64).reshape(-1, 2) for p in polygons]
        self.circles = [tuple(float(v) for v in c) for c in circles]


def get_goal_area_shape_group(planning_problem, scenario):
    """helper_functions.py:226-271 without pycrcc: lanelet polygons of ``goal.lanelets_of_goal_position`` or the goal
    states' position shapes (anything with ``vertices``; rectangles by centre / length / width / orientation; circles
    by centre / radius); ``None`` for a survival scenario."""
    goal = planning_problem.goal
    if getattr(goal, "lanelets_of_goal_position", None) is not None:
        return GoalArea([scenario.lanelet_network.find_lanelet_by_id(i).convert_to_polygon().vertices
                         for i in goal.lanelets_of_goal_position[0]])
    if hasattr(goal.state_list[0], "position"):
        polygons, circles = [], []
        for state in goal.state_list:
            shape = state.position
            if hasattr(shape, "radius"):
                circles.append((shape.center[0], shape.center[1], shape.radius))
            elif hasattr(shape, "vertices"):
                polygons.append(shape.vertices)
            else:
                c, s = np.cos(shape.orientation), np.sin(shape.orientation)
                hl, hw = shape.length / 2.0, shape.width / 2.0
                polygons.append([(shape.center[0] + a * hl * c - b * hw * s, shape.center[1] + a * hl * s + b * hw * c)
                                 for a, b in ((1, 1), (-1, 1), (-1, -1), (1, -1))])
        return GoalArea(polygons, circles)
    return None


def goal_context(ego_state, planning_problem, scenario, goal_area):
    """``GoalContext`` of one planning step, or ``None`` when the goal logic cannot run on the device (no planning
    problem, a pre-evaluated context, or a goal area that is not plain geometry)."""
    if planning_problem is None or hasattr(planning_problem, "velocity_cost_mode") or not hasattr(planning_problem, "goal"):
        return None
    if goal_area is not None and not isinstance(goal_area, GoalArea):
        return None
    goal = planning_problem.goal
    state0 = goal.state_list[0]
    centers = lines = points = None
    if getattr(goal, "lanelets_of_goal_position", None) is not None:  # :473-484, :773-786
        centers, lines = [], []
        for lanelet_id in goal.lanelets_of_goal_position[0]:
            lanelet = scenario.lanelet_network.find_lanelet_by_id(lanelet_id)
            centers.append(lanelet.center_vertices[int(len(lanelet.center_vertices) / 2.0)])
            lines.append(np.asarray(lanelet.center_vertices, dtype=np.float64).reshape(-1, 2))
    elif hasattr(state0, "position"):                                  # :485-488, :788-797
        centers = [state0.position.center] if hasattr(state0.position, "center") else []
        if hasattr(state0.position, "center"):
            points = [np.asarray(s.position.center, dtype=np.float64) for s in goal.state_list]
    interval = lambda name: ((getattr(state0, name).start, getattr(state0, name).end) if hasattr(state0, name) else None)  # noqa: E731
    return GoalContext(
        polygons=None if goal_area is None else goal_area.polygons, circles=None if goal_area is None else goal_area.circles,
        velocity=interval("velocity"), orientation=interval("orientation"), time_step=interval("time_step") or (0, -1),
        ego_time_step=int(ego_state.time_step), ego_position=tuple(np.asarray(ego_state.position, dtype=np.float64)[:2]),
        goal_centers=centers, goal_lines=lines, goal_points=points)


def build_cost_dict(traj, c, params, validity_level, predictions):
    """``(cost, cost_dict)`` of calc_trajectory_costs (:321-346) from the cost rows ``c`` of one candidate."""
    weights, modes = params["weights"], params["modes"]
    cost_dict = {}
    if weights["risk_cost"] > 0.0 and predictions is not None:  # :103-148
        traj.risk_dict = {"bayes": c[COST["bayes"]], "equality": c[COST["equality"]], "maximin": c[COST["maximin"]],
                          "responsibility": c[COST["responsibility"]], "ego": c[COST["ego"]],
                          "total": c[COST["risk_cost"]]}
        if len(predictions) == 0:
            # no obstacle: the principles return the integer 0 (risk_costs.py:145-146,173-174,199-200,222-223; the
            # responsibility loop never runs, :250-253) and the log line prints them as such
            for key in ("bayes", "equality", "maximin", "responsibility", "ego"):
                traj.risk_dict[key] = 0
        cost_dict["risk_cost"] = traj.risk_dict["total"]
    else:
        traj.risk_dict = {}
    for key in ("lon_jerk", "lat_jerk", "velocity", "dist_to_global_path", "travelled_dist", "dist_to_goal_pos"):  # :269-305
        if weights[key] > 0.0:
            cost_dict[key] = c[COST[key]]
    if validity_level < 10 and modes["multiple_cost_functions"]:  # :322-327
        curr_weights = {k: (weights[k] if k in RISK_COSTS_KEY else 0) for k in weights}
    else:
        curr_weights = weights
    total = float(c[COST["total"]])
    return total, {"risk_cost_dict": traj.risk_dict, "weights": curr_weights, "factor": float(c[COST["factor"]]),
                   "unweighted_cost": {k: float(v) for k, v in cost_dict.items()}, "total_cost": total}


def calc_trajectory_costs(traj, global_path, planning_problem, ego_state, validity_level, scenario, params: dict,
                          ego_id: int, dt: float, predictions: dict, goal_area, vehicle_params,
                          sensor_radius: float = 30.0, exec_timer=None, mode=None, reach_set=None, obstacle_types=None):
    """Total cost of one frenet trajectory: ``(cost, cost_dict)`` (calc_trajectory_cost.py:33-346)."""
    w = params["weights"]
    if w.get("responsibility", 0.0) > 0.0:
        raise NotImplementedError("reach-set responsibility is outside the scored path")
    for k in ("visible_area", "dist_to_lane_center"):
        if w.get(k, 0.0) > 0.0:
            raise NotImplementedError(f"cost term {k!r} needs lanelet / sensor geometry and stays on the host")
    types = dropin.obstacle_type_names(scenario, predictions, obstacle_types) if predictions else {}
    vel_mode, vel_target, factor = cost_context([traj], ego_state, planning_problem, scenario, dt, goal_area)
    bd = np.array([float(getattr(traj, "bd_harm", 0) or 0)])
    r = dropin.score_trajectories([traj], predictions, types, vehicle_params, params, mode=(mode or "risk"), bd_harm=bd,
                                  velocity_cost_mode=vel_mode, velocity_target=vel_target,
                                  cost_factor=np.broadcast_to(factor, (1,)),
                                  goal=goal_context(ego_state, planning_problem, scenario, goal_area), dt=dt)
    c = r.cost[:, 0].copy()
    own_risk_only = int(r.level[0]) < 10 and bool(params["modes"]["multiple_cost_functions"])
    want_risk_only = validity_level < 10 and bool(params["modes"]["multiple_cost_functions"])
    if own_risk_only != want_risk_only:
        # the caller costs this trajectory under another list level than its own (:322-327): re-weight the device's
        # unweighted terms accordingly
        risk = w["risk_cost"] * c[COST["risk_cost"]] if (w["risk_cost"] > 0.0 and predictions is not None) else 0.0
        rest = 0.0 if want_risk_only else sum(
            w[k] * c[COST[k]] for k in ("lon_jerk", "lat_jerk", "velocity", "dist_to_global_path", "travelled_dist", "dist_to_goal_pos")
            if w[k] > 0.0)
        c[COST["total"]] = c[COST["factor"]] * (risk + rest)
    return build_cost_dict(traj, c, params, validity_level, predictions)
