"""TEST INFRASTRUCTURE -- not part of the product path (only tests/, smoke() and bench.py's CPU legs import oracle/).

fp64 restatement of the goal logic of the cost function: ``get_cost_factor`` (calc_trajectory_cost.py:349-415),
``velocity_costs`` (:438-519), ``goal_position_reached_and_left_again`` (:520-550), ``reached_goal_state`` (:553-571),
``reached_target_position / _velocity / _yaw / _time`` (:574-660), ``calc_final_diff`` (:663-699),
``calc_remaining_time_steps`` (:702-723), ``check_orientation`` (:881-932), ``is_in_interval``
(helper_functions.py:408-421), on a ``GoalContext`` (ethicaltrajectoryplanning_b200/synthetic.py) instead of CommonRoad
objects.  Pinned by tests/golden/goal_cases.npz, written by the reference's own functions under oracle/ref_shim.py with
``pycrcc.Point.collide`` replaced by the even-odd containment test below (pycrcc is not installable here; its treatment of
points exactly on an edge is the one thing not reproduced).
"""
from __future__ import annotations

import math

import numpy as np


def point_in_goal(goal, x, y):
    """reached_target_position (:574-594): goal_area None -> True; else the point lies in one of the shapes."""
    if not goal.has_area:
        return True
    for poly in goal.polygons or []:
        p = np.asarray(poly, dtype=np.float64)
        inside = False
        j = len(p) - 1
        for i in range(len(p)):
            xi, yi, xj, yj = p[i, 0], p[i, 1], p[j, 0], p[j, 1]
            if ((yi > y) != (yj > y)) and (x < (xj - xi) * (y - yi) / (yj - yi) + xi):
                inside = not inside
            j = i
        if inside:
            return True
    for cx, cy, r in goal.circles or []:
        if (x - cx) ** 2 + (y - cy) ** 2 <= r * r:
            return True
    return False


def is_in_interval(value, interval):
    return interval[0] < value < interval[1]


def check_orientation(value, interval):
    """:881-932: interval ends folded into [-pi, pi]; if exactly one end was folded the test is inverted."""
    n_in_range = 0
    new = []
    for v in interval:
        if v > math.pi:
            while v > math.pi:
                v -= 2 * math.pi
            new.append(v)
        elif v < -math.pi:
            while v < -math.pi:
                v += 2 * math.pi
            new.append(v)
        else:
            n_in_range += 1
            new.append(v)
    new.sort()
    if n_in_range == 1:
        return not is_in_interval(value, new)
    return is_in_interval(value, new)


def reached_target_velocity(goal, vel):
    return goal.velocity is None or is_in_interval(vel, goal.velocity)        # calc_final_diff(...) == 0.0


def reached_target_yaw(goal, yaw):
    return goal.orientation is None or check_orientation(yaw, goal.orientation)


def reached_target_time(goal, t, dt):
    considered = int(goal.ego_time_step + t / dt)                              # :718
    return goal.time_step[0] - considered <= 0.0 and goal.time_step[1] - considered >= 0.0


def get_cost_factor(goal, traj, dt):
    """:349-415."""
    factor = 1.0
    goal_reached = False
    for i in range(len(traj.x)):
        if point_in_goal(goal, traj.x[i], traj.y[i]) and reached_target_velocity(goal, traj.v[i]) \
                and reached_target_yaw(goal, traj.yaw[i]):
            goal_reached = True
            break
    check_good = True
    if goal_reached:
        factor = 0.01
        for i in range(len(traj.x)):
            if reached_target_time(goal, traj.t[i], dt):
                check_good = False
                break
    if check_good is False:
        return 0.0001
    if point_in_goal(goal, goal.ego_position[0], goal.ego_position[1]):       # :520-550
        for i in range(len(traj.x)):
            if not point_in_goal(goal, traj.x[i], traj.y[i]):
                return 10.0
    return factor


def velocity_costs(goal, traj, dt):
    """:438-519."""
    if point_in_goal(goal, traj.x[0], traj.y[0]):
        if goal.velocity is not None:
            return abs((goal.velocity[0] + (goal.velocity[1] - goal.velocity[0]) / 2) - np.mean(traj.v))
        return np.mean(traj.v)
    if goal.goal_centers is None or len(goal.goal_centers) == 0:
        # no centres: `else: return 0.0` (:489-490); an empty list cannot happen in the reference (np.mean([]) is nan)
        return 0.0
    pos = np.asarray(goal.ego_position, dtype=np.float64)
    avg_dist = np.mean([np.linalg.norm(np.asarray(c, dtype=np.float64) - pos) for c in goal.goal_centers])
    remaining_time = (goal.time_step[1] - int(goal.ego_time_step + 0.0 / dt)) * dt
    if remaining_time > 0.0:
        return abs(avg_dist / remaining_time - np.mean(traj.v))
    return 30.0 - np.mean(traj.v)


def dist_to_nearest_point(center_vertices, pos):
    """:835-853: shapely's ``linestring.interpolate(linestring.project(point))`` is the point of the polyline closest to
    ``pos``; restated segment by segment (shapely itself is not installed here: this branch is pinned to the published
    semantics of project / interpolate, not to a run of the reference)."""
    p = np.asarray(center_vertices, dtype=np.float64).reshape(-1, 2)
    pos = np.asarray(pos, dtype=np.float64)
    if len(p) == 1:
        return float(np.linalg.norm(pos - p[0]))
    best = np.inf
    for a, b in zip(p[:-1], p[1:]):
        e = b - a
        l2 = float(e @ e)
        u = 0.0 if l2 == 0.0 else min(1.0, max(0.0, float((pos - a) @ e) / l2))
        best = min(best, float(np.linalg.norm(pos - (a + u * e))))
    return best


def calc_dist_to_goal_pos(goal, x_end, y_end):
    """:760-803: distance of the trajectory's last position to the closest goal position -- centre lines of the goal
    lanelets, else centres of the goal states' shapes, else 0 (survival scenario)."""
    pos = np.array([x_end, y_end], dtype=np.float64)
    lines = getattr(goal, "goal_lines", None) if goal is not None else None
    points = getattr(goal, "goal_points", None) if goal is not None else None
    if lines:
        return min(dist_to_nearest_point(l, pos) for l in lines)
    if points is not None and len(points) > 0:
        return min(float(np.linalg.norm(np.asarray(c, dtype=np.float64) - pos)) for c in points)
    return 0


# ------------------------------------------------------------------------------------------------
# vectorised over candidates of equal length (oracle/vec.py)
# ------------------------------------------------------------------------------------------------
def points_in_goal(goal, x, y):
    if not goal.has_area:
        return np.ones(np.shape(x), dtype=bool)
    out = np.zeros(np.shape(x), dtype=bool)
    for poly in goal.polygons or []:
        p = np.asarray(poly, dtype=np.float64)
        inside = np.zeros(np.shape(x), dtype=bool)
        j = len(p) - 1
        for i in range(len(p)):
            xi, yi, xj, yj = p[i, 0], p[i, 1], p[j, 0], p[j, 1]
            with np.errstate(divide="ignore", invalid="ignore"):
                cross = ((yi > y) != (yj > y)) & (x < (xj - xi) * (y - yi) / (yj - yi) + xi)
            inside ^= cross
            j = i
        out |= inside
    for cx, cy, r in goal.circles or []:
        out |= (x - cx) ** 2 + (y - cy) ** 2 <= r * r
    return out


def factor_and_velocity_cost(goal, x, y, v, yaw, t, dt):
    """x, y, v, yaw: [n, T]; t: [T] -> (factor[n], velocity cost[n])."""
    n, T = x.shape
    inside = points_in_goal(goal, x, y)
    vok = np.ones_like(inside) if goal.velocity is None else (goal.velocity[0] < v) & (v < goal.velocity[1])
    if goal.orientation is None:
        ook = np.ones_like(inside)
    else:
        ook = np.vectorize(lambda a: check_orientation(a, goal.orientation))(yaw)
    reached = (inside & vok & ook).any(axis=1)
    in_time = any(reached_target_time(goal, ti, dt) for ti in t)
    factor = np.where(reached, 0.0001 if in_time else 0.01, 1.0)
    if point_in_goal(goal, goal.ego_position[0], goal.ego_position[1]):
        left = (~inside).any(axis=1)
        factor = np.where(left & ~(reached & in_time), 10.0, factor)
    mv = v.mean(axis=1)
    if goal.velocity is not None:
        inside_cost = np.abs((goal.velocity[0] + (goal.velocity[1] - goal.velocity[0]) / 2) - mv)
    else:
        inside_cost = mv
    if goal.goal_centers is None or len(goal.goal_centers) == 0:
        outside_cost = np.zeros(n)
    else:
        pos = np.asarray(goal.ego_position, dtype=np.float64)
        avg_dist = np.mean([np.linalg.norm(np.asarray(c, dtype=np.float64) - pos) for c in goal.goal_centers])
        remaining_time = (goal.time_step[1] - int(goal.ego_time_step + 0.0 / dt)) * dt
        outside_cost = np.abs(avg_dist / remaining_time - mv) if remaining_time > 0.0 else 30.0 - mv
    return factor, np.where(inside[:, 0], inside_cost, outside_cost)
