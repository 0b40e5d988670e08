"""Shared test helpers: golden fixture loading and comparison utilities."""
from __future__ import annotations

import glob
import json
import os

import numpy as np

from ethicaltrajectoryplanning_b200.spline import CubicSpline2D
from ethicaltrajectoryplanning_b200.synthetic import PlanningStep
from ethicaltrajectoryplanning_b200.vehicleparams import VehicleParameters

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "case_*.npz")))


def load_golden(name):
    """Returns (PlanningStep, dict of reference outputs) of a committed fixture."""
    z = np.load(os.path.join(GOLDEN_DIR, f"case_{name}.npz"), allow_pickle=False)
    oids = [int(o) for o in z["obstacle_ids"]]
    types = [str(t) for t in z["obstacle_types"]]
    predictions = {}
    for j, oid in enumerate(oids):
        predictions[oid] = {
            "pos_list": z[f"pos_{j}"].copy(), "cov_list": z[f"cov_{j}"].copy(),
            "orientation_list": z[f"yaw_{j}"].copy(), "v_list": z[f"vel_{j}"].copy(),
            "shape": {"length": float(z[f"shape_{j}"][0]), "width": float(z[f"shape_{j}"][1])},
            "responsibility": int(z[f"resp_{j}"]),
        }
    st = z["start"]
    step = PlanningStep(
        c_s=float(st[0]), c_s_d=float(st[1]), c_s_dd=float(st[2]), c_d=float(st[3]), c_d_d=float(st[4]),
        c_d_dd=float(st[5]), v_list=z["v_list"].copy(), t_list=z["t_list"].copy(), d_list=z["d_list"].copy(),
        dt=float(z["dt"]), v_thr=float(z["v_thr"]),
        csp=CubicSpline2D(knots=z["knots"], coef_x=z["coef_x"], coef_y=z["coef_y"]),
        predictions=predictions, obstacle_types=dict(zip(oids, types)),
        ego_position=z["ego_position"].copy(), ego_orientation=float(z["ego_orientation"]),
        vehicle=VehicleParameters(str(z["vehicle"])), params=json.loads(str(z["params_json"])),
        mode=str(z["mode"]), velocity_cost_mode=int(z["velocity_cost_mode"]),
        velocity_target=float(z["velocity_target"]), cost_factor=float(z["cost_factor"]),
    )
    ref = {k[4:]: z[k] for k in z.files if k.startswith("ref_")}
    return step, ref


def rel_err(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / (np.abs(b) + floor)


def goal_cases():
    """Planning steps with a goal context: (name, PlanningStep).  The expected cost factors and
    velocity costs of tests/golden/goal_cases.npz were written for exactly these steps by the reference's own
    get_cost_factor / velocity_costs (tests/golden/make_golden_goal.py)."""
    from ethicaltrajectoryplanning_b200.synthetic import GoalContext, make_step

    box = lambda x0, x1, y0, y1: [(x0, y0), (x1, y0), (x1, y1), (x0, y1)]  # noqa: E731
    ahead = box(24.0, 60.0, 0.5, 4.2)                    # on the path ahead: fast, well-placed candidates get in
    here = box(0.0, 17.0, -3.0, 4.0)                     # around the ego: most candidates leave it again
    lshape = [(20.0, -1.0), (60.0, -1.0), (60.0, 6.0), (40.0, 6.0), (40.0, 2.0), (20.0, 2.0)]  # concave
    specs = [
        ("reached_in_time", dict(), dict(polygons=[ahead], velocity=(5.0, 30.0), orientation=(-0.5, 0.5), time_step=(10, 60),
                                         ego_time_step=0, goal_centers=[(42.0, 2.3)])),
        ("reached_too_early", dict(), dict(polygons=[ahead], velocity=(5.0, 30.0), time_step=(40, 90), ego_time_step=3,
                                           goal_centers=[(42.0, 2.3)])),
        ("velocity_window_narrow", dict(), dict(polygons=[ahead], velocity=(13.0, 16.0), orientation=(-0.3, 0.12),
                                                time_step=(0, 25), ego_time_step=2, goal_centers=[(42.0, 2.3), (50.0, 3.0)])),
        ("starts_inside_and_leaves", dict(), dict(polygons=[here], velocity=(0.0, 9.0), time_step=(100, 200), ego_time_step=50,
                                                  goal_centers=[(8.0, 0.5)])),
        ("starts_inside_no_velocity", dict(v0=6.0), dict(polygons=[here], time_step=(100, 200), ego_time_step=120,
                                                         goal_centers=[(8.0, 0.5)])),
        ("survival", dict(), dict(polygons=None, circles=None, time_step=(0, 300), ego_time_step=7, goal_centers=None)),
        ("orientation_folded", dict(), dict(polygons=[ahead], orientation=(3.0, 6.5), time_step=(0, 80), ego_time_step=0,
                                            goal_centers=[(42.0, 2.3)])),
        ("orientation_both_folded", dict(), dict(polygons=[ahead], orientation=(6.0, 6.6), time_step=(0, 80), ego_time_step=0,
                                                 goal_centers=[(42.0, 2.3)])),
        ("circle_and_concave", dict(v0=20.0), dict(polygons=[lshape], circles=[(30.0, 4.5, 2.0)], velocity=(2.0, 40.0),
                                                   time_step=(5, 12), ego_time_step=0, goal_centers=[(45.0, 3.0)])),
        ("time_is_up", dict(), dict(polygons=[ahead], velocity=(5.0, 30.0), time_step=(0, 20), ego_time_step=20,
                                    goal_centers=[(42.0, 2.3)])),
        ("no_centres", dict(), dict(polygons=[ahead], time_step=(0, 20), ego_time_step=0, goal_centers=None)),
        # int(ego_time + t / dt): sample 43 is time step 42 (4.3 / 0.1 = 42.99...), no sample is at 43 -> never "in time"
        ("time_step_43_never", dict(t_list=(5.0,)), dict(polygons=[box(24.0, 130.0, -2.0, 9.0)], time_step=(43, 43), ego_time_step=0,
                                                         goal_centers=[(42.0, 2.3)])),
        ("time_step_42_twice", dict(t_list=(5.0,)), dict(polygons=[box(24.0, 130.0, -2.0, 9.0)], time_step=(42, 42), ego_time_step=0,
                                                         goal_centers=[(42.0, 2.3)])),
        ("two_t", dict(t_list=(2.0, 3.0)), dict(polygons=[ahead], velocity=(5.0, 30.0), time_step=(24, 60), ego_time_step=0,
                                                goal_centers=[(42.0, 2.3)])),
    ]
    out = []
    for name, kw, g in specs:
        step = make_step("planning", seed=77, n_obstacles=4, nv=9, nd=11, **kw)
        # goal positions of calc_dist_to_goal_pos: one centre = the goal state's position shape, several = lanelet centre lines
        # (one-vertex lines, like the lanelet network tests/golden/make_golden_goal.py hands the reference)
        c = g.get("goal_centers")
        if c is not None and len(c) == 1:
            g = dict(g, goal_points=[c[0]])
        elif c is not None:
            g = dict(g, goal_lines=[[p] for p in c])
        step.goal = GoalContext(ego_position=tuple(step.ego_position), **g)
        out.append((name, step))
    return out
