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
