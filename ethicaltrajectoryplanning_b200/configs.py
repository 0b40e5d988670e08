"""Config surface of the scoring path.

Same file names, keys and loader names as the reference's
``planner/Frenet/configs/load_json.py:8-86`` (``weights*.json``, ``risk.json``,
``harm_parameters.json``, ``planning*.json``).  A user switching over points
``config_dir`` at their own directory of JSON files; without one the built-in
tables below (the values the reference ships) are used.
"""
from __future__ import annotations

import copy
import json
import os

import numpy as np

# order of the weight vector handed to the C-ABI (include/etp_b200.h: ETP_W_*)
WEIGHT_KEYS = (
    "bayes", "equality", "maximin", "responsibility", "ego", "risk_cost",
    "visible_area", "lon_jerk", "lat_jerk", "velocity", "dist_to_global_path",
    "travelled_dist", "dist_to_goal_pos", "dist_to_lane_center",
)
# calc_trajectory_cost.py:30
RISK_COSTS_KEY = ("bayes", "equality", "maximin", "responsibility", "ego", "risk_cost")


def _w(bayes, equality, maximin, responsibility, ego, risk_cost):
    return {
        "bayes": bayes, "equality": equality, "maximin": maximin,
        "responsibility": responsibility, "ego": ego, "risk_cost": risk_cost,
        "visible_area": 0, "lon_jerk": 0.0, "lat_jerk": 0.0, "velocity": 1.0,
        "dist_to_global_path": 10.0, "travelled_dist": 0.0,
        "dist_to_goal_pos": 0.0, "dist_to_lane_center": 0.0,
    }


_BUILTIN = {
    "weights.json": _w(33.3, 33.3, 33.3, 0.5, 0.0, 250),
    "weights_ethical.json": _w(33.3, 33.3, 33.3, 0.0, 0.0, 250),
    "weights_ego.json": _w(0.0, 0.0, 0.0, 0.0, 100.0, 250),
    "weights_standard.json": _w(0.0, 0.0, 0.0, 0.0, 0.0, 1.0),
    "risk.json": {
        "harm_mode": "log_reg",
        "sensor_occlusion_model": False,
        "occlusion_mode": False,
        "ignore_angle": False,
        "sym_angle": True,
        "reduced_angle_areas": True,
        "trajectory_risk": "max",
        "max_acceptable_risk": 1,
        "multiple_cost_functions": False,
        "scale_factor_time": 0.9,
        "crash_angle_accuracy": 10,
        "crash_angle_simplified": True,
        "fast_prob_mahalanobis": False,
        "figures": {"create_figures": False, "number_plotted_trajectories": 3},
        "risk_dashboard": False,
        "collision_report": True,
    },
    "harm_parameters.json": {
        "log_reg": {
            "complete_angle_areas": {
                "const": -4.626, "speed": 0.189,
                "Imp_1": -0.039, "Imp_2": 0.018, "Imp_3": 0.459, "Imp_4": -0.125,
                "Imp_5": -1.413, "Imp_6": -0.116, "Imp_7": -1.782, "Imp_8": -0.434,
                "Imp_9": 0.482, "Imp_10": 0.142, "Imp_11": 0.400,
            },
            "reduced_angle_areas": {
                "const": -4.476, "speed": 0.179,
                "driver_side": 0.250, "right_side": 0.259, "rear": -0.445,
            },
            "ignore_angle": {"const": -4.591, "speed": 0.185},
            "complete_sym_angle_areas": {
                "const": -4.620, "speed": 0.189,
                "Imp_1_11": 0.209, "Imp_2_10": 0.086, "Imp_3_9": 0.470,
                "Imp_4_8": -0.259, "Imp_5_7": -1.590, "Imp_6": -0.118,
            },
            "reduced_sym_angle_areas": {
                "const": -4.457, "speed": 0.177, "side": 0.244, "rear": -0.431,
            },
        },
        "gidas": {"const": -5.820, "speed": 0.292},
        "pedestrian": {"const": 3.164, "speed": 0.288},
        "pedestrian_MAIS2+": {"const": 1.786, "speed": 0.259},
    },
    "planning.json": {
        "evaluation_settings": {
            "timing_enabled": True, "show_visualization": False, "vehicle_type": "bmw_320i",
        },
        "frenet_settings": {
            "mode": "risk",
            "frenet_parameters": {
                "t_list": [2.0], "v_list_generation_mode": "linspace", "n_v_samples": 15,
                "d_list": {"d_max_abs": 3.5, "n": 15}, "dt": 0.1, "v_thr": 3.0,
            },
        },
    },
    "planning_fast.json": {
        "evaluation_settings": {
            "timing_enabled": True, "show_visualization": False, "vehicle_type": "bmw_320i",
        },
        "frenet_settings": {
            "mode": "risk",
            "frenet_parameters": {
                "t_list": [2.0], "v_list_generation_mode": "linspace", "n_v_samples": 5,
                "d_list": {"d_max_abs": 2.5, "n": 5}, "dt": 0.1, "v_thr": 3.0,
            },
        },
    },
}


def _load(filename, config_dir):
    if config_dir is not None:
        with open(os.path.join(config_dir, filename), "r") as f:
            return json.load(f)
    if filename not in _BUILTIN:
        raise FileNotFoundError(f"no built-in config named {filename!r}; pass config_dir")
    return copy.deepcopy(_BUILTIN[filename])


def load_risk_json(filename="risk.json", config_dir=None):
    """Modes of the risk assessment (reference load_json.py:8-25)."""
    return _load(filename, config_dir)


def load_harm_parameter_json(config_dir=None):
    """Harm model coefficients (reference load_json.py:28-41)."""
    return _load("harm_parameters.json", config_dir)


def load_weight_json(filename="weights.json", config_dir=None):
    """Cost weights (reference load_json.py:44-61)."""
    return _load(filename, config_dir)


def load_planning_json(filename="planning.json", config_dir=None):
    """Planner settings; ``d_list`` is expanded with linspace (reference load_json.py:64-83)."""
    data = _load(filename, config_dir)
    fp = data["frenet_settings"]["frenet_parameters"]
    fp["d_list"] = np.linspace(-fp["d_list"]["d_max_abs"], fp["d_list"]["d_max_abs"], fp["d_list"]["n"])
    return data


def default_params(weights="weights_ethical.json", config_dir=None):
    """``params_dict`` of FrenetPlanner.__init__ (reference frenet_planner.py:149-153)."""
    w = load_weight_json(weights, config_dir) if isinstance(weights, str) else dict(weights)
    return {
        "weights": w,
        "modes": load_risk_json(config_dir=config_dir),
        "harm": load_harm_parameter_json(config_dir=config_dir),
    }
