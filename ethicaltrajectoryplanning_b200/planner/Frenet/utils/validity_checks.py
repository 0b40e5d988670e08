"""Drop-in for ``planner/Frenet/utils/validity_checks.py``: ``VALIDITY_LEVELS`` (:22-28), ``check_validity`` (:31-122).

The kinematic tests and the maximum-risk test run on the device.  The collision-with-prediction sweep (:91-103) and
the road-boundary test (:110-115) are pycrcc geometry the reference does not own; their per-trajectory results enter
as host inputs (last row) through ``set_host_checks``.
"""
from __future__ import annotations

import numpy as np

from .... import _abi, dropin

VALIDITY_LEVELS = dropin.VALIDITY_LEVELS

_HOST_CHECKS = None


def set_host_checks(fn):
    """Register ``fn(fp_list, ego_state, scenario, vehicle_params, predictions, road_boundary, collision_checker,
    params) -> (ext_level, bd_harm)``: ``ext_level[i]`` is 1 if trajectory i collides with a prediction
    (validity_checks.py:244-274), 2 if it leaves the road without predictions in play (:216-241), else 10;
    ``bd_harm[i]`` is the boundary harm of risk_costs.py:104-123.  ``None`` restores the default (all pass, 0)."""
    global _HOST_CHECKS
    _HOST_CHECKS = fn


class RoadBoundary:
    """Road boundary as plain triangles [n, 3, 2] (the region outside the road), standing where the reference holds the
    pycrcc ShapeGroup of ``boundary.create_road_boundary_obstacle`` (frenet_planner.py:226-236).  Passed as
    ``road_boundary`` it moves the boundary test (level 2, boundary harm) onto the device; ``check`` = "swept" merges
    consecutive ego boxes like ``trajectory_preprocess_obb_sum`` (own hull), "discrete" tests them as they are."""

    def __init__(self, triangles, check="swept"):
        self.triangles = np.ascontiguousarray(triangles, dtype=np.float64).reshape(-1, 3, 2)
        self.check = check

    def as_step_input(self):
        return (self.triangles, self.check)


def device_road_boundary(road_boundary):
    """``(triangles, check)`` for PlanningStep.road_boundary / ScoringContext.set_road_boundary, or None."""
    return road_boundary.as_step_input() if isinstance(road_boundary, RoadBoundary) else None


def host_validity_inputs(fp_list, ego_state, scenario, vehicle_params, predictions, road_boundary, collision_checker,
                         params):
    if _HOST_CHECKS is None:
        return None, None
    ext, bd = _HOST_CHECKS(fp_list, ego_state, scenario, vehicle_params, predictions, road_boundary, collision_checker,
                           params)
    ext = None if ext is None else np.asarray(ext, dtype=np.uint8)
    bd = None if bd is None else np.asarray(bd, dtype=np.float64)
    return ext, bd


def check_validity(ft, ego_state, scenario, vehicle_params, risk_params, mode: str, road_boundary, collision_checker,
                   predictions: dict = None, exec_timer=None, obstacle_types=None):
    """Validity level and reason of one frenet trajectory: ``(level, reason)`` with the reference's level codes and
    reason strings (validity_checks.py:31-122)."""
    if mode not in ("ground_truth", "WaleNet", "risk"):
        raise ValueError("mode must be ground_truth, WaleNet, or risk")
    params = dropin.params_with(modes=risk_params)
    types = dropin.obstacle_type_names(scenario, predictions, obstacle_types) if predictions else {}
    ext, bd = host_validity_inputs([ft], ego_state, scenario, vehicle_params, predictions, road_boundary,
                                   collision_checker, params)
    if bd is None and getattr(ft, "bd_harm", 0):  # validity_checks.py:110-115 reads ft.bd_harm
        bd = np.array([float(ft.bd_harm)])
    r = dropin.score_trajectories([ft], predictions, types, vehicle_params, params, mode=mode, ext_level=ext, bd_harm=bd,
                                  road_boundary=device_road_boundary(road_boundary))
    return int(r.level[0]), _abi.REASON_NAMES[int(r.reason[0])]
