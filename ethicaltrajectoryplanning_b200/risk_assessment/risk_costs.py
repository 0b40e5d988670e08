"""Drop-in for ``risk_assessment/risk_costs.py``: ``calc_risk`` (:16-125) and the principle costs (:128-255)."""
from __future__ import annotations

from .. import dropin
from ..synthetic import assign_responsibility_by_action_space


def calc_risk(traj, ego_state, predictions: dict, scenario, ego_id: int, vehicle_params, params, road_boundary,
              reach_set=None, exec_timer=None, obstacle_types=None, boundary_harm=None):
    """Per-obstacle maxima over time of ego / obstacle risk and harm, plus the boundary harm.

    Returns ``(ego_risk_max, obst_risk_max, ego_harm_max, obst_harm_max, boundary_harm)`` like the reference.
    The road-boundary query is pycrcc geometry (risk_costs.py:104-123) and stays with the caller: pass its result as
    ``boundary_harm`` (default 0 = road never left).
    """
    types = dropin.obstacle_type_names(scenario, predictions, obstacle_types)
    res = dropin.score_trajectories([traj], predictions, types, vehicle_params, params)
    oids = res.obstacle_ids
    T = len(traj.x)
    dicts = [{}, {}, {}, {}]
    for j, oid in enumerate(oids):
        if min(T - 1, len(predictions[oid]["pos_list"])) == 0:  # harm_estimation.py:235-236: no entry at all
            continue
        for r in range(4):
            dicts[r][oid] = float(res.red[r, 0, j])
    return dicts[0], dicts[1], dicts[2], dicts[3], (0 if boundary_harm is None else boundary_harm)


def get_bayesian_costs(ego_risk_max, obst_risk_max, boundary_harm):
    """Bayes principle (risk_costs.py:128-150)."""
    if len(ego_risk_max) == 0:
        return 0
    return float(dropin.principle_costs(ego_risk_max, obst_risk_max, None, None, boundary_harm)[0])


def get_equality_costs(ego_risk_max, obst_risk_max):
    """Equality principle (risk_costs.py:153-178)."""
    if len(ego_risk_max) == 0:
        return 0
    return float(dropin.principle_costs(ego_risk_max, obst_risk_max, None, None, 0.0)[1])


def get_maximin_costs(ego_risk_max, obst_risk_max, ego_harm_max, obst_harm_max, boundary_harm, eps=10e-10, scale_factor=10):
    """Maximin principle incl. its inverted gate (risk_costs.py:181-208)."""
    if len(ego_harm_max) == 0:
        return 0
    return float(dropin.principle_costs(ego_risk_max, obst_risk_max, ego_harm_max, obst_harm_max, boundary_harm,
                                        eps=eps, scale_factor=scale_factor)[2])


def get_ego_costs(ego_risk_max, boundary_harm):
    """Ego principle (risk_costs.py:211-226)."""
    if len(ego_risk_max) == 0:
        return 0
    return float(dropin.principle_costs(ego_risk_max, ego_risk_max, None, None, boundary_harm)[4])


def get_responsibility_cost(scenario, traj, ego_state, obst_risk_max, predictions, reach_set, mode="reach_set"):
    """Responsibility by action space (risk_costs.py:229-255 with reach_set None): ``(cost, None)``."""
    if "reach_set" in mode and reach_set is not None:
        raise NotImplementedError("reach-set responsibility (planner/utils/reachable_set.py) is outside the scored path")
    if ego_state is not None and hasattr(ego_state, "position"):
        # the reference re-assigns the flags here (risk_costs.py:246-248), mutating `predictions`
        assign_responsibility_by_action_space(ego_state.position, ego_state.orientation, predictions)
    if len(predictions) == 0:
        return 0, None
    resp = {k: predictions[k]["responsibility"] for k in predictions}
    risk = {k: obst_risk_max[k] for k in predictions}
    zero = {k: 0.0 for k in predictions}
    return float(dropin.principle_costs(zero, risk, None, None, 0.0, responsibility=resp)[3]), None
