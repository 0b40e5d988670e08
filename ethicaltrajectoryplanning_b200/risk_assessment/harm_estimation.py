"""Drop-in for ``risk_assessment/harm_estimation.py:get_harm`` (:202-338, simplified crash-angle branch)."""
from __future__ import annotations

import numpy as np

from .. import dropin


def get_harm(scenario, traj, predictions, ego_id, vehicle_params, modes, coeffs, timer=None, obstacle_types=None):
    """Harm of ego and obstacle per time step: two dicts ``{obstacle_id: ndarray[min(len(traj.t) - 1, Tp)]}``.

    ``obstacle_types`` ({id: ObstacleType name}) may replace the ``scenario.obstacle_by_id`` lookups.
    """
    if modes["harm_mode"] not in ("log_reg", "ref_speed", "gidas"):
        raise ValueError("Please select a valid mode for harm estimation (log_reg, ref_speed, gidas)")
    params = dropin.params_with(modes=modes, coeffs=coeffs)
    types = dropin.obstacle_type_names(scenario, predictions, obstacle_types)
    res = dropin.score_trajectories([traj], predictions, types, vehicle_params, params, want_harm=True)
    T = len(traj.x)
    ego, obst = {}, {}
    for j, oid in enumerate(res.obstacle_ids):
        n = min(T - 1, len(predictions[oid]["pos_list"]))  # harm_estimation.py:234
        if n == 0:
            continue
        ego[oid] = np.array(res.harm[0, 0, :n, j])
        obst[oid] = np.array(res.harm[1, 0, :n, j])
    return ego, obst
