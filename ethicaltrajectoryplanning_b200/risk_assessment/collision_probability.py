"""Drop-in for ``risk_assessment/collision_probability.py`` (entry points at :136 and :259).

Same signatures and return types; the arithmetic runs in the fused CUDA kernel through
``etp_score_trajectories`` (``include/etp_b200.h``).
"""
from __future__ import annotations

import copy

import numpy as np

from .. import dropin


def _prob_dict(traj, predictions, vehicle_params, mahalanobis):
    params = dropin.params_with()
    params["modes"] = copy.deepcopy(params["modes"])
    params["modes"]["fast_prob_mahalanobis"] = bool(mahalanobis)
    # the probability does not depend on the obstacle class: score every obstacle as a protected one of mass 1
    res = dropin.score_trajectories([traj], predictions, None, vehicle_params, params,
                                    masses={oid: 1.0 for oid in predictions}, protections={oid: 1 for oid in predictions})
    T = len(traj.x)
    return {oid: np.array(res.prob[0, :T - 1, j]) for j, oid in enumerate(res.obstacle_ids)}


def get_collision_probability_fast(traj, predictions: dict, vehicle_params, safety_margin=1.0):
    """Collision probability per time step and obstacle: ``{obstacle_id: ndarray[len(traj.x) - 1]}``
    (collision_probability.py:136-256; ``safety_margin`` is unused there as well)."""
    return _prob_dict(traj, predictions, vehicle_params, mahalanobis=False)


def get_inv_mahalanobis_dist(traj, predictions: dict, vehicle_params, safety_margin=1.0):
    """``1e-4 / mahalanobis`` surrogate of the collision probability (collision_probability.py:259-292)."""
    return _prob_dict(traj, predictions, vehicle_params, mahalanobis=True)
