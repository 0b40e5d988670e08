"""Drop-in for ``planner/utils/responsibility.py:assign_responsibility_by_action_space`` (:57-89): the +-pi/4 view-cone
flag of every predicted obstacle (no angle wrap), derived on the device."""
from __future__ import annotations

import numpy as np

from ... import dropin


def assign_responsibility_by_action_space(scenario, ego_state, predictions):
    """Sets ``predictions[id]['responsibility']`` to 0 inside the ego's view cone, else 1; returns ``predictions``."""
    if len(predictions) == 0:
        return predictions
    ctx = dropin.default_context()
    first = {oid: {"pos_list": np.asarray(p["pos_list"], dtype=np.float64)[:1], "cov_list": np.zeros((1, 2, 2))}
             for oid, p in predictions.items()}
    ones = {oid: 1.0 for oid in predictions}
    ctx.set_params(dropin.params_with(), dropin.default_vehicle(), "risk")
    ctx.set_raw_predictions(first, {oid: (1.0, 1.0) for oid in first}, {oid: 0.0 for oid in first}, {oid: 0.0 for oid in first},
                            None, 0.1, np.asarray(ego_state.position, dtype=np.float64), float(ego_state.orientation),
                            masses=ones, protections={oid: 1 for oid in first})
    ctx.enriched(first)
    for oid in predictions:
        predictions[oid]["responsibility"] = first[oid]["responsibility"]
    return predictions
