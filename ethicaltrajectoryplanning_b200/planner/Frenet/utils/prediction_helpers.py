"""Drop-in for ``planner/Frenet/utils/prediction_helpers.py:get_orientation_velocity_and_shape_of_prediction`` (:75-135):
orientation (np.gradient of the predicted positions, "barely moves" rule), velocity and shape margins are derived on the
device (``etp_set_raw_predictions``) and read back into the prediction dict."""
from __future__ import annotations

import numpy as np  # noqa: F401

from .... import dropin


def get_orientation_velocity_and_shape_of_prediction(predictions: dict, scenario, safety_margin_length=1.0,
                                                     safety_margin_width=0.5):
    """Extend the prediction by orientation, velocity and the shape (incl. safety margins) of the predicted obstacle.
    Mutates and returns ``predictions``; predictions without any sample are deleted (:98-100)."""
    ctx = dropin.default_context()
    shapes, ori0, vel0 = {}, {}, {}
    for oid in predictions:
        ob = scenario.obstacle_by_id(oid)
        shapes[oid] = (ob.obstacle_shape.length, ob.obstacle_shape.width)
        ori0[oid] = ob.initial_state.orientation
        vel0[oid] = ob.initial_state.velocity
    keep = {oid: predictions[oid].get("responsibility") for oid in predictions}
    ones = {oid: 1.0 for oid in predictions}
    ctx.set_params(dropin.params_with(), dropin.default_vehicle(), "risk")
    ctx.set_raw_predictions(predictions, shapes, ori0, vel0, None, scenario.dt, (0.0, 0.0), 0.0, safety_margin_length,
                            safety_margin_width, masses=ones, protections={oid: 1 for oid in predictions})
    ctx.enriched(predictions)
    for oid in predictions:  # the reference function does not touch the responsibility flag
        if keep.get(oid) is None:
            predictions[oid].pop("responsibility", None)
        else:
            predictions[oid]["responsibility"] = keep[oid]
    return predictions
