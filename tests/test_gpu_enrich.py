"""Prediction enrichment on the device against golden vectors of the reference's own
get_orientation_velocity_and_shape_of_prediction / assign_responsibility_by_action_space."""
import copy
import os
from types import SimpleNamespace

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "enrich_cases.npz")


def load_cases():
    z = np.load(GOLD)
    for c in range(int(z["n_cases"])):
        ids = [int(i) for i in z[f"c{c}_ids"]]
        raw, shapes, ori0, vel0, ref = {}, {}, {}, {}, {}
        for j, oid in enumerate(ids):
            pos = z[f"c{c}_{j}_pos"]
            raw[oid] = {"pos_list": pos.copy(), "cov_list": np.tile(np.eye(2) * 0.1, (len(pos), 1, 1))}
            a = z[f"c{c}_{j}_in"]
            shapes[oid], ori0[oid], vel0[oid] = (float(a[0]), float(a[1])), float(a[2]), float(a[3])
            ref[oid] = (z[f"c{c}_{j}_yaw"], z[f"c{c}_{j}_vel"], z[f"c{c}_{j}_out"])
        yield c, float(z[f"c{c}_dt"]), z[f"c{c}_ego"], raw, shapes, ori0, vel0, ref


def _check(pred, ref):
    for oid, (yaw, vel, out) in ref.items():
        np.testing.assert_allclose(pred[oid]["orientation_list"], yaw, rtol=0, atol=1e-12)
        np.testing.assert_allclose(pred[oid]["v_list"], vel, rtol=1e-12, atol=1e-13)
        assert pred[oid]["shape"] == {"length": out[0], "width": out[1]}
        assert pred[oid]["responsibility"] == int(out[2])


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_device_enrichment_matches_the_reference(precision):
    from ethicaltrajectoryplanning_b200 import dropin
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext

    ctx = ScoringContext(0, precision)
    try:
        ctx.set_params(dropin.params_with(), dropin.default_vehicle(), "risk")
        for c, dt, ego, raw, shapes, ori0, vel0, ref in load_cases():
            pred = copy.deepcopy(raw)
            types = {oid: "CAR" for oid in raw}
            ctx.set_raw_predictions(pred, shapes, ori0, vel0, types, dt, ego[:2], ego[2])
            _check(ctx.enriched(pred), ref)
    finally:
        ctx.close()


def test_reference_named_entry_points():
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils.prediction_helpers import (
        get_orientation_velocity_and_shape_of_prediction)
    from ethicaltrajectoryplanning_b200.planner.utils.responsibility import assign_responsibility_by_action_space

    for c, dt, ego, raw, shapes, ori0, vel0, ref in load_cases():
        sc = SimpleNamespace(dt=dt, obstacle_by_id=lambda oid: SimpleNamespace(
            initial_state=SimpleNamespace(orientation=ori0.get(oid, 0.0), velocity=vel0.get(oid, 0.0)),
            obstacle_shape=SimpleNamespace(length=shapes.get(oid, (1.0, 1.0))[0], width=shapes.get(oid, (1.0, 1.0))[1])))
        pred = copy.deepcopy(raw)
        pred[999] = {"pos_list": np.zeros((0, 2)), "cov_list": np.zeros((0, 2, 2))}   # deleted like the reference does
        out = get_orientation_velocity_and_shape_of_prediction(pred, sc)
        assert out is pred and 999 not in pred and "responsibility" not in pred[next(iter(pred))]
        assign_responsibility_by_action_space(sc, SimpleNamespace(position=ego[:2], orientation=float(ego[2])), pred)
        _check(pred, ref)


def test_resident_raw_predictions_plan_like_host_enriched_ones():
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    step = make_step("planning", seed=5, n_obstacles=7)
    ctx = ScoringContext(0, "fp32")
    try:
        base = ctx.plan(step, materialize=False).best
        raw = {oid: {"pos_list": p["pos_list"], "cov_list": p["cov_list"]} for oid, p in step.predictions.items()}
        # recover the generator's raw shapes / initial states: margins (1.0, 0.5); orientation / velocity of moving
        # obstacles come from the gradient, the static one keeps its initial orientation
        shapes = {oid: (p["shape"]["length"] - 1.0, p["shape"]["width"] - 0.5) for oid, p in step.predictions.items()}
        ori0 = {oid: float(np.asarray(p["orientation_list"])[0]) for oid, p in step.predictions.items()}
        vel0 = {oid: float(np.asarray(p["v_list"])[0]) for oid, p in step.predictions.items()}
        ctx.ensure_params(step.params, step.vehicle, step.mode)
        ctx.set_raw_predictions(raw, shapes, ori0, vel0, step.obstacle_types, step.dt, step.ego_position, step.ego_orientation)
        import dataclasses

        res = ctx.plan(dataclasses.replace(step, predictions=raw, predictions_resident=True), materialize=False).best
        assert (res["index"], res["level"]) == (base["index"], base["level"])
        assert abs(res["cost"] - base["cost"]) <= 1e-9 * abs(base["cost"])
    finally:
        ctx.close()
