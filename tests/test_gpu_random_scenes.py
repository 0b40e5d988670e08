"""Randomised parity against the vectorised fp64 oracle: reference paths heading in every direction (ego yaw on both
sides of +-pi, un-wrapped angles beyond pi: quirk Q4), obstacles with arbitrary headings (incl. static ones whose stored
orientation lies outside (-pi, pi]), correlations up to +-0.9 (fp64 fallback of the rectangle probability), every harm
model, ragged prediction lengths, low- and high-speed lateral modes."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MODELS = [  # (ignore_angle, sym_angle, reduced_angle_areas)
    (False, True, True), (False, False, True), (False, True, False), (False, False, False), (True, True, True)]


def random_scene(seed):
    from ethicaltrajectoryplanning_b200.configs import default_params
    from ethicaltrajectoryplanning_b200.spline import CubicSpline2D
    from ethicaltrajectoryplanning_b200.synthetic import (PlanningStep, assign_responsibility_by_action_space,
                                                          enrich_predictions, planner_v_list)
    from ethicaltrajectoryplanning_b200.vehicleparams import VehicleParameters

    rng = np.random.default_rng(seed)
    alpha = rng.uniform(-np.pi, np.pi) if seed % 3 else np.pi - 0.02 * (seed % 5)  # every third path runs along the +-pi cut
    x = 5.0 * np.arange(61, dtype=np.float64)
    y = rng.uniform(1.0, 4.0) * np.sin(x / rng.uniform(15.0, 40.0))
    ca, sa = np.cos(alpha), np.sin(alpha)
    off = rng.uniform(-300.0, 300.0, size=2)
    csp = CubicSpline2D(ca * x - sa * y + off[0], sa * x + ca * y + off[1])
    veh = VehicleParameters("bmw_320i")
    v0 = float(rng.choice([1.5, 2.9, 8.0, 15.0, 28.0]))
    t_list = np.array([float(rng.choice([2.0, 3.0]))])
    dt = 0.1
    T = len(np.arange(0.0, t_list[0], dt))
    n_obs = int(rng.integers(1, 9))
    c_s = float(rng.uniform(2.0, 20.0))
    predictions, types, shapes, ori0, vel0 = {}, {}, {}, {}, {}
    k = np.arange(T + 5, dtype=np.float64)
    for i in range(n_obs):
        oid = 200 + i
        n_pred = int(rng.integers(3, T + 5))
        s0 = c_s + rng.uniform(-6.0, 8.0 + v0 * t_list[0])
        lat = rng.uniform(-5.0, 5.0)
        px, py = csp.calc_position(s0)
        yaw_p = float(csp.calc_yaw(s0))
        p0 = np.array([px - lat * np.sin(yaw_p), py + lat * np.cos(yaw_p)])
        heading = rng.uniform(-np.pi, np.pi)
        speed = 0.0 if i % 4 == 1 else rng.uniform(0.5, 15.0)
        pos = p0[None] + (speed * dt * k[:n_pred])[:, None] * np.array([np.cos(heading), np.sin(heading)])[None]
        sx = rng.uniform(0.15, 0.6) + rng.uniform(0.0, 0.08) * k[:n_pred]
        sy = rng.uniform(0.5, 1.2) * sx
        rho = rng.uniform(-0.9, 0.9) if i % 3 else rng.uniform(-0.3, 0.3)
        cov = np.zeros((n_pred, 2, 2))
        cov[:, 0, 0], cov[:, 1, 1] = sx * sx, sy * sy
        cov[:, 0, 1] = cov[:, 1, 0] = rho * sx * sy
        if i % 5 == 4:
            cov[:] = 0.0  # Q14
        predictions[oid] = {"pos_list": pos, "cov_list": cov}
        types[oid] = str(rng.choice(["CAR", "TRUCK", "BUS", "PEDESTRIAN", "BICYCLE", "MOTORCYCLE", "TAXI"]))
        shapes[oid] = (rng.uniform(0.4, 8.0), rng.uniform(0.4, 2.5))
        # static obstacles keep the stored initial orientation, which need not lie in (-pi, pi]
        ori0[oid] = heading + (2 * np.pi * int(rng.integers(-1, 2)) if speed == 0.0 else 0.0)
        vel0[oid] = speed
    enrich_predictions(predictions, shapes, ori0, vel0, dt)
    c_d = float(rng.uniform(-0.5, 0.5))
    px, py = csp.calc_position(c_s)
    yaw = float(csp.calc_yaw(c_s))
    ego_pos = np.array([px - c_d * np.sin(yaw), py + c_d * np.cos(yaw)])
    assign_responsibility_by_action_space(ego_pos, yaw, predictions)
    params = copy.deepcopy(default_params(str(rng.choice(["weights_ethical.json", "weights_ego.json"]))))
    ia, sym, red = MODELS[seed % len(MODELS)]
    params["modes"].update(ignore_angle=ia, sym_angle=sym, reduced_angle_areas=red)
    params["weights"]["lon_jerk"], params["weights"]["travelled_dist"] = 0.1, 0.05
    return PlanningStep(
        c_s=c_s, c_s_d=v0, c_s_dd=float(rng.uniform(-1.0, 1.0)), c_d=c_d, c_d_d=float(rng.uniform(-0.3, 0.3)),
        c_d_dd=float(rng.uniform(-0.2, 0.2)), v_list=planner_v_list(v0, t_list, veh, 12), t_list=t_list,
        d_list=np.linspace(-3.0, 3.0, 41), dt=dt, v_thr=3.0, csp=csp, predictions=predictions, obstacle_types=types,
        ego_position=ego_pos, ego_orientation=yaw, vehicle=veh, params=params, mode="risk", velocity_cost_mode=0,
        velocity_target=v0, cost_factor=1.0)


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("ETP_RANDOM_SCENES", "24"))))
def test_random_scene_against_the_vectorised_oracle(seed):
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext
    from oracle import vec

    step = random_scene(seed)
    ref = vec.score_dense(step)
    for precision, rtol, atol, ctol in (("fp32", 1e-4, 1e-9, 1e-3), ("fp64", 1e-9, 1e-14, 1e-8)):
        ctx = ScoringContext(0, precision)
        try:
            res = ctx.plan(step)
            out = res.numpy()
            np.testing.assert_array_equal(out["level"].astype(np.int64), ref["level"], err_msg=f"{precision} level")
            np.testing.assert_array_equal(out["reason"].astype(np.int64), ref["reason"])
            bad = ~(np.abs(out["prob"] - ref["prob"]) <= rtol * np.abs(ref["prob"]) + atol)
            assert not bad.any(), (precision, "prob", np.argwhere(bad)[:4], out["prob"][bad][:4], ref["prob"][bad][:4])
            bad = ~(np.abs(out["red"] - ref["red"]) <= rtol * np.abs(ref["red"]) + atol)
            assert not bad.any(), (precision, "red", np.argwhere(bad)[:4], out["red"][bad][:4], ref["red"][bad][:4])
            live = ref["level"] != vec.LEVEL_SKIPPED
            bad = ~(np.abs(out["cost"][:, live] - ref["cost"][:, live]) <= ctol * np.abs(ref["cost"][:, live]) + 1e-6)
            assert not bad.any(), (precision, "cost", np.argwhere(bad)[:4])
            # the selection is exact in both modes (fp32: near-ties re-scored in fp64 on the device)
            assert res.best["level"] == ref["best"]["level"]
            assert res.best["index"] == ref["best"]["index"] and res.best["list_index"] == ref["best"]["list_index"]
            # the same step as one planning cycle (etp_plan_cycle: clustered tiles, selection in the fused kernel's last CTA,
            # ragged prediction lengths, every harm model): the same record, twice (direct launch, then captured)
            for rep in range(2):
                best, traj = ctx.plan_cycle(step)
                assert (best["index"], best["level"], best["list_index"], best["n_scored"], best["cost"]) == \
                    (res.best["index"], res.best["level"], res.best["list_index"], res.best["n_scored"], res.best["cost"])
            np.testing.assert_array_equal(traj["x_m"], res.best_trajectory()["x_m"])
        finally:
            ctx.close()
