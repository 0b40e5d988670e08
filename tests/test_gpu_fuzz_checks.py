"""Randomised parity of the optional device-side checks switched on TOGETHER (collision with the predicted boxes, road
boundary, goal logic) against the vectorised fp64 oracle: random paths in every direction, random boundary widths, random
goal polygons / intervals, both box modes; plus the same scenes through the batched sweep entry point."""
import copy
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

_spec = importlib.util.spec_from_file_location("_rs2", os.path.join(os.path.dirname(__file__), "test_gpu_random_scenes.py"))
_rs = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_rs)


def scene(seed):
    from ethicaltrajectoryplanning_b200.synthetic import GoalContext, road_boundary_triangles

    rng = np.random.default_rng(10_000 + seed)
    step = _rs.random_scene(seed)
    step.params = copy.deepcopy(step.params)
    mode = ("discrete", "swept")[seed % 2]
    step.params["modes"]["collision_check_device"] = ("host", "discrete", "swept")[seed % 3]
    step.road_boundary = (road_boundary_triangles(step.csp, half_width=float(rng.uniform(2.4, 4.5)), outer=float(rng.uniform(8.0, 15.0)),
                                                  ds=float(rng.uniform(2.5, 7.0)), s_end=min(float(step.csp.s[-1]), step.c_s + 140.0)), mode)
    # goal: a quadrilateral across the road somewhere ahead (sometimes around the ego), random intervals
    s0 = step.c_s + (rng.uniform(-3.0, 3.0) if seed % 4 == 0 else rng.uniform(8.0, 45.0))
    ring = []
    for ds, side in ((0.0, 1), (rng.uniform(8.0, 30.0), 1), (rng.uniform(8.0, 30.0), -1), (0.0, -1)):
        px, py = step.csp.calc_position(s0 + ds)
        yaw = step.csp.calc_yaw(s0 + ds)
        w = rng.uniform(1.0, 5.0)
        ring.append((px - side * w * np.sin(yaw), py + side * w * np.cos(yaw)))
    yaw0 = float(step.csp.calc_yaw(s0))
    step.goal = GoalContext(
        polygons=[ring], circles=[(ring[1][0], ring[1][1], float(rng.uniform(0.5, 3.0)))] if seed % 5 == 0 else None,
        velocity=None if seed % 3 == 0 else tuple(sorted(rng.uniform(0.0, 30.0, size=2))),
        orientation=None if seed % 4 == 1 else (yaw0 - rng.uniform(0.1, 1.0) + (2 * np.pi if seed % 6 == 2 else 0.0),
                                                yaw0 + rng.uniform(0.1, 1.0) + (2 * np.pi if seed % 6 in (2, 3) else 0.0)),
        time_step=(int(rng.integers(0, 30)), int(rng.integers(30, 80))), ego_time_step=int(rng.integers(0, 40)),
        ego_position=tuple(step.ego_position), goal_centers=[ring[1], ring[2]] if seed % 7 else None)
    return step


@pytest.mark.parametrize("seed", range(int(os.environ.get("ETP_FUZZ_SCENES", "20"))))
def test_all_device_checks_together(seed):
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext
    from oracle import vec

    step = scene(seed)
    ref = vec.score_dense(step)
    live = ref["level"] != vec.LEVEL_SKIPPED
    for precision, rtol in (("fp32", 2e-4), ("fp64", 1e-9)):
        ctx = ScoringContext(0, precision)
        try:
            res = ctx.plan(step)
            out = res.numpy()
            np.testing.assert_array_equal(out["level"].astype(np.int64), ref["level"], err_msg=precision)
            np.testing.assert_array_equal(out["reason"].astype(np.int64), ref["reason"], err_msg=precision)
            np.testing.assert_array_equal(out["cost"][11, live].astype(np.float32), ref["cost"][11, live].astype(np.float32))
            np.testing.assert_allclose(out["cost"][:, live], ref["cost"][:, live], rtol=rtol, atol=1e-6, err_msg=precision)
            assert res.best["level"] == ref["best"]["level"]
            assert res.best["index"] == ref["best"]["index"]  # exact selection in both modes
            # the same step as one planning cycle (etp_plan_cycle with goal context, road boundary and collision check:
            # clustered tiles fold collision flags and first boundary contacts through distributed shared memory)
            for rep in range(2):
                best, traj = ctx.plan_cycle(step)
                assert (best["index"], best["level"], best["list_index"], best["cost"]) == \
                    (res.best["index"], res.best["level"], res.best["list_index"], res.best["cost"])
        finally:
            ctx.close()


def test_batched_sweep_honours_goal_and_road_boundary():
    """etp_plan_batch: every scenario carries its own goal; the road boundary is the context's."""
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext
    from ethicaltrajectoryplanning_b200.synthetic import GoalContext, make_step, road_boundary_triangles
    from oracle import vec

    steps = []
    for k in range(6):
        st = make_step("planning", seed=40 + k, n_obstacles=3 + k % 3, nv=6, nd=9)
        st.goal = GoalContext(polygons=[[(20.0 + k, -1.0), (60.0, -1.0), (60.0, 6.0), (20.0 + k, 6.0)]], velocity=(4.0, 25.0),
                              time_step=(5 + k, 40), ego_time_step=k, ego_position=tuple(st.ego_position), goal_centers=[(40.0, 2.0)])
        steps.append(st)
    tris = road_boundary_triangles(steps[0].csp, s_end=140.0)
    ctx = ScoringContext(0, "fp32")
    try:
        ctx.set_params(steps[0].params, steps[0].vehicle, steps[0].mode)
        ctx.set_road_boundary(tris, "swept")
        bests = ctx.plan_batch(steps)
        for st, b in zip(steps, bests):
            st.road_boundary = (tris, "swept")
            ref = vec.score_dense(st)
            assert b["level"] == ref["best"]["level"]
            if b["index"] != ref["best"]["index"]:
                d = ref["dense_index"].tolist()
                assert abs(ref["cost"][12][d.index(b["index"])] - ref["best"]["cost"]) <= 1e-4 * abs(ref["best"]["cost"]) + 1e-7
    finally:
        ctx.close()


@pytest.mark.parametrize("seed,offset", [(1, (4.1e5, 5.4e6)), (3, (-7.7e5, 2.2e6)), (6, (3.3e4, -9.1e4))])
def test_map_coordinates_far_from_the_origin(seed, offset):
    """UTM-sized map coordinates: the fp32 path works relative to a step origin, so shifting a whole scene (path,
    predictions, ego, boundary, goal) by hundreds of kilometres changes nothing but the absolute trajectory fields."""
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext
    from ethicaltrajectoryplanning_b200.spline import CubicSpline2D
    from ethicaltrajectoryplanning_b200.synthetic import GoalContext
    from oracle import vec

    step = scene(seed)
    dx, dy = offset
    cx, cy = step.csp.coef_x.copy(), step.csp.coef_y.copy()
    cx[:, 0] += dx
    cy[:, 0] += dy
    step.csp = CubicSpline2D(knots=step.csp.s, coef_x=cx, coef_y=cy)
    for p in step.predictions.values():
        p["pos_list"] = np.asarray(p["pos_list"]) + np.array([dx, dy])
    step.ego_position = np.asarray(step.ego_position) + np.array([dx, dy])
    step.road_boundary = (step.road_boundary[0] + np.array([dx, dy]), step.road_boundary[1])
    g = step.goal
    step.goal = GoalContext(polygons=[np.asarray(q) + np.array([dx, dy]) for q in g.polygons],
                            circles=None if g.circles is None else [(c[0] + dx, c[1] + dy, c[2]) for c in g.circles],
                            velocity=g.velocity, orientation=g.orientation, time_step=g.time_step, ego_time_step=g.ego_time_step,
                            ego_position=tuple(step.ego_position),
                            goal_centers=None if g.goal_centers is None else [(c[0] + dx, c[1] + dy) for c in g.goal_centers])
    ref = vec.score_dense(step)
    live = ref["level"] != vec.LEVEL_SKIPPED
    # at 10^6 m one fp64 ulp is 2e-10 m; the reference's chained difference quotients over millimetre arclength steps
    # (creeping profiles) turn that into ~1e-6 rad of heading noise, i.e. ~1e-5 relative on a probability: the fp64
    # oracle and the fp64 kernel both sit on that floor, the fp32 path must not be worse than its usual bar by more
    for precision, rtol, ptol in (("fp32", 5e-4, 3e-4), ("fp64", 1e-4, 3e-5)):
        ctx = ScoringContext(0, precision)
        try:
            res = ctx.plan(step)
            out = res.numpy()
            np.testing.assert_array_equal(out["level"].astype(np.int64), ref["level"], err_msg=precision)
            bad = ~(np.abs(out["prob"] - ref["prob"]) <= ptol * np.abs(ref["prob"]) + 1e-9)
            assert not bad.any(), (precision, np.argwhere(bad)[:4], out["prob"][bad][:4], ref["prob"][bad][:4])
            np.testing.assert_allclose(out["cost"][:, live], ref["cost"][:, live], rtol=rtol, atol=1e-6, err_msg=precision)
            assert res.best["level"] == ref["best"]["level"]
        finally:
            ctx.close()
