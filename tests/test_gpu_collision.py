"""Device-side collision test with the predicted obstacle boxes (validity level 1; ETP_COLLISION_DISCRETE / _SWEPT,
SURVEY.md 8f N1) against oracle/boxes.py: bit-exact levels and reasons, identical winner, in fp32 and fp64, on generated
candidates and on caller-provided trajectories."""
import copy
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

_spec = importlib.util.spec_from_file_location("_rs", os.path.join(os.path.dirname(__file__), "test_gpu_random_scenes.py"))
_rs = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_rs)


def _with_mode(step, mode):
    step.params = copy.deepcopy(step.params)
    step.params["modes"]["collision_check_device"] = mode
    return step


def _compare(step, precisions=("fp32", "fp64")):
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext
    from oracle import vec

    ref = vec.score_dense(step)
    for precision in precisions:
        ctx = ScoringContext(0, precision)
        try:
            res = ctx.plan(step)
            out = res.numpy()
            np.testing.assert_array_equal(out["level"].astype(np.int64), ref["level"], err_msg=precision)
            np.testing.assert_array_equal(out["reason"].astype(np.int64), ref["reason"], err_msg=precision)
            assert res.best["level"] == ref["best"]["level"]
            if res.best["index"] != ref["best"]["index"]:
                assert precision == "fp32"
                d = ref["dense_index"].tolist()
                assert abs(ref["cost"][12][d.index(res.best["index"])] - ref["best"]["cost"]) <= 1e-4 * abs(ref["best"]["cost"]) + 1e-7
        finally:
            ctx.close()
    return ref


@pytest.mark.parametrize("mode", ["discrete", "swept"])
def test_random_scenes_with_device_collision_check(mode):
    from oracle import vec

    n_hit = n_live = 0
    for seed in range(16):
        ref = _compare(_with_mode(_rs.random_scene(seed), mode))
        n_hit += int(np.sum(ref["level"] == 1))
        n_live += int(np.sum(ref["level"] != vec.LEVEL_SKIPPED))
    assert 0 < n_hit < n_live, (n_hit, n_live)


@pytest.mark.parametrize("mode", ["discrete", "swept"])
@pytest.mark.parametrize("kw", [dict(), dict(lateral_spread=10.0), dict(lateral_spread=8.0, seed=7)])
def test_cfg2_step_with_device_collision_check(mode, kw):
    """The 10^4-candidate step: the stock scene (obstacle boxes on the path, nearly everything collides) and two wider
    ones where about half of the candidates get through."""
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    ref = _compare(_with_mode(make_step("cfg2", **kw), mode))
    hit = int(np.sum(ref["level"] == 1))
    assert (hit >= 9999) if not kw else (3000 < hit < 8000), hit


def test_host_mask_and_device_test_are_combined():
    from ethicaltrajectoryplanning_b200.synthetic import make_step
    from oracle import vec

    step = _with_mode(make_step("planning", seed=11, n_obstacles=6), "discrete")
    C = len(step.v_list) * len(step.t_list) * len(step.d_list)
    step.ext_level = np.full(C, 10, dtype=np.uint8)
    step.ext_level[::7] = 1
    step.ext_level[3::11] = 2
    ref = _compare(step)
    live = ref["level"] != vec.LEVEL_SKIPPED
    assert np.all(ref["level"][live & (step.ext_level[ref["dense_index"]] == 1)] <= 1)


def test_ground_truth_mode_leaves_the_test_to_the_host():
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    step = _with_mode(make_step("planning", seed=12, n_obstacles=6), "swept")
    step.mode = "ground_truth"
    ref = _compare(step, precisions=("fp32",))
    assert not np.any(ref["level"] == 1)


@pytest.mark.parametrize("mode", ["discrete", "swept"])
def test_given_trajectories_with_device_collision_check(mode):
    """sort_frenet_trajectories on an arbitrary fp_list (etp_score_trajectories) against the loop oracle."""
    from ethicaltrajectoryplanning_b200 import dropin
    from ethicaltrajectoryplanning_b200.synthetic import make_step
    from oracle import port

    step = _with_mode(make_step("planning", seed=21, n_obstacles=5, t_list=(2.0, 3.0)), mode)
    oid = list(step.predictions)[1]
    p = step.predictions[oid]
    px, py = step.csp.calc_position(step.c_s + 15.0)
    p["pos_list"] = np.tile(np.array([[px, py - 0.5]]), (len(p["pos_list"]), 1))
    _, _, _, fps = port.plan_step(step)
    want = np.array([fp.valid_level for fp in fps])
    assert 0 < np.sum(want == 1) < len(fps)
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext

    for precision in ("fp32", "fp64"):
        ctx = ScoringContext(0, precision)
        try:
            res = dropin.score_trajectories(fps, step.predictions, step.obstacle_types, step.vehicle, step.params, step.mode,
                                            velocity_cost_mode=step.velocity_cost_mode, velocity_target=step.velocity_target,
                                            cost_factor=step.cost_factor, ctx=ctx)
            np.testing.assert_array_equal(np.asarray(res.level).astype(np.int64), want, err_msg=precision)
        finally:
            ctx.close()


def test_discrete_mode_against_the_reference_written_fixtures():
    """collision_cases.npz: the reference's own create_collision_object + collision_checker_prediction (fallback branch,
    stand-in pycrcc primitives) decided which candidates collide."""
    import importlib.util

    from ethicaltrajectoryplanning_b200 import _abi
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext

    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("_mgc", os.path.join(here, "golden", "make_golden_collision.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    z = np.load(os.path.join(here, "golden", "collision_cases.npz"))
    for name, step, _ in mod.collision_cases():
        want = z[name].astype(bool)
        _with_mode(step, "discrete")
        for precision in ("fp32", "fp64"):
            ctx = ScoringContext(0, precision)
            try:
                lvl = ctx.plan(step).numpy()["level"]
                lvl = lvl[lvl != _abi.LEVEL_SKIPPED].astype(np.int64)
                assert len(lvl) == len(want)
                # level 0 (velocity / curvature) is decided before the collision test
                np.testing.assert_array_equal(lvl[lvl > 0] == 1, want[lvl > 0], err_msg=f"{name} {precision}")
            finally:
                ctx.close()
