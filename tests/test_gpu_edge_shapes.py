"""Edge shapes against the vectorised fp64 oracle: one obstacle, obstacle counts that are not multiples of the
kernel's obstacle blocking, more than 64 obstacles, one-sample predictions, the shortest and the longest trajectories,
a single lateral target, candidate counts that are not multiples of the 32-candidate tile, several horizons in one
step, every candidate skipped (the reference raises), and shapes beyond the kernel's limits (loud errors)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _compare(step, precisions=(("fp32", 1e-4, 1e-9, 1e-3), ("fp64", 1e-9, 1e-14, 1e-8))):
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext
    from oracle import vec

    ref = vec.score_dense(step)
    for precision, rtol, atol, ctol in precisions:
        ctx = ScoringContext(0, precision)
        try:
            res = ctx.plan(step)
            out = res.numpy()
            np.testing.assert_array_equal(out["level"].astype(np.int64), ref["level"])
            np.testing.assert_array_equal(out["reason"].astype(np.int64), ref["reason"])
            if ref["prob"].size:
                assert np.all(np.abs(out["prob"] - ref["prob"]) <= rtol * np.abs(ref["prob"]) + atol), precision
                assert np.all(np.abs(out["red"] - ref["red"]) <= rtol * np.abs(ref["red"]) + atol), precision
            live = ref["level"] != vec.LEVEL_SKIPPED
            assert np.all(np.abs(out["cost"][:, live] - ref["cost"][:, live]) <= ctol * np.abs(ref["cost"][:, live]) + 1e-6)
            traj_err = np.abs(out["traj"] - ref["traj"]) / np.maximum(np.abs(ref["traj"]), 1e-3)
            assert traj_err.max() <= (2e-5 if precision == "fp32" else 1e-6)
            assert res.best["index"] == ref["best"]["index"] and res.best["list_index"] == ref["best"]["list_index"]
            assert res.best["n_scored"] == ref["best"]["n_scored"]
        finally:
            ctx.close()


@pytest.mark.parametrize("n_obstacles", [1, 3, 5, 9, 33, 70])
def test_obstacle_counts(n_obstacles):
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    _compare(make_step("planning", n_obstacles=n_obstacles, nv=6, nd=11, seed=7 + n_obstacles))


@pytest.mark.parametrize("nv,nd", [(1, 1), (3, 1), (1, 7), (7, 9), (2, 33), (5, 13)])
def test_candidate_counts_around_the_tile_size(nv, nd):
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    _compare(make_step("planning", nv=nv, nd=nd, n_obstacles=4, seed=nv * 100 + nd))


def test_one_sample_and_ragged_predictions():
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    step = make_step("planning", nv=4, nd=9, n_obstacles=6, seed=3)
    for j, (oid, p) in enumerate(step.predictions.items()):
        keep = [1, 2, 5, 20, 3, 12][j]
        for k in ("pos_list", "cov_list", "orientation_list", "v_list"):
            p[k] = np.asarray(p[k])[:keep]
    _compare(step)


@pytest.mark.parametrize("t_list", [[0.2], [0.3, 1.0], [1.0, 2.5, 4.0]])
def test_horizons(t_list):
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    # 0.2 s -> 2 samples, the minimum the reference can transform
    _compare(make_step("planning", nv=3, nd=5, n_obstacles=3, t_list=t_list, seed=11, v0=6.0))


def test_longest_horizons_that_fit_the_shared_memory_plan():
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    # the ego samples of a 32-candidate tile live in shared memory: ~230 samples in fp32, ~130 in the fp64 check mode.
    # End velocities are given explicitly: the planner's own rule clamps them to exactly v_max over such horizons, and
    # `|s_d| <= v_max` at exact equality is decided by the last ulp of numpy's vectorised pow in the reference.
    for t_end, prec in ((22.0, ("fp32", 1e-4, 1e-9, 1e-3)), (12.0, ("fp64", 1e-9, 1e-14, 1e-8))):
        step = make_step("planning", nv=3, nd=5, n_obstacles=2, t_list=[t_end], seed=11, v0=6.0)
        step.v_list = np.array([2.0, 9.0, 6.0])
        _compare(step, (prec,))


def test_every_candidate_skipped_raises_like_the_reference():
    from ethicaltrajectoryplanning_b200.scoring import NoCandidateError, ScoringContext
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    step = make_step("planning", nv=3, nd=5, n_obstacles=2, t_list=[0.3], v0=0.05)
    step.v_list = np.array([0.01, 0.02, 0.03])
    step.d_list = np.array([-3.0, -2.0, 2.0, 2.5, 3.0])  # every |dT| exceeds the travelled arc length
    ctx = ScoringContext(0, "fp32")
    try:
        with pytest.raises(NoCandidateError):  # ValueError subclass: max([]) in frenet_functions.py:562-564
            ctx.plan(step)
    finally:
        ctx.close()


def test_shapes_beyond_the_limits_fail_loudly():
    from ethicaltrajectoryplanning_b200.scoring import EtpError, ScoringContext
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    ctx = ScoringContext(0, "fp32")
    try:
        with pytest.raises(EtpError, match="256 samples"):
            ctx.plan(make_step("planning", nv=2, nd=3, n_obstacles=1, t_list=[30.0]))
        with pytest.raises((EtpError, NotImplementedError), match="shared memory"):
            ctx.plan(make_step("planning", nv=2, nd=3, n_obstacles=300, t_list=[20.0], seed=5))
    finally:
        ctx.close()


def test_malformed_inputs_of_the_optional_checks_fail_loudly():
    """Bad collision mode, degenerate goal polygon, non-finite boundary vertex, goal without a time step: errors, not
    silently ignored inputs."""
    import copy

    from ethicaltrajectoryplanning_b200.scoring import EtpError, ScoringContext
    from ethicaltrajectoryplanning_b200.synthetic import GoalContext, make_step

    ctx = ScoringContext(0, "fp32")
    try:
        step = make_step("cfg1", seed=5)
        step.params = copy.deepcopy(step.params)
        step.params["modes"]["collision_check_device"] = "sometimes"
        with pytest.raises(ValueError, match="collision_check_device"):
            ctx.plan(step)
        step = make_step("cfg1", seed=5)
        step.goal = GoalContext(polygons=[[(0.0, 0.0), (1.0, 1.0)]], time_step=(0, 10), ego_position=(0.0, 0.0))
        with pytest.raises(EtpError, match="fewer than 3 vertices"):
            ctx.plan(step)
        tris = np.zeros((2, 3, 2))
        tris[1, 2, 0] = np.nan
        with pytest.raises(EtpError, match="not finite"):
            ctx.set_road_boundary(tris, "swept")
        with pytest.raises(KeyError):
            ctx.set_road_boundary(np.zeros((1, 3, 2)), "maybe")
        # the context still works afterwards
        step = make_step("cfg1", seed=5)
        assert ctx.plan(step).best["index"] >= 0
    finally:
        ctx.close()
