"""Goal logic of the cost function on the device (etp_goal) against the fixtures written by the
reference's own get_cost_factor / velocity_costs: cost factors exact, velocity costs and totals within tolerance,
identical winner -- generated candidates, caller-provided trajectories and the drop-in sort_frenet_trajectories."""
import os
from types import SimpleNamespace

import numpy as np
import pytest

from helpers import GOLDEN_DIR, goal_cases

pytestmark = pytest.mark.gpu

CASES = goal_cases()
FACTOR_F32 = {1.0: np.float32(1.0), 0.01: np.float32(0.01), 0.0001: np.float32(0.0001), 10.0: np.float32(10.0)}


def _golden():
    return np.load(os.path.join(GOLDEN_DIR, "goal_cases.npz"))


@pytest.mark.parametrize("name,step", CASES, ids=[n for n, _ in CASES])
def test_generated_candidates(name, step):
    from ethicaltrajectoryplanning_b200 import _abi
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext
    from oracle import vec

    z = _golden()
    ref = vec.score_dense(step)
    for precision, rtol in (("fp32", 1e-5), ("fp64", 1e-12)):
        ctx = ScoringContext(0, precision)
        try:
            res = ctx.plan(step)
            out = res.numpy()
            keep = out["level"] != _abi.LEVEL_SKIPPED
            want = z[f"{name}_factor"]
            if precision == "fp32":
                want = np.array([FACTOR_F32[f] for f in want.tolist()], dtype=np.float32)
            np.testing.assert_array_equal(out["cost"][11, keep], want, err_msg=f"{name} {precision} factor")
            np.testing.assert_allclose(out["cost"][6, keep], z[f"{name}_velocity"], rtol=rtol, atol=1e-5 if precision == "fp32" else 1e-11)
            np.testing.assert_allclose(out["cost"][12, keep], ref["cost"][12, keep], rtol=2e-4 if precision == "fp32" else 1e-9, atol=1e-7)
            np.testing.assert_array_equal(out["level"].astype(np.int64), ref["level"])
            if res.best["index"] != ref["best"]["index"]:
                assert precision == "fp32"
                d = ref["dense_index"].tolist()
                assert abs(ref["cost"][12][d.index(res.best["index"])] - ref["best"]["cost"]) <= 1e-4 * abs(ref["best"]["cost"]) + 1e-7
        finally:
            ctx.close()


def test_given_trajectories_and_drop_in_sort():
    """The same goal through etp_score_trajectories (arbitrary fp_list) and through the drop-in
    sort_frenet_trajectories with a CommonRoad-shaped planning problem and a GoalArea."""
    from ethicaltrajectoryplanning_b200 import dropin
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils import frenet_functions as ff
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils.calc_trajectory_cost import GoalArea
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext
    from oracle import port

    z = _golden()
    for name, step in CASES:
        if name not in ("reached_in_time", "starts_inside_and_leaves", "two_t", "circle_and_concave", "orientation_folded"):
            continue
        g = step.goal
        fps = port.calc_frenet_trajectories(step.c_s, step.c_s_d, step.c_s_dd, step.c_d, step.c_d_d, step.c_d_dd, step.d_list,
                                            step.t_list, step.v_list, step.dt, step.csp, step.v_thr)
        ctx = ScoringContext(0, "fp64")
        try:
            r = dropin.score_trajectories(fps, step.predictions, step.obstacle_types, step.vehicle, step.params, step.mode,
                                          ctx=ctx, goal=g, dt=step.dt)
            np.testing.assert_array_equal(r.cost[11], z[f"{name}_factor"], err_msg=name)
            np.testing.assert_allclose(r.cost[6], z[f"{name}_velocity"], rtol=1e-12, atol=1e-11, err_msg=name)
        finally:
            ctx.close()
        # drop-in level: planning problem with intervals, goal area as plain geometry
        state = SimpleNamespace(time_step=SimpleNamespace(start=g.time_step[0], end=g.time_step[1]))
        if g.velocity is not None:
            state.velocity = SimpleNamespace(start=g.velocity[0], end=g.velocity[1])
        if g.orientation is not None:
            state.orientation = SimpleNamespace(start=g.orientation[0], end=g.orientation[1])
        assert len(g.goal_centers) == 1
        state.position = SimpleNamespace(center=np.asarray(g.goal_centers[0]))
        problem = SimpleNamespace(goal=SimpleNamespace(state_list=[state]))
        ego = SimpleNamespace(position=np.asarray(g.ego_position), orientation=step.ego_orientation, time_step=g.ego_time_step)
        scenario = SimpleNamespace(dt=step.dt, obstacle_by_id=lambda oid, _t=step.obstacle_types: SimpleNamespace(
            obstacle_type=SimpleNamespace(name=_t[oid])))
        fp_list = ff.calc_frenet_trajectories(step.c_s, step.c_s_d, step.c_s_dd, step.c_d, step.c_d_d, step.c_d_dd, step.d_list,
                                              step.t_list, step.v_list, step.dt, step.csp, step.v_thr)
        valid, invalid, _ = ff.sort_frenet_trajectories(
            ego, fp_list, None, step.predictions, step.mode, step.params, problem, scenario, step.vehicle, 1, step.dt, 30.0, None,
            None, GoalArea(g.polygons or (), g.circles or ()))
        got = {id(fp): fp.cost_dict["factor"] for fp in valid}
        want = z[f"{name}_factor"]
        top = [i for i, fp in enumerate(fp_list) if id(fp) in got]
        assert len(top) > 0
        # the drop-in modules run the fp32 context: the factor comes back as the fp32 image of 1, 0.01, 0.0001 or 10
        np.testing.assert_array_equal(np.array([got[id(fp_list[i])] for i in top]).astype(np.float32),
                                      want[top].astype(np.float32), err_msg=name)


def test_dist_to_goal_pos_on_the_device():
    """weights['dist_to_goal_pos'] > 0 (calc_trajectory_cost.py:299-305, 760-803): the distance of every candidate's last
    position to the closest goal position, against the reference's own calc_dist_to_goal_pos (goal-shape / survival
    branches, tests/golden/goal_cases.npz) and the oracle (lanelet centre lines); it enters the weighted cost."""
    import copy

    from ethicaltrajectoryplanning_b200 import _abi
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext
    from oracle import vec

    z = _golden()
    row = _abi.COST_NAMES.index("dist_to_goal_pos")
    checked = 0
    for name, base in CASES:
        if name not in ("reached_in_time", "velocity_window_narrow", "survival", "circle_and_concave", "two_t", "no_centres"):
            continue
        step = copy.copy(base)
        step.params = copy.deepcopy(base.params)
        step.params["weights"]["dist_to_goal_pos"] = 2.5
        ref = vec.score_dense(step)
        for precision, rtol in (("fp32", 1e-5), ("fp64", 1e-12)):
            ctx = ScoringContext(0, precision)
            try:
                res = ctx.plan(step)
                out = res.numpy()
                keep = out["level"] != _abi.LEVEL_SKIPPED
                if f"{name}_dist_goal" in z.files:
                    np.testing.assert_allclose(out["cost"][row, keep], z[f"{name}_dist_goal"], rtol=rtol, atol=1e-12, err_msg=name)
                np.testing.assert_allclose(out["cost"][row], ref["cost"][13], rtol=rtol, atol=1e-12, err_msg=name)
                np.testing.assert_allclose(out["cost"][12, keep], ref["cost"][12, keep], rtol=1e-4 if precision == "fp32" else 1e-9, atol=1e-9)
                assert res.best["index"] == ref["best"]["index"], (name, precision)
                checked += 1
            finally:
                ctx.close()
    assert checked == 12
    # weighted without goal positions: the step is refused, loudly
    step = copy.copy(CASES[0][1])
    step.params = copy.deepcopy(step.params)
    step.params["weights"]["dist_to_goal_pos"] = 1.0
    step.goal = None
    ctx = ScoringContext(0, "fp32")
    try:
        with pytest.raises(Exception, match="dist_to_goal_pos"):
            ctx.plan(step)
    finally:
        ctx.close()


def test_goal_position_without_a_centre_is_refused():
    """A goal state whose position has no centre point and no goal lanelets: the reference's velocity cost averages an empty
    list of distances (calc_trajectory_cost.py:473-497, NaN).  The library refuses the step instead of diverging silently --
    unless the velocity weight is zero, where the reference never evaluates that term (:279)."""
    import copy
    import dataclasses

    from ethicaltrajectoryplanning_b200.scoring import ScoringContext

    name, step = CASES[0]
    step = copy.deepcopy(step)
    step.goal = dataclasses.replace(step.goal, goal_centers=[], goal_points=None, goal_lines=None)
    ctx = ScoringContext(0, "fp32")
    try:
        assert step.params["weights"]["velocity"] > 0.0
        with pytest.raises(NotImplementedError, match="no centre point"):  # ETP_ERR_UNSUPPORTED
            ctx.plan(step, materialize=False)
        step.params["weights"]["velocity"] = 0.0
        assert ctx.plan(step, materialize=False).best["index"] >= 0
    finally:
        ctx.close()
