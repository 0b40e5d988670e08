"""CPU checks of the goal-logic restatement (oracle/goal.py) against the fixtures the reference's own get_cost_factor /
velocity_costs wrote (tests/golden/make_golden_goal.py), loop form and vectorised form."""
import os

import numpy as np

from helpers import GOLDEN_DIR, goal_cases
from oracle import goal as ogoal
from oracle import port, vec


def _golden():
    return np.load(os.path.join(GOLDEN_DIR, "goal_cases.npz"))


def test_every_branch_of_the_cost_factor_is_in_the_fixtures():
    z = _golden()
    seen = set()
    for name, _ in goal_cases():
        seen |= set(np.unique(z[f"{name}_factor"]).tolist())
    assert seen == {1.0, 0.01, 0.0001, 10.0}


def test_loop_restatement_matches_the_reference():
    z = _golden()
    for name, step in goal_cases():
        fps = port.calc_frenet_trajectories(step.c_s, step.c_s_d, step.c_s_dd, step.c_d, step.c_d_d, step.c_d_dd, step.d_list,
                                            step.t_list, step.v_list, step.dt, step.csp, step.v_thr)
        factor = np.array([ogoal.get_cost_factor(step.goal, fp, step.dt) for fp in fps])
        vel = np.array([ogoal.velocity_costs(step.goal, fp, step.dt) for fp in fps])
        np.testing.assert_array_equal(factor, z[f"{name}_factor"], err_msg=name)
        np.testing.assert_allclose(vel, z[f"{name}_velocity"], rtol=1e-13, atol=1e-13, err_msg=name)


def test_vectorised_restatement_matches_the_reference():
    z = _golden()
    for name, step in goal_cases():
        ref = vec.score_dense(step)
        live = ref["level"] != vec.LEVEL_SKIPPED
        np.testing.assert_array_equal(ref["cost"][11, live], z[f"{name}_factor"], err_msg=name)
        np.testing.assert_allclose(ref["cost"][6, live], z[f"{name}_velocity"], rtol=1e-12, atol=1e-12, err_msg=name)


def test_time_step_truncation_quirk():
    """int(ego_time + t / dt) truncates: sample 43 of a 0.1 s grid is time step 42 (4.3 / 0.1 = 42.999...), so no sample
    is ever at time step 43."""
    from ethicaltrajectoryplanning_b200.synthetic import GoalContext

    t = np.arange(0.0, 5.0, 0.1)
    assert int(0 + t[43] / 0.1) == 42
    g = GoalContext(time_step=(43, 43), ego_time_step=0)
    assert [i for i, ti in enumerate(t) if ogoal.reached_target_time(g, ti, 0.1)] == []
    g = GoalContext(time_step=(42, 42), ego_time_step=0)
    assert [i for i, ti in enumerate(t) if ogoal.reached_target_time(g, ti, 0.1)] == [42, 43]


def test_check_orientation_folding():
    assert ogoal.check_orientation(0.1, (-0.5, 0.5)) and not ogoal.check_orientation(0.6, (-0.5, 0.5))
    # one end folded: (3.0, 6.5) -> sorted (0.2168, 3.0), test inverted
    assert ogoal.check_orientation(0.1, (3.0, 6.5)) and not ogoal.check_orientation(1.0, (3.0, 6.5))
    # both ends folded: plain interval again
    assert ogoal.check_orientation(0.0, (6.0, 6.6)) and not ogoal.check_orientation(1.0, (6.0, 6.6))


def test_dist_to_goal_pos_restatement_matches_the_reference():
    """calc_dist_to_goal_pos (calc_trajectory_cost.py:760-803): goal-shape and survival branches against the reference's own
    function; the lanelet branch (closest point of a centre line, :835-853) against hand-checked geometry."""
    from ethicaltrajectoryplanning_b200.synthetic import GoalContext

    z = _golden()
    n = 0
    for name, step in goal_cases():
        key = f"{name}_dist_goal"
        if key not in z.files:
            continue
        fps = port.calc_frenet_trajectories(step.c_s, step.c_s_d, step.c_s_dd, step.c_d, step.c_d_d, step.c_d_dd, step.d_list,
                                            step.t_list, step.v_list, step.dt, step.csp, step.v_thr)
        got = np.array([ogoal.calc_dist_to_goal_pos(step.goal, fp.x[-1], fp.y[-1]) for fp in fps], dtype=np.float64)
        np.testing.assert_allclose(got, z[key], rtol=1e-14, atol=0, err_msg=name)
        n += 1
    assert n >= 10
    g = GoalContext(goal_lines=[[(0.0, 0.0), (10.0, 0.0), (10.0, 10.0)], [(20.0, 5.0)]])
    assert ogoal.calc_dist_to_goal_pos(g, 4.0, 3.0) == 3.0            # foot of the perpendicular on the first segment
    assert ogoal.calc_dist_to_goal_pos(g, -3.0, 4.0) == 5.0           # before the first vertex: the vertex itself
    assert ogoal.calc_dist_to_goal_pos(g, 12.0, 12.0) == np.hypot(2.0, 2.0)
    assert ogoal.calc_dist_to_goal_pos(g, 19.0, 5.0) == 1.0           # a one-vertex centre line


def test_dist_to_goal_pos_enters_the_weighted_cost_of_both_oracles():
    import copy

    name, step = [c for c in goal_cases() if c[0] == "velocity_window_narrow"][0]   # two lanelet centre lines
    step = copy.copy(step)
    step.params = copy.deepcopy(step.params)
    step.params["weights"]["dist_to_goal_pos"] = 2.5
    ref = vec.score_dense(step)
    best, valid, invalid, allfp = port.plan_step(step)
    live = np.flatnonzero(ref["level"] != vec.LEVEL_SKIPPED)
    for li, c in enumerate(live):
        fp = allfp[li]
        if hasattr(fp, "cost_dict"):
            assert abs(fp.cost_dict["unweighted_cost"]["dist_to_goal_pos"] - ref["cost"][13, c]) <= 1e-12
            assert abs(fp.cost - ref["cost"][12, c]) <= 1e-10 * abs(fp.cost)
    assert ref["best"]["list_index"] == allfp.index(best)
