"""The selected trajectory is the reference's, whatever fp32 does to the costs.

`north_star`: identical selected-trajectory index (ties broken by index) on every config, bit-exact validity masks.
The fused kernel scores in fp32; etp_select.cu re-decides the arg-min among the near-ties, the max-risk level and the
maximin gate in fp64.  Checked here against
  * the committed full-size fixtures of the vectorised fp64 oracle (tests/golden/fullsize_cfg2.npz / fullsize_cfg3.npz,
    written by tests/golden/make_golden_fullsize.py over EVERY profile: winner, level histogram and level digest per
    profile, per-profile minima) -- unsharded and 2 / 4 / 8-way sharded,
  * adversarial steps whose runner-up is 1e-7 ... 1e-10 (relative) behind or ahead of the cheapest candidate,
  * bit-identical duplicate candidates (the lower index wins),
  * the selection's own diagnostics: no re-scored candidate may lie outside the bound the fp32 pass claimed for it.
"""
import copy
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RANK = {0: 0, 1: 1, 2: 2, 3: 3, 10: 4}


def _ctx(precision="fp32"):
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext

    return ScoringContext(0, precision)


def _check_against_fixture(step, res, fx, rtol):
    from ethicaltrajectoryplanning_b200 import _abi

    assert res.best["index"] == int(fx["winner_index"])
    assert res.best["list_index"] == int(fx["winner_list_index"])
    assert res.best["level"] == int(fx["winner_level"])
    assert res.best["n_scored"] == int(fx["n_scored"])
    # the key's cost is the fp64 one when the winner had to be told apart from near-ties, else the fp32 pass's own
    assert abs(res.best["cost"] - float(fx["winner_cost"])) <= (1e-9 if rtol < 1e-6 else 1e-5) * abs(float(fx["winner_cost"]))
    out = res.numpy()
    nd = len(step.d_list)
    lv = out["level"].astype(np.int64).reshape(-1, nd)
    hist = np.stack([(lv == k).sum(axis=1) for k in (0, 1, 2, 3, 10, 255)], axis=1)
    np.testing.assert_array_equal(hist, fx["level_hist"])                       # bit-exact levels, every candidate
    digest = (np.where(lv == 255, 0, lv + 1) * (np.arange(nd)[None, :] + 1)).sum(axis=1)
    np.testing.assert_array_equal(digest, fx["level_digest"])
    total = out["cost"][_abi.COST_NAMES.index("total")].astype(np.float64).reshape(-1, nd)
    for k, r in RANK.items():
        c = np.where(lv == k, total, np.inf).min(axis=1)
        want = fx["min_cost"][:, r]
        both = np.isfinite(want)
        np.testing.assert_array_equal(np.isfinite(c), both)
        assert np.all(np.abs(c[both] - want[both]) <= rtol * np.abs(want[both]) + 1e-9)
    # the cheapest candidates of the top level: their materialised costs agree with the oracle's
    ti = fx["top_index"][:512]
    got = total.ravel()[ti]
    assert np.all(np.abs(got - fx["top_cost"][:512]) <= rtol * np.abs(fx["top_cost"][:512]) + 1e-9)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_cfg2_full_step_matches_the_full_size_oracle_fixture(precision):
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    fx = np.load(os.path.join(GOLDEN, "fullsize_cfg2.npz"))
    step = make_step("cfg2")
    ctx = _ctx(precision)
    try:
        res = ctx.plan(step)
        _check_against_fixture(step, res, fx, 1e-4 if precision == "fp32" else 1e-9)
    finally:
        ctx.close()


def test_cfg3_full_step_matches_the_full_size_oracle_fixture_unsharded_and_sharded():
    from ethicaltrajectoryplanning_b200.distributed import combine_best_records
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    fx = np.load(os.path.join(GOLDEN, "fullsize_cfg3.npz"))
    step = make_step("cfg3")
    ctx = _ctx("fp32")
    try:
        res = ctx.plan(step)
        _check_against_fixture(step, res, fx, 1e-4)
        dbg = ctx.debug_counters()
        assert dbg["rescored"] >= 2, dbg            # the runner-up is 2.5e-8 (relative) behind: fp32 cannot tell them apart
        assert dbg["bound_violations"] == 0, dbg
        lean = ctx.plan(step, materialize=False, configure=False)
        assert lean.best["index"] == int(fx["winner_index"]) and lean.best["cost"] == res.best["cost"]
        for world in (2, 4, 8):
            recs = []
            for r in range(world):
                ctx.plan(step, materialize=False, configure=False, shard_index=r, shard_count=world)
                recs.append(ctx.best_raw())
            rc, out = combine_best_records(recs)
            assert rc == 0 and out.index == int(fx["winner_index"]) and out.list_index == int(fx["winner_list_index"])
            assert out.n_scored == int(fx["n_scored"]) and out.cost == res.best["cost"]
        assert ctx.debug_counters()["bound_violations"] == 0
    finally:
        ctx.close()


def _tie_steps():
    """(name, step, expected dense index): the planning.json grid with per-candidate cost factors that put one chosen
    candidate a relative `gap` behind / ahead of the oracle's cheapest one."""
    from ethicaltrajectoryplanning_b200.synthetic import make_step
    from oracle import vec

    base = make_step("planning", seed=77)
    ref = vec.score_dense(base)
    lv, cost = ref["level"], ref["cost"][12]
    top = np.flatnonzero(lv == lv[lv != 255].max() if 10 not in lv else lv == 10)
    a = int(top[np.argmin(cost[top])])
    others = [int(c) for c in top if c != a and cost[c] > 0]
    lower = [c for c in others if c < a][-1:] + [c for c in others if c < a][:1]
    higher = [c for c in others if c > a][:1] + [c for c in others if c > a][-1:]
    out = []
    for gap in (1e-7, 1e-8, 1e-9, 1e-10):
        for b in lower + higher:
            for sign in (+1, -1):
                step = copy.copy(base)
                fl = np.full(base.n_dense, 1e3)
                fl[a] = 1.0
                fl[b] = cost[a] * (1.0 + sign * gap) / cost[b]
                step.cost_factor_list = fl
                want = vec.score_dense(step)["best"]["index"]
                assert want == (a if sign > 0 else b)
                out.append((f"gap{gap:g}_b{b}_{'behind' if sign > 0 else 'ahead'}", step, want))
    return out


def test_runner_up_within_1e_7_to_1e_10_of_the_winner_is_told_apart():
    ctx32, ctx64 = _ctx("fp32"), _ctx("fp64")
    try:
        for name, step, want in _tie_steps():
            for ctx in (ctx32, ctx64):
                got = ctx.plan(step, materialize=False).best["index"]
                assert got == want, (name, got, want)
                got = ctx.plan(step, materialize=True).best["index"]
                assert got == want, (name, got, want)
        dbg = ctx32.debug_counters()
        assert dbg["rescored"] > 0 and dbg["bound_violations"] == 0, dbg
    finally:
        ctx32.close()
        ctx64.close()


def test_bit_identical_duplicates_go_to_the_lower_index():
    """A duplicated end velocity (quirk Q8: v_cur appended to a linspace that already ends there) gives bit-identical
    candidates; the reference's stable sort keeps the first."""
    from ethicaltrajectoryplanning_b200.synthetic import make_step
    from oracle import vec

    step = make_step("planning", seed=5)
    ref0 = vec.score_dense(step)
    iv = ref0["best"]["index"] // (len(step.t_list) * len(step.d_list))
    step.v_list = np.append(step.v_list, step.v_list[iv])  # the winner's profile once more, at the end
    ref = vec.score_dense(step)
    assert ref["best"]["index"] == ref0["best"]["index"]
    for precision in ("fp32", "fp64"):
        ctx = _ctx(precision)
        try:
            res = ctx.plan(step)
            assert res.best["index"] == ref["best"]["index"] and res.best["list_index"] == ref["best"]["list_index"]
            total = res.numpy()["cost"][12]
            dup = (len(step.v_list) - 1) * len(step.t_list) * len(step.d_list) + ref["best"]["index"] % len(step.d_list)
            assert total[dup] == total[ref["best"]["index"]]
            # survivors that are one and the same computation are settled by generation order, without an fp64 re-scoring --
            # unless a third candidate is within the fp32 bound as well
            before = ctx.debug_counters()["rescored"]
            best, _ = ctx.plan_winner(step)
            assert best["index"] == ref["best"]["index"] and best["list_index"] == ref["best"]["list_index"]
            assert ctx.debug_counters()["rescored"] - before in (0, 3)
            best, _ = ctx.plan_cycle(step)
            assert best["index"] == ref["best"]["index"] and best["list_index"] == ref["best"]["list_index"]
        finally:
            ctx.close()


def test_max_risk_level_and_maximin_gate_close_to_their_thresholds():
    """Candidates whose ego-risk sum sits within fp32 rounding of max_acceptable_risk, and obstacle risks around the
    maximin gate's 1e-9: levels stay bit-exact and costs on the fp64 side of the decision."""
    from ethicaltrajectoryplanning_b200 import _abi
    from ethicaltrajectoryplanning_b200.synthetic import make_step
    from oracle import vec

    base = make_step("planning", seed=11, lateral_spread=1.5, n_obstacles=6)
    ref = vec.score_dense(base)
    live = ref["level"] != 255
    sums = ref["red"][0].sum(axis=1)[live]
    assert sums.max() > 0
    checked = 0
    for q in (0.3, 0.5, 0.7):
        # put the threshold ON a candidate's own ego-risk sum (and a hair either side of it)
        target = float(np.sort(sums)[int(q * (len(sums) - 1))])
        if target <= 0:
            continue
        for scale in (1.0, 1.0 + 3e-7, 1.0 - 3e-7):
            step = copy.copy(base)
            step.params = copy.deepcopy(base.params)
            step.params["modes"]["max_acceptable_risk"] = target * scale
            want = vec.score_dense(step)
            ctx = _ctx("fp32")
            try:
                res = ctx.plan(step)
                out = res.numpy()
                np.testing.assert_array_equal(out["level"].astype(np.int64), want["level"])
                np.testing.assert_array_equal(out["reason"].astype(np.int64), want["reason"])
                assert res.best["index"] == want["best"]["index"]
                assert ctx.debug_counters()["bound_violations"] == 0
                checked += 1
            finally:
                ctx.close()
    assert checked >= 6


def _near_tie_variants(base, gaps=(1e-8, 1e-10)):
    """Steps derived from `base` whose runner-up is a relative gap behind / ahead of the oracle's winner (per-candidate cost
    factors), with the expected dense index."""
    from oracle import vec

    ref = vec.score_dense(base)
    lv, cost = ref["level"], ref["cost"][12]
    live = lv != 255
    top_level = 10 if (lv[live] == 10).any() else int(lv[live].max())
    top = np.flatnonzero(lv == top_level)
    a = int(top[np.argmin(cost[top])])
    others = [int(c) for c in top if c != a and cost[c] > 0]
    picks = ([c for c in others if c < a][-1:] + [c for c in others if c > a][:1])[:2]
    out = []
    for gap in gaps:
        for b in picks:
            for sign in (+1, -1):
                step = copy.copy(base)
                fl = np.full(base.n_dense, 1e3)
                fl[a] = 1.0
                fl[b] = cost[a] * (1.0 + sign * gap) / cost[b]
                step.cost_factor_list = fl
                want = vec.score_dense(step)["best"]
                assert want["index"] == (a if sign > 0 else b)
                out.append((step, want))
    return out


def test_near_ties_with_every_optional_stage_switched_on():
    """The exact re-scoring reproduces everything that enters a candidate's cost and level: Mahalanobis probabilities,
    road boundary harm and collision with the predicted boxes on the device, all harm models, risk-only weights below
    level 10, jerk / travelled-distance terms."""
    from ethicaltrajectoryplanning_b200.synthetic import make_step, road_boundary_triangles

    bases = []
    s = make_step("planning", seed=21, lateral_spread=2.0, nv=7, nd=9)
    s.params = copy.deepcopy(s.params)
    s.params["modes"]["fast_prob_mahalanobis"] = True
    for p in s.predictions.values():
        if not np.any(p["cov_list"]):
            p["cov_list"][:, 0, 0] = p["cov_list"][:, 1, 1] = 0.1
    bases.append(("mahalanobis", s))
    s = make_step("planning", seed=22, lateral_spread=5.0, nv=7, nd=9)
    s.road_boundary = (road_boundary_triangles(s.csp, half_width=2.5), "discrete")
    s.params = copy.deepcopy(s.params)
    s.params["modes"]["collision_check_device"] = "discrete"
    bases.append(("boundary_and_collision", s))
    for name, (ign, sym, red) in {"lr1s": (True, True, True), "lr12a": (False, False, False), "lr4a": (False, False, True)}.items():
        s = make_step("planning", seed=23, lateral_spread=1.5, nv=7, nd=9)
        s.params = copy.deepcopy(s.params)
        s.params["modes"].update(ignore_angle=ign, sym_angle=sym, reduced_angle_areas=red)
        bases.append((name, s))
    s = make_step("planning", seed=24, lateral_spread=1.0, nv=7, nd=9)
    s.params = copy.deepcopy(s.params)
    s.params["modes"]["multiple_cost_functions"] = True
    s.params["modes"]["max_acceptable_risk"] = 1e-9   # everything level 3: risk-only weights
    s.params["weights"].update(lon_jerk=0.3, lat_jerk=0.2, travelled_dist=0.1)
    bases.append(("multi_cost_level3", s))
    s = make_step("planning", seed=25, nv=7, nd=9, weights="weights_ego.json")
    s.params = copy.deepcopy(s.params)
    s.params["weights"].update(lon_jerk=0.3, lat_jerk=0.2, travelled_dist=0.1)
    bases.append(("ego_weights_jerk", s))
    ctx = _ctx("fp32")
    try:
        n = 0
        for name, base in bases:
            for step, want in _near_tie_variants(base):
                res = ctx.plan(step)
                assert res.best["index"] == want["index"], (name, res.best, want)
                assert res.best["level"] == want["level"], name
                assert abs(res.best["cost"] - want["cost"]) <= 1e-9 * abs(want["cost"]) + 1e-15, (name, res.best["cost"], want["cost"])
                n += 1
        dbg = ctx.debug_counters()
        assert n >= 40 and dbg["rescored"] >= n and dbg["bound_violations"] == 0, dbg
    finally:
        ctx.close()


def test_near_ties_among_caller_provided_trajectories():
    """etp_score_trajectories (arbitrary fp_list): duplicates and a 1e-10 runner-up are told apart like generated candidates."""
    from ethicaltrajectoryplanning_b200 import dropin
    from ethicaltrajectoryplanning_b200.synthetic import make_step
    from oracle import port, vec

    step = make_step("planning", seed=31, nv=5, nd=7)
    fps = port.calc_frenet_trajectories(step.c_s, step.c_s_d, step.c_s_dd, step.c_d, step.c_d_d, step.c_d_dd, step.d_list,
                                        step.t_list, step.v_list, step.dt, step.csp, step.v_thr)
    ref = vec.score_dense(step)
    live = np.flatnonzero(ref["level"] != 255)
    cost = ref["cost"][12][live]
    lv = ref["level"][live]
    top = np.flatnonzero(lv == (10 if (lv == 10).any() else lv.max()))
    a = int(top[np.argmin(cost[top])])
    b = int([c for c in top if c != a][0])
    ctx = _ctx("fp32")
    try:
        for sign in (+1, -1):
            factor = np.full(len(fps), 1e3)
            factor[a] = 1.0
            factor[b] = cost[a] * (1.0 + sign * 1e-10) / cost[b]
            r = dropin.score_trajectories(fps, step.predictions, step.obstacle_types, step.vehicle, step.params, step.mode, ctx=ctx,
                                          velocity_cost_mode=step.velocity_cost_mode, velocity_target=step.velocity_target,
                                          cost_factor=factor)
            assert r.best["index"] == (a if sign > 0 else b), (sign, r.best, a, b)
        # the winner twice: the first of the two identical trajectories is selected
        dup = list(fps) + [fps[a]]
        r = dropin.score_trajectories(dup, step.predictions, step.obstacle_types, step.vehicle, step.params, step.mode, ctx=ctx,
                                      velocity_cost_mode=step.velocity_cost_mode, velocity_target=step.velocity_target)
        assert r.best["index"] == a and r.cost[12][a] == r.cost[12][len(fps)]
        assert ctx.debug_counters()["bound_violations"] == 0
    finally:
        ctx.close()
