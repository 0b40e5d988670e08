"""Road-boundary test on the device (etp_set_road_boundary; SURVEY.md 8f N1, second half) against oracle/boxes.py:
first colliding ego box, boundary harm LR1S(v[i]) in the risk costs, validity level 2 -- bit-exact levels, costs within
tolerance, with and without predictions, discrete and swept boxes, generated and caller-provided trajectories."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _step(mode, **kw):
    from ethicaltrajectoryplanning_b200.synthetic import make_step, road_boundary_triangles

    step = make_step(**kw)
    step.road_boundary = (road_boundary_triangles(step.csp, half_width=kw.pop("half_width", 3.6), s_end=140.0), mode)
    return step


def _compare(step, precisions=("fp32", "fp64")):
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext
    from oracle import vec

    ref = vec.score_dense(step)
    for precision, rtol in (("fp32", 2e-4), ("fp64", 1e-9)):
        if precision not in precisions:
            continue
        ctx = ScoringContext(0, precision)
        try:
            res = ctx.plan(step)
            out = res.numpy()
            np.testing.assert_array_equal(out["level"].astype(np.int64), ref["level"], err_msg=precision)
            np.testing.assert_array_equal(out["reason"].astype(np.int64), ref["reason"], err_msg=precision)
            live = ref["level"] != vec.LEVEL_SKIPPED
            np.testing.assert_allclose(out["cost"][:, live], ref["cost"][:, live], rtol=rtol, atol=1e-6, err_msg=precision)
            assert res.best["level"] == ref["best"]["level"]
            if res.best["index"] != ref["best"]["index"]:
                assert precision == "fp32"
                d = ref["dense_index"].tolist()
                assert abs(ref["cost"][12][d.index(res.best["index"])] - ref["best"]["cost"]) <= 1e-4 * abs(ref["best"]["cost"]) + 1e-7
        finally:
            ctx.close()
    return ref


@pytest.mark.parametrize("mode", ["discrete", "swept"])
def test_planning_grid_with_road_boundary(mode):
    from oracle import vec

    ref = _compare(_step(mode, name="planning", seed=5, n_obstacles=4, nv=9, nd=21))
    lvl = ref["level"][ref["level"] != vec.LEVEL_SKIPPED]
    assert 0 < np.sum(lvl == 2) < len(lvl), np.bincount(lvl, minlength=11)


@pytest.mark.parametrize("mode", ["discrete", "swept"])
def test_cfg2_with_road_boundary(mode):
    ref = _compare(_step(mode, name="cfg2", lateral_spread=10.0))
    n2 = int(np.sum(ref["level"] == 2))
    assert 500 < n2 < 9500, n2


def test_without_predictions_any_contact_is_level_two():
    step = _step("swept", name="planning", seed=6, n_obstacles=3, nv=7, nd=15)
    step.predictions = None
    ref = _compare(step)
    assert np.any(ref["level"] == 2) and np.any(ref["level"] == 10)


def test_boundary_is_persistent_and_removable():
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext

    step = _step("discrete", name="planning", seed=7, n_obstacles=3, nv=7, nd=15)
    ctx = ScoringContext(0, "fp32")
    try:
        with_rb = ctx.plan(step).numpy()["level"].copy()
        again = ctx.plan(step).numpy()["level"].copy()
        np.testing.assert_array_equal(with_rb, again)
        step2 = copy.copy(step)
        step2.road_boundary = None
        without = ctx.plan(step2).numpy()["level"]
        assert np.any(with_rb == 2) and not np.any(without == 2)
    finally:
        ctx.close()


def test_given_trajectories_with_road_boundary():
    from ethicaltrajectoryplanning_b200 import dropin
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext
    from oracle import port

    step = _step("swept", name="planning", seed=8, n_obstacles=3, nv=7, nd=15, t_list=(2.0, 3.0))
    _, _, _, fps = port.plan_step(step)
    want = np.array([fp.valid_level for fp in fps])
    assert 0 < np.sum(want == 2) < len(fps)
    for precision in ("fp32", "fp64"):
        ctx = ScoringContext(0, precision)
        try:
            ctx.set_road_boundary(*step.road_boundary)
            res = dropin.score_trajectories(fps, step.predictions, step.obstacle_types, step.vehicle, step.params, step.mode,
                                            velocity_cost_mode=step.velocity_cost_mode, velocity_target=step.velocity_target,
                                            cost_factor=step.cost_factor, ctx=ctx)
            np.testing.assert_array_equal(np.asarray(res.level).astype(np.int64), want, err_msg=precision)
        finally:
            ctx.close()


def test_drop_in_sort_with_a_road_boundary_object():
    """sort_frenet_trajectories(..., road_boundary=RoadBoundary(...)): levels and bd_harm like the loop oracle."""
    from types import SimpleNamespace

    from ethicaltrajectoryplanning_b200.planner.Frenet.utils import frenet_functions as ff
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils.validity_checks import RoadBoundary
    from oracle import port

    step = _step("swept", name="planning", seed=9, n_obstacles=3, nv=7, nd=15)
    _, _, _, fps = port.plan_step(step)
    want = {fp.dense_index: fp.valid_level for fp in fps}
    ego = SimpleNamespace(position=np.asarray(step.ego_position), orientation=step.ego_orientation, time_step=0)
    scenario = SimpleNamespace(dt=step.dt, obstacle_by_id=lambda oid, _t=step.obstacle_types: SimpleNamespace(
        obstacle_type=SimpleNamespace(name=_t[oid])))
    problem = SimpleNamespace(velocity_cost_mode=step.velocity_cost_mode, velocity_target=step.velocity_target, cost_factor=1.0)
    fp_list = ff.calc_frenet_trajectories(step.c_s, step.c_s_d, step.c_s_dd, step.c_d, step.c_d_d, step.c_d_dd, step.d_list,
                                          step.t_list, step.v_list, step.dt, step.csp, step.v_thr)
    valid, invalid, vd = ff.sort_frenet_trajectories(
        ego, fp_list, None, step.predictions, step.mode, step.params, problem, scenario, step.vehicle, 1, step.dt, 30.0,
        RoadBoundary(*step.road_boundary), None, None)
    got = [fp.valid_level for fp in fp_list]
    assert got == [want[i] for i in sorted(want)]
    assert 0 < len(vd[2]) < len(fp_list)


def test_against_the_reference_written_fixtures():
    """boundary_cases.npz: the reference's own sort_frenet_trajectories with a road boundary (stand-in pycrcc primitives)
    decided boundary harm, levels, costs and the winner."""
    import importlib.util
    import os

    from ethicaltrajectoryplanning_b200 import _abi
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext

    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("_mgb", os.path.join(here, "golden", "make_golden_boundary.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    z = np.load(os.path.join(here, "golden", "boundary_cases.npz"))
    for name, step, _ in mod.boundary_cases():
        for precision, rtol in (("fp32", 2e-4), ("fp64", 1e-9)):
            ctx = ScoringContext(0, precision)
            try:
                res = ctx.plan(step)
                out = res.numpy()
                keep = out["level"] != _abi.LEVEL_SKIPPED
                np.testing.assert_array_equal(out["level"][keep].astype(np.int64), z[f"{name}_level"], err_msg=f"{name} {precision}")
                np.testing.assert_array_equal(out["reason"][keep].astype(np.int64), z[f"{name}_reason"], err_msg=f"{name} {precision}")
                if step.predictions is not None:
                    np.testing.assert_allclose(out["bd_harm"][keep], z[f"{name}_bd"], rtol=1e-12, atol=0)
                top = z[f"{name}_level"] == z[f"{name}_level"].max()
                np.testing.assert_allclose(out["cost"][12, keep][top], z[f"{name}_cost"][top], rtol=rtol, atol=1e-6)
                assert res.best["list_index"] == int(z[f"{name}_best"])
            finally:
                ctx.close()
