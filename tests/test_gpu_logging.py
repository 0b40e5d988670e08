"""Columnar side-car log: the legacy text line rebuilt from the binary record equals, character by
character, the line FrenetLogging.log_data writes for the object path of the same planning step."""
import os
import time
from types import SimpleNamespace

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _object_path_line(step, tmp_path, name):
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils import frenet_functions as ff
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils.logging import FrenetLogging

    ego = SimpleNamespace(position=np.asarray(step.ego_position), orientation=step.ego_orientation, time_step=0)
    scenario = SimpleNamespace(dt=step.dt, obstacle_by_id=lambda oid, _t=step.obstacle_types: SimpleNamespace(
        obstacle_type=SimpleNamespace(name=_t[oid])))
    problem = SimpleNamespace(velocity_cost_mode=step.velocity_cost_mode, velocity_target=step.velocity_target,
                              cost_factor=step.cost_factor)
    t0 = time.perf_counter()
    fp_list = ff.calc_frenet_trajectories(step.c_s, step.c_s_d, step.c_s_dd, step.c_d, step.c_d_d, step.c_d_dd, step.d_list,
                                          step.t_list, step.v_list, step.dt, step.csp, step.v_thr)
    valid, invalid, _ = ff.sort_frenet_trajectories(ego, fp_list, None, step.predictions, step.mode, step.params, problem,
                                                    scenario, step.vehicle, 1, step.dt, 30.0, None, None, None)
    valid.sort(key=lambda fp: fp.cost, reverse=False)
    path = os.path.join(tmp_path, name)
    log = FrenetLogging(path)
    log.log_data(3, 3 * step.dt, [d.__dict__ for d in valid], [d.__dict__ for d in invalid], step.predictions, 0, None)
    dt = time.perf_counter() - t0
    return open(path).read().split("\n")[1], dt, os.path.getsize(path)


def _columnar(step, tmp_path, name):
    from ethicaltrajectoryplanning_b200 import dropin
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils.logging import ColumnarLogging, legacy_line, read_columnar

    ctx = dropin.default_context()
    path = os.path.join(tmp_path, name)
    log = ColumnarLogging(path)
    t0 = time.perf_counter()
    res = ctx.plan(step)
    n = log.log_step(3, 3 * step.dt, res, step, step.predictions)
    dt = time.perf_counter() - t0
    records = read_columnar(path)
    assert len(records) == 1 and os.path.getsize(path) == n
    return legacy_line(*records[0]), dt, n, path


@pytest.mark.parametrize("kw", [dict(name="planning", seed=3, n_obstacles=4, nv=5, nd=7),
                                dict(name="planning", seed=4, n_obstacles=3, nv=4, nd=9, t_list=(2.0, 3.0), v0=2.5),
                                dict(name="planning", seed=5, n_obstacles=0, nv=4, nd=5)])
def test_legacy_line_from_the_side_car_is_identical(tmp_path, kw):
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    step = make_step(**kw)
    want, _, _ = _object_path_line(step, str(tmp_path), "text.log")
    got, _, _, _ = _columnar(step, str(tmp_path), "side.etpl")
    assert got == want


def test_with_host_masks_and_several_records(tmp_path):
    from ethicaltrajectoryplanning_b200 import dropin
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils.logging import ColumnarLogging, legacy_line, read_columnar
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    path = os.path.join(str(tmp_path), "multi.etpl")
    log = ColumnarLogging(path)
    ctx = dropin.default_context()
    lines = []
    for k in range(3):
        step = make_step("planning", seed=10 + k, n_obstacles=2 + k, nv=4, nd=5)
        step.params["weights"]["lon_jerk"] = 0.2
        res = ctx.plan(step)
        log.log_step(k, k * step.dt, res, step, step.predictions)
        lines.append((step, res))
    records = read_columnar(path)
    assert [h["simulation_step"] for h, _ in records] == [0, 1, 2]
    for (h, a), (step, res) in zip(records, lines):
        line = legacy_line(h, a)
        assert line.count(";") >= 6
        fields = line.split(";")
        assert int(fields[0]) == h["simulation_step"]
        import json
        valid = json.loads(fields[2])
        assert valid and list(valid[0].keys())[:14] == ["t", "d", "d_d", "d_dd", "d_ddd", "s", "s_d", "s_dd", "s_ddd", "x", "y", "yaw", "v", "curv"]
        costs = [v["cost"] for v in valid]
        assert costs == sorted(costs)


def test_sizes_and_times_of_a_full_step(tmp_path, capsys):
    """10^4 candidates: report what the side-car saves (not asserted beyond sanity)."""
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    step = make_step("cfg2")
    _, t_text, b_text = _object_path_line(step, str(tmp_path), "text_cfg2.log")
    _, t_col, b_col, _ = _columnar(step, str(tmp_path), "side_cfg2.etpl")
    with capsys.disabled():
        print(f"\ncfg2 step log: text line {b_text / 1e6:.1f} MB in {t_text:.2f} s (objects + json), "
              f"side-car {b_col / 1e6:.1f} MB in {t_col * 1e3:.1f} ms")
    assert b_col < b_text / 3 and t_col < t_text


def test_closed_loop_with_side_car(tmp_path):
    """FrenetPlanner's fused path with a columnar log: one record per cycle, winner equal to the un-logged run."""
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils.logging import legacy_line, read_columnar
    from ethicaltrajectoryplanning_b200.workloads import closed_loop

    path = os.path.join(str(tmp_path), "loop.etpl")
    _, logged = closed_loop(n_steps=5, grid="planning", n_obstacles=4, columnar_log_path=path)
    _, plain = closed_loop(n_steps=5, grid="planning", n_obstacles=4)
    assert [s[3] for s in logged] == [s[3] for s in plain]
    records = read_columnar(path)
    assert len(records) == 5
    import json
    first_valid = json.loads(legacy_line(*records[2]).split(";")[2])[0]
    assert abs(first_valid["cost"] - min(v["cost"] for v in json.loads(legacy_line(*records[2]).split(";")[2]))) == 0.0
