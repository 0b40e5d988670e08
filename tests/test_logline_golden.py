"""The per-step log line against lines the reference's own FrenetLogging.log_data wrote (tests/golden/logline_*.txt.gz,
tests/golden/make_golden_logline.py; planner/Frenet/utils/logging.py:31-61,166-182, frenet_planner.py:469-477).

CPU: this repository's writer re-emits the reference's line byte for byte from the parsed values (same separators, key order,
number formatting).  GPU: the line of the drop-in object path (calc_frenet_trajectories -> sort_frenet_trajectories ->
FrenetLogging) and the line rebuilt from the columnar side-car have the reference line's exact skeleton (every character
that is not part of a float literal) and its values within 1e-9 (fp64 check mode) / 1e-4 (fp32) after parsing.
"""
import gzip
import json
import os
import re
from types import SimpleNamespace

import numpy as np
import pytest

from helpers import load_golden

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("cfg1_ethical", "max_risk_level3", "multi_t", "no_obstacles")
FLOAT = re.compile(r"-?(?:\d+\.\d*(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+|NaN|Infinity)")


def _golden_text(name):
    with gzip.open(os.path.join(GOLDEN, f"logline_{name}.txt.gz"), "rb") as fh:
        return fh.read().decode("utf-8")


@pytest.mark.parametrize("name", CASES)
def test_writer_reproduces_the_reference_line_byte_for_byte(name, tmp_path):
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils.logging import FrenetLogging, get_data_from_line

    text = _golden_text(name)
    header, line = text.split("\n")
    cells = [json.loads(c) for c in line.split(";")]
    path = os.path.join(tmp_path, "sub", "frenet.csv")
    FrenetLogging(path).log_data(*cells)
    assert open(path).read() == text
    got = get_data_from_line(path, 1)
    assert list(got.keys()) == header.split(";") and got["simulation_step"] == 3


TRAJ_KEYS = ("t", "d", "d_d", "d_dd", "d_ddd", "s", "s_d", "s_dd", "s_ddd", "x", "y", "yaw", "v", "curv")


def _close(a, b, rtol, atol, where):
    if isinstance(b, dict):
        assert isinstance(a, dict) and list(a.keys()) == list(b.keys()), where
        for k in b:
            _close(a[k], b[k], rtol, atol, where + "/" + str(k))
    elif isinstance(b, list):
        assert isinstance(a, list) and len(a) == len(b), where
        if b and all(isinstance(v, (int, float)) for v in b):
            x, y = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
            key = where.rsplit("/", 1)[-1]
            if key in TRAJ_KEYS:
                # the bar of tests/test_gpu_golden.py: relative to max(|value|, 1e-3); the curvature of a profile that creeps
                # at centimetres per second is a third difference quotient of map coordinates
                tt = 1e-4 if key == "curv" else (1e-6 if rtol < 1e-6 else 2e-5)
                assert np.all(np.abs(x - y) <= tt * np.maximum(np.abs(y), 1e-3)), (where, float(np.abs(x - y).max()))
            else:
                assert np.all(np.abs(x - y) <= rtol * np.abs(y) + atol), (where, float(np.abs(x - y).max()))
        else:
            for i, (x, y) in enumerate(zip(a, b)):
                _close(x, y, rtol, atol, where + f"[{i}]")
    elif isinstance(b, float):
        assert isinstance(a, (int, float)) and abs(a - b) <= rtol * abs(b) + atol, (where, a, b)
    else:
        assert a == b and type(a) is type(b), (where, a, b)


def _compare_lines(got, want, rtol, atol):
    # integers that stand for floats (a cost of exactly 0 prints "0" when it was never assigned, "0.0" when computed) are
    # part of the skeleton on purpose: the reference leaves `cost = 0` on candidates below the top level
    assert FLOAT.sub("#", got) == FLOAT.sub("#", want)
    for i, (a, b) in enumerate(zip(got.split(";"), want.split(";"))):
        _close(json.loads(a), json.loads(b), rtol, atol, f"cell{i}")


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("name", CASES)
def test_object_path_and_side_car_lines_equal_the_reference_line(name, precision, tmp_path, monkeypatch):
    from ethicaltrajectoryplanning_b200 import dropin
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils import frenet_functions as ff
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils.logging import ColumnarLogging, FrenetLogging, legacy_line, read_columnar

    step, ref = load_golden(name)
    want = _golden_text(name).split("\n")[1]
    rtol, atol = (1e-9, 1e-15) if precision == "fp64" else (1e-4, 1e-9)  # north star: probability / harm / cost bars
    monkeypatch.setenv("ETP_B200_PRECISION", precision)  # the drop-in entry points score on dropin.default_context()
    ctx = dropin.default_context()
    if True:
        ego = SimpleNamespace(position=np.asarray(step.ego_position), orientation=step.ego_orientation, time_step=0)
        scenario = SimpleNamespace(dt=step.dt, obstacle_by_id=lambda oid, _t=step.obstacle_types: SimpleNamespace(
            obstacle_type=SimpleNamespace(name=_t[oid])))
        problem = SimpleNamespace(velocity_cost_mode=step.velocity_cost_mode, velocity_target=step.velocity_target,
                                  cost_factor=step.cost_factor)
        fp_list = ff.calc_frenet_trajectories(step.c_s, step.c_s_d, step.c_s_dd, step.c_d, step.c_d_d, step.c_d_dd,
                                              step.d_list, step.t_list, step.v_list, step.dt, step.csp, step.v_thr)
        valid, invalid, _ = ff.sort_frenet_trajectories(ego, fp_list, None, step.predictions, step.mode, step.params, problem,
                                                        scenario, step.vehicle, 1, step.dt, 30.0, None, None, None)
        valid.sort(key=lambda fp: fp.cost, reverse=False)
        path = os.path.join(tmp_path, "text.log")
        FrenetLogging(path).log_data(3, 3 * step.dt, [d.__dict__ for d in valid], [d.__dict__ for d in invalid],
                                     step.predictions, 0, None)
        _compare_lines(open(path).read().split("\n")[1], want, rtol, atol)
        # the same line rebuilt from the binary side-car record
        side = os.path.join(tmp_path, "side.etpl")
        res = ctx.plan(step)
        ColumnarLogging(side).log_step(3, 3 * step.dt, res, step, step.predictions)
        _compare_lines(legacy_line(*read_columnar(side)[0]), want, rtol, atol)
