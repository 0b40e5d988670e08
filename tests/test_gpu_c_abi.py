"""The C ABI used from plain C (tests/c_abi/abi_demo.c: gcc, include/etp_b200.h, libetp_b200.so -- no Python, no torch, no
CUDA headers): same selection as the Python layer on the same planning step."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_step(step, path):
    from ethicaltrajectoryplanning_b200 import _abi
    from ethicaltrajectoryplanning_b200.scoring import make_params, pack_predictions
    from ethicaltrajectoryplanning_b200.spline import CubicSpline2D

    p = make_params(step.params, step.vehicle, step.mode)
    nums = list(p.weights) + [p.planner_mode, p.fast_prob_mahalanobis, p.multiple_cost_functions, p.n_angle_thresholds,
                              p.max_acceptable_risk] + list(p.angle_thr) + list(p.angle_coef_pos) + list(p.angle_coef_neg)
    nums += [p.lr_const, p.lr_speed, p.lr1s_const, p.lr1s_speed, p.ped_const, p.ped_speed, p.veh_l, p.veh_w, p.veh_l_r, p.veh_m,
             p.veh_steer_max, p.veh_lat_a_max, p.veh_v_max, p.veh_a_max, p.collision_check]
    csp = CubicSpline2D.from_foreign(step.csp)
    nums += [len(csp.s) - 1] + list(np.ravel(csp.s)) + list(np.ravel(csp.coef_x)) + list(np.ravel(csp.coef_y))
    pk = pack_predictions(step.predictions, step.obstacle_types)
    nums += [pk["O"], pk["Tp"]]
    for k in ("pred_len", "pos", "cov", "yaw", "vel", "length", "width", "mass", "prot", "resp"):
        nums += list(np.ravel(pk[k]))
    nums += [step.c_s, step.c_s_d, step.c_s_dd, step.c_d, step.c_d_d, step.c_d_dd]
    nums += [len(step.v_list)] + list(step.v_list) + [len(step.t_list)] + list(step.t_list) + list(step.n_samples)
    nums += [len(step.d_list)] + list(step.d_list) + [step.dt, step.v_thr, step.velocity_cost_mode, step.velocity_target, step.cost_factor]
    with open(path, "w") as fh:
        fh.write("\n".join(repr(float(v)) for v in nums))
    assert _abi.NUM_WEIGHTS == 14


def test_plain_c_program_selects_the_same_trajectory(tmp_path):
    from ethicaltrajectoryplanning_b200 import _abi
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    _abi.load_library()
    libdir = os.path.dirname(_abi.LIB_PATH)
    exe = os.path.join(str(tmp_path), "abi_demo")
    subprocess.run(["gcc", "-std=c99", "-O1", os.path.join(ROOT, "tests", "c_abi", "abi_demo.c"), "-I", os.path.join(ROOT, "include"),
                    "-L", libdir, "-letp_b200", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    ctx = ScoringContext(0, "fp32")
    try:
        for kw in (dict(name="planning", seed=3, n_obstacles=5), dict(name="cfg2")):
            step = make_step(**kw)
            path = os.path.join(str(tmp_path), "step.txt")
            _write_step(step, path)
            res = subprocess.run([exe, path], capture_output=True, text=True)
            assert res.returncode == 0, res.stderr
            lines = res.stdout.strip().split("\n")
            tag, index, list_index, level, n_scored, cost = lines[0].split()
            want = ctx.plan(step, materialize=False)
            assert tag == "best" and int(index) == want.best["index"] and int(list_index) == want.best["list_index"]
            assert int(level) == want.best["level"] and int(n_scored) == want.best["n_scored"]
            assert float(cost) == want.best["cost"]
            traj = want.best_trajectory()
            got = [float(v) for v in lines[1].split()[2:]]
            np.testing.assert_array_equal(got[0::2], traj["x_m"][:3])
            np.testing.assert_array_equal(got[1::2], traj["y_m"][:3])
            assert "sm_100a" in lines[2]
            cyc = lines[3].split()
            assert cyc[0] == "cycle" and [int(v) for v in cyc[1:5]] == [want.best[k] for k in ("index", "list_index", "level", "n_scored")]
            assert float(cyc[5]) == want.best["cost"] and int(cyc[6]) == len(traj["x_m"])
            got = [float(v) for v in cyc[7:]]
            np.testing.assert_array_equal(got[0::2], traj["x_m"][:3])
            np.testing.assert_array_equal(got[1::2], traj["y_m"][:3])
    finally:
        ctx.close()
