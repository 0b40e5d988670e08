"""Host-side logic that needs no GPU: configs, spline tables, synthetic steps, ABI surface, packing."""
import ctypes
import os
import re

import numpy as np
import pytest

from ethicaltrajectoryplanning_b200 import _abi, configs, synthetic
from ethicaltrajectoryplanning_b200.scoring import harm_model_tables, make_params, pack_predictions
from ethicaltrajectoryplanning_b200.spline import CubicSpline2D
from ethicaltrajectoryplanning_b200.vehicleparams import VehicleParameters

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    """include/etp_b200.h <-> libetp_b200.so <-> ctypes table (no compute calls: there is no GPU here)."""
    _abi.build_library()
    lib = _abi.load_library()
    header = open(os.path.join(ROOT, "include", "etp_b200.h")).read()
    declared = set(re.findall(r"\b(etp_[a-z_]+)\s*\(", header))
    assert declared == set(_abi.SYMBOLS), declared ^ set(_abi.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.etp_version()


def test_abi_struct_sizes_match_the_header_layout():
    assert ctypes.sizeof(_abi.Best) == 40
    assert ctypes.sizeof(_abi.Params) == 8 * 14 + 4 * 4 + 8 + 8 * 8 + 8 * 9 * 2 + 8 * 6 + 8 * 8
    assert ctypes.sizeof(_abi.Out) == 8 + 6 * 8 + 8


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_abi, "_LIB", None)
    monkeypatch.setattr(_abi, "LIB_PATH", "/nonexistent/libetp_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _abi.load_library()


def test_config_loaders_and_weight_files():
    w = configs.load_weight_json("weights_ethical.json")
    assert (w["bayes"], w["equality"], w["maximin"], w["risk_cost"], w["responsibility"]) == (33.3, 33.3, 33.3, 250, 0.0)
    assert configs.load_weight_json("weights_ego.json")["ego"] == 100.0
    assert configs.load_weight_json("weights_standard.json")["risk_cost"] == 1.0
    assert configs.load_weight_json()["responsibility"] == 0.5
    p = configs.load_planning_json("planning_fast.json")["frenet_settings"]["frenet_parameters"]
    np.testing.assert_allclose(p["d_list"], np.linspace(-2.5, 2.5, 5))
    assert configs.load_risk_json()["harm_mode"] == "log_reg"
    assert configs.load_harm_parameter_json()["log_reg"]["reduced_sym_angle_areas"]["const"] == -4.457
    assert tuple(configs.WEIGHT_KEYS)[:6] == configs.RISK_COSTS_KEY


def test_config_dir_override(tmp_path):
    import json

    (tmp_path / "weights.json").write_text(json.dumps({k: 1.5 for k in configs.WEIGHT_KEYS}))
    assert configs.load_weight_json(config_dir=str(tmp_path))["bayes"] == 1.5


def test_spline_interpolates_and_is_smooth():
    x = np.linspace(0, 100, 21)
    y = 2 * np.sin(x / 15)
    csp = CubicSpline2D(x, y)
    px, py = csp.calc_position(csp.s)
    np.testing.assert_allclose(px, x, atol=1e-12)
    np.testing.assert_allclose(py, y, atol=1e-12)
    # C1 at interior knots
    eps = 1e-6
    for k in (3, 10, 17):
        a = np.array(csp.calc_position(csp.s[k] - eps))
        b = np.array(csp.calc_position(csp.s[k] + eps))
        c = np.array(csp.calc_position(csp.s[k]))
        np.testing.assert_allclose((c - a) / eps, (b - c) / eps, atol=1e-4)
    # clamped extrapolation uses the end segments
    assert np.isfinite(csp.calc_position(-5.0)[0]) and np.isfinite(csp.calc_position(csp.s[-1] + 5.0)[0])
    again = CubicSpline2D.from_foreign(csp)
    assert again is csp


def test_get_v_list_quirk_q8():
    v = synthetic.get_v_list(1.0, 20.0, 7.5, n_samples=5)
    np.testing.assert_allclose(v, [1.0, 1.0 + 19 / 3, 1.0 + 38 / 3, 20.0, 7.5])
    with pytest.raises(ValueError):
        synthetic.get_v_list(1.0, 2.0, 1.5, n_samples=0)
    with pytest.raises(AttributeError):
        synthetic.get_v_list(1.0, 2.0, 1.5, v_goal_min=1.0)


def test_synthetic_steps_have_the_named_shapes():
    s = synthetic.make_step("cfg2")
    assert (len(s.v_list), len(s.d_list), int(s.n_samples[0]), len(s.predictions)) == (100, 100, 30, 8)
    s3 = synthetic.make_step("cfg3", nv=4, nd=4)
    assert int(s3.n_samples[0]) == 50 and len(s3.predictions) == 32
    kinds = list(s.obstacle_types.values())
    assert kinds.count("PEDESTRIAN") == 2 and "PARKED_VEHICLE" in kinds
    assert not np.any(s.predictions[101]["cov_list"])               # Q14 all-zero covariance
    assert np.allclose(s.predictions[102]["cov_list"][:, 0, 0], 0.02)  # static obstacle


def test_harm_model_tables_cover_every_log_reg_variant():
    coeffs = configs.load_harm_parameter_json()
    modes = configs.load_risk_json()
    thr, pos, neg, const, speed = harm_model_tables(modes, coeffs)  # default LR4S
    assert len(thr) == 2 and pos == neg == [0.0, 0.244, -0.431] and (const, speed) == (-4.457, 0.177)
    m = dict(modes, sym_angle=False)
    thr, pos, neg, _, _ = harm_model_tables(m, coeffs)
    assert pos[1] == 0.250 and neg[1] == 0.259
    m = dict(modes, reduced_angle_areas=False)
    assert len(harm_model_tables(m, coeffs)[0]) == 6
    m = dict(modes, ignore_angle=True)
    assert harm_model_tables(m, coeffs)[0] == []
    with pytest.raises(NotImplementedError):
        harm_model_tables(dict(modes, harm_mode="gidas"), coeffs)
    with pytest.raises(ValueError):
        harm_model_tables(dict(modes, harm_mode="nope"), coeffs)
    p = make_params(configs.default_params(), VehicleParameters("bmw_320i"), "risk")
    assert p.n_angle_thresholds == 2 and p.veh_v_max == 50.8 and p.weights[5] == 250
    with pytest.raises(ValueError):
        make_params(configs.default_params(), VehicleParameters("bmw_320i"), "bogus")


def test_pack_predictions_layout_and_ragged_lengths():
    s = synthetic.make_step("cfg1", n_pred=9)
    first = next(iter(s.predictions))
    s.predictions[first]["pos_list"] = s.predictions[first]["pos_list"][:5]
    s.predictions[first]["cov_list"] = s.predictions[first]["cov_list"][:5]
    pk = pack_predictions(s.predictions, s.obstacle_types)
    assert pk["O"] == 5 and pk["Tp"] == 9 and pk["pred_len"][0] == 5
    assert pk["pos"].shape == (5, 9, 2) and pk["cov"].shape == (5, 9, 4)
    assert np.all(pk["pos"][0, 5:] == 0)
    assert pk["prot"].tolist() == [1, 1, 1, 0, 1]
    assert pk["mass"][3] == 75.0 and pk["mass"][0] == pytest.approx(-1333.5 + 526.9 * (5.5 * 2.3) ** 0.8)
    with pytest.raises(ValueError):
        pack_predictions(s.predictions, {k: "BUILDING" for k in s.predictions})
