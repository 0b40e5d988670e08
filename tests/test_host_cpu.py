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
    assert ctypes.sizeof(_abi.Params) == 8 * 14 + 4 * 4 + 8 + 8 * 8 + 8 * 9 * 2 + 8 * 6 + 8 * 8 + 8
    assert ctypes.sizeof(_abi.Out) == 8 + 6 * 8 + 8 + 8 + 8
    # sizeof(etp_goal), sizeof(etp_step), sizeof(etp_trajectories) as g++ lays them out (x86-64)
    assert (ctypes.sizeof(_abi.Goal), ctypes.sizeof(_abi.Step), ctypes.sizeof(_abi.Trajectories)) == (176, 192, 88)


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


# ---------------------------------------------------------------------------------------------------------------
# call surface of the reference: same module paths, names and leading parameters
# ---------------------------------------------------------------------------------------------------------------
REFERENCE_SURFACE = {
    # module (under ethicaltrajectoryplanning_b200 / EthicalTrajectoryPlanning): {function: parameters of the reference}
    "planner.Frenet.utils.frenet_functions": {
        # frenet_functions.py:207-221
        "calc_frenet_trajectories": ["c_s", "c_s_d", "c_s_dd", "c_d", "c_d_d", "c_d_dd", "d_list", "t_list", "v_list", "dt",
                                     "csp", "v_thr", "exec_timer"],
        # frenet_functions.py:477-495
        "sort_frenet_trajectories": ["ego_state", "fp_list", "global_path", "predictions", "mode", "params", "planning_problem",
                                     "scenario", "vehicle_params", "ego_id", "dt", "sensor_radius", "road_boundary",
                                     "collision_checker", "goal_area", "exec_timer", "reach_set"],
        # frenet_functions.py:760-768
        "get_v_list": ["v_min", "v_max", "v_cur", "v_goal_min", "v_goal_max", "n_samples", "mode"],
    },
    "planner.Frenet.utils.validity_checks": {
        # validity_checks.py:31-42
        "check_validity": ["ft", "ego_state", "scenario", "vehicle_params", "risk_params", "mode", "road_boundary",
                           "collision_checker", "predictions", "exec_timer"],
    },
    "planner.Frenet.utils.calc_trajectory_cost": {
        # calc_trajectory_cost.py:33-50
        "calc_trajectory_costs": ["traj", "global_path", "planning_problem", "ego_state", "validity_level", "scenario", "params",
                                  "ego_id", "dt", "predictions", "goal_area", "vehicle_params", "sensor_radius", "exec_timer",
                                  "mode", "reach_set"],
    },
    "risk_assessment.risk_costs": {
        # risk_costs.py:16-27, 128, 153, 181, 211, 229
        "calc_risk": ["traj", "ego_state", "predictions", "scenario", "ego_id", "vehicle_params", "params", "road_boundary",
                      "reach_set", "exec_timer"],
        "get_bayesian_costs": ["ego_risk_max", "obst_risk_max", "boundary_harm"],
        "get_equality_costs": ["ego_risk_max", "obst_risk_max"],
        "get_maximin_costs": ["ego_risk_max", "obst_risk_max", "ego_harm_max", "obst_harm_max", "boundary_harm", "eps",
                              "scale_factor"],
        "get_ego_costs": ["ego_risk_max", "boundary_harm"],
        "get_responsibility_cost": ["scenario", "traj", "ego_state", "obst_risk_max", "predictions", "reach_set", "mode"],
    },
    "risk_assessment.collision_probability": {
        # collision_probability.py:136, 259
        "get_collision_probability_fast": ["traj", "predictions", "vehicle_params", "safety_margin"],
        "get_inv_mahalanobis_dist": ["traj", "predictions", "vehicle_params", "safety_margin"],
    },
    "risk_assessment.harm_estimation": {
        # harm_estimation.py:202
        "get_harm": ["scenario", "traj", "predictions", "ego_id", "vehicle_params", "modes", "coeffs", "timer"],
    },
}


def test_dropin_modules_keep_the_reference_call_surface():
    import importlib
    import inspect

    for mod, fns in REFERENCE_SURFACE.items():
        m = importlib.import_module("ethicaltrajectoryplanning_b200." + mod)
        for name, params in fns.items():
            got = list(inspect.signature(getattr(m, name)).parameters)
            assert got[:len(params)] == params, (mod, name, got)
    # the same check against the reference sources when they are mounted (build container only)
    ref_root = "/root/reference"
    if os.path.isdir(ref_root):
        import ast

        for mod, fns in REFERENCE_SURFACE.items():
            tree = ast.parse(open(os.path.join(ref_root, *mod.split(".")) + ".py").read())
            defs = {n.name: [a.arg for a in n.args.args] for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)}
            for name, params in fns.items():
                assert defs[name] == params, (mod, name, defs[name])


def test_frenet_planner_and_trajectory_keep_the_reference_attributes():
    import inspect

    from ethicaltrajectoryplanning_b200.planner.Frenet.frenet_planner import FrenetPlanner
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils.frenet_functions import FrenetTrajectory

    # frenet_planner.py:79-92
    want = ["self", "scenario", "planning_problem", "ego_id", "vehicle_params", "mode", "exec_timer", "frenet_parameters",
            "sensor_radius", "plot_frenet_trajectories", "weights", "settings"]
    assert list(inspect.signature(FrenetPlanner.__init__).parameters)[:len(want)] == want
    assert hasattr(FrenetPlanner, "_step_planner") and hasattr(FrenetPlanner, "step")
    # attribute insertion order = key order of the log line (frenet_functions.py:72-102)
    ft = FrenetTrajectory()
    assert list(ft.__dict__) == ["t", "d", "d_d", "d_dd", "d_ddd", "s", "s_d", "s_dd", "s_ddd", "x", "y", "yaw", "v", "curv",
                                 "valid_level", "reason_invalid", "cost", "ego_risk_dict", "obst_risk_dict"]


def test_log_line_format_round_trip(tmp_path):
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils.logging import FrenetLogging, get_data_from_line

    p = os.path.join(tmp_path, "a", "log.csv")
    lg = FrenetLogging(p)
    lg.log_data(3, 0.3, [{"t": np.arange(2.0), "cost": 1.5, "ego_risk_dict": {7: 0.25}}], [], {7: {"pos_list": np.zeros((1, 2))}}, 0, None)
    txt = open(p).read().split("\n")
    assert txt[0] == "simulation_step;time;valid_trajectories;invalid_trajectories;predictions;calc_time_avg;reach_set"
    assert txt[1].count(";") == 6
    d = get_data_from_line(p, 1)
    assert d["simulation_step"] == 3 and d["valid_trajectories"][0]["ego_risk_dict"] == {"7": 0.25} and d["reach_set"] is None


def test_host_enrichment_matches_the_reference_golden_vectors():
    """synthetic.enrich_predictions / assign_responsibility_by_action_space (host side, used by the scene generator)
    against the vectors of the reference's own functions (tests/golden/make_golden_enrich.py)."""
    import copy

    z = np.load(os.path.join(ROOT, "tests", "golden", "enrich_cases.npz"))
    for c in range(int(z["n_cases"])):
        ids = [int(i) for i in z[f"c{c}_ids"]]
        raw, shapes, ori0, vel0 = {}, {}, {}, {}
        for j, oid in enumerate(ids):
            pos = z[f"c{c}_{j}_pos"]
            raw[oid] = {"pos_list": pos.copy(), "cov_list": np.zeros((len(pos), 2, 2))}
            a = z[f"c{c}_{j}_in"]
            shapes[oid], ori0[oid], vel0[oid] = (float(a[0]), float(a[1])), float(a[2]), float(a[3])
        pred = synthetic.enrich_predictions(copy.deepcopy(raw), shapes, ori0, vel0, float(z[f"c{c}_dt"]))
        ego = z[f"c{c}_ego"]
        synthetic.assign_responsibility_by_action_space(ego[:2], float(ego[2]), pred)
        for j, oid in enumerate(ids):
            np.testing.assert_array_equal(np.asarray(pred[oid]["orientation_list"], dtype=np.float64), z[f"c{c}_{j}_yaw"])
            np.testing.assert_array_equal(np.asarray(pred[oid]["v_list"], dtype=np.float64), z[f"c{c}_{j}_vel"])
            out = z[f"c{c}_{j}_out"]
            assert (pred[oid]["shape"]["length"], pred[oid]["shape"]["width"], pred[oid]["responsibility"]) == tuple(out)


def test_exec_timer_keeps_the_reference_surface_and_label_structure():
    """planner/utils/timers.py:5-113: same methods, same dict-of-lists result, disabled timer is a no-op."""
    import time

    from ethicaltrajectoryplanning_b200.planner.utils.timers import ExecTimer, timer_or_noop

    t = ExecTimer()
    t.start_timer("simulation/total")
    with t.time_with_cm("simulation/calculate trajectories/total"):
        time.sleep(0.001)
    with t.time_with_cm("simulation/calculate trajectories/total"):
        pass

    @t.time_with_dec("decorated")
    def f(a, b=2):
        return a + b

    assert f(1) == 3 and f.__name__ == "f"
    t.stop_timer("simulation/total")
    t.record("simulation/sort trajectories/device/fused scoring kernel", 1.5e-5)
    d = t.get_timing_dict()
    assert d is t.finished and t.running == {}
    assert len(d["simulation/calculate trajectories/total"]) == 2 and d["simulation/calculate trajectories/total"][0] >= 0.001
    assert d["simulation/total"][0] >= d["simulation/calculate trajectories/total"][0]
    assert d["simulation/sort trajectories/device/fused scoring kernel"] == [1.5e-5] and len(d["decorated"]) == 1
    t.reset()
    assert t.finished == {} and t.running == {}
    off = timer_or_noop(None)
    with off.time_with_cm("x"):
        pass
    off.start_timer("y")
    off.stop_timer("y")
    off.record("z", 1.0)
    assert off.get_timing_dict() == {}
    ref_root = "/root/reference"
    if os.path.isdir(ref_root):  # the reference class itself, when mounted: same public methods and arguments
        import ast

        tree = ast.parse(open(os.path.join(ref_root, "planner", "utils", "timers.py")).read())
        cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "ExecTimer"][0]
        import inspect

        for fn in [n for n in cls.body if isinstance(n, ast.FunctionDef)]:
            assert hasattr(ExecTimer, fn.name), fn.name
            got = list(inspect.signature(getattr(ExecTimer, fn.name)).parameters)
            assert got == [a.arg for a in fn.args.args], (fn.name, got)


def test_velocity_samples_equal_numpy_linspace_bit_for_bit():
    """get_v_list writes the linspace of frenet_functions.py:760-807 by hand (a planning cycle pays for it every 0.1 s): the
    same bits as np.append(np.linspace(...), v_cur) for every sample count, degenerate ranges and goal-velocity clamps."""
    from ethicaltrajectoryplanning_b200.synthetic import get_v_list

    rng = np.random.default_rng(7)
    for _ in range(5000):
        lo, hi = sorted(rng.uniform(0.0, 40.0, 2))
        n = int(rng.integers(2, 60))
        v_cur = float(rng.uniform(0.0, 40.0))
        if rng.random() < 0.1:
            hi = lo
        if rng.random() < 0.5:
            want = np.append(np.linspace(lo, hi, n - 1), v_cur)
            got = get_v_list(lo, hi, v_cur, None, None, n)
        else:
            g0, g1 = sorted(rng.uniform(0.0, 40.0, 2))
            want = np.append(np.linspace(max(min(g0, lo), 0.001), max(g1, hi), n - 1), v_cur)
            got = get_v_list(lo, hi, v_cur, g0, g1, n)
        assert got.dtype == np.float64 and got.tobytes() == want.tobytes(), (lo, hi, n)
    assert get_v_list(1.0, 2.0, 1.5, None, None, 2).tolist() == [1.0, 1.5]  # a single sample plus the current velocity
    with pytest.raises(ValueError):
        get_v_list(1.0, 2.0, 1.5, None, None, 0)


def test_array_wise_sweep_packer_equals_the_per_scenario_packing():
    """ScoringContext.pack_scenarios for a sweep over one grid shape fills the ABI structs array-wise (one stacked copy per
    input field, struct columns through numpy views of the ctypes layouts, one class table per obstacle type, spline tables
    shared when their bytes are equal): every record it hands to etp_plan_batch equals what the per-scenario packing gives."""
    import ctypes as C

    from ethicaltrajectoryplanning_b200.scoring import ScoringContext, pack_predictions
    from ethicaltrajectoryplanning_b200.spline import CubicSpline2D
    from ethicaltrajectoryplanning_b200.workloads import sweep_steps

    steps = [st for _, st in sweep_steps(9)]
    steps[4].csp = steps[3].csp  # the same spline object twice in a row
    arr, keep = ScoringContext._pack_scenarios_uniform(steps)
    assert len(arr) == len(steps)

    def view(ptr, shape, dtype):
        n = int(np.prod(shape))
        if n == 0:
            return np.zeros(shape, dtype=dtype)
        ct = C.c_double if dtype == np.float64 else C.c_int32
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).reshape(shape).copy()

    for sc, st in zip(arr, steps):
        o, s = sc.obstacles.contents, sc.step.contents
        pk = pack_predictions(st.predictions, st.obstacle_types)
        O, Tp = pk["O"], pk["Tp"]
        assert (o.n_obstacles, o.max_pred) == (O, Tp)
        np.testing.assert_array_equal(view(o.pred_len, (O,), np.int32), pk["pred_len"])
        for name, shape in (("pos", (O, Tp, 2)), ("cov", (O, Tp, 4)), ("yaw", (O, Tp)), ("vel", (O, Tp)), ("length", (O,)),
                            ("width", (O,)), ("mass", (O,))):
            np.testing.assert_array_equal(view(getattr(o, name), shape, np.float64), np.asarray(pk[name]).reshape(shape), err_msg=name)
        np.testing.assert_array_equal(view(o.is_protected, (O,), np.int32), pk["prot"])
        np.testing.assert_array_equal(view(o.responsibility, (O,), np.int32), pk["resp"])
        assert (s.nv, s.nt, s.nd) == (len(st.v_list), len(st.t_list), len(st.d_list))
        assert (s.c_begin, s.c_end, s.shard_count) == (0, s.nv * s.nt * s.nd, 1)
        for name in ("c_s", "c_s_d", "c_s_dd", "c_d", "c_d_d", "c_d_dd", "dt", "v_thr", "velocity_target", "cost_factor"):
            assert getattr(s, name) == getattr(st, name), name
        np.testing.assert_array_equal(view(s.v_list, (s.nv,), np.float64), st.v_list)
        np.testing.assert_array_equal(view(s.t_list, (s.nt,), np.float64), st.t_list)
        np.testing.assert_array_equal(view(s.d_list, (s.nd,), np.float64), st.d_list)
        np.testing.assert_array_equal(view(s.n_samples, (s.nt,), np.int32), [len(np.arange(0.0, tT, st.dt)) for tT in st.t_list])
        csp = CubicSpline2D.from_foreign(st.csp)
        assert sc.n_seg == len(csp.s) - 1
        np.testing.assert_array_equal(view(sc.knots, (sc.n_seg + 1,), np.float64), csp.s)
        np.testing.assert_array_equal(view(sc.coef_x, (sc.n_seg, 4), np.float64), csp.coef_x)
        np.testing.assert_array_equal(view(sc.coef_y, (sc.n_seg, 4), np.float64), csp.coef_y)
