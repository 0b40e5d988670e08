"""GPU parity of the reference-named entry points (planner/... and risk_assessment/... mirrors) against the golden
fixtures of the unmodified reference.  The calls below are the ones ``sort_frenet_trajectories`` and
``FrenetPlanner._step_planner`` make in the reference (frenet_functions.py:529-593, frenet_planner.py:326-576)."""
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest

from helpers import golden_cases, load_golden

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-4, 1e-9  # fp32 path (north star: 1e-4 relative), absolute floor of SURVEY.md section 7
FIELDS = ("d", "d_d", "d_dd", "d_ddd", "s", "s_d", "s_dd", "s_ddd", "x", "y", "yaw", "v", "curv")
CASES = ["cfg1_ethical", "cfg1_ego", "harm_lr12a", "short_pred", "multi_t", "max_risk_level3", "mahalanobis",
         "no_obstacles", "lowspeed_2p9"]


def _close(got, ref, rtol=RTOL, atol=ATOL):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return np.all(np.abs(got - ref) <= rtol * np.abs(ref) + atol)


class _Scenario:
    """``scenario.obstacle_by_id(id).obstacle_type`` as harm_estimation.py:110-113 uses it."""

    def __init__(self, types, dt=0.1):
        self.types, self.dt = types, dt

    def obstacle_by_id(self, oid):
        return SimpleNamespace(obstacle_type=SimpleNamespace(name=self.types[oid]))


def _ref_trajectories(step, ref):
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils.frenet_functions import FrenetTrajectory

    out = []
    for li in range(len(ref["nT"])):
        T = int(ref["nT"][li])
        kw = {n: ref["traj"][i, li, :T].copy() for i, n in enumerate(FIELDS)}
        out.append(FrenetTrajectory(t=np.arange(T) * step.dt, **kw))
    return out


def _problem(step):
    return SimpleNamespace(velocity_cost_mode=step.velocity_cost_mode, velocity_target=step.velocity_target,
                           cost_factor=step.cost_factor)


def _ego(step):
    return SimpleNamespace(position=step.ego_position, orientation=step.ego_orientation, time_step=0,
                           velocity=step.c_s_d, acceleration=step.c_s_dd)


@pytest.mark.parametrize("name", CASES)
def test_calc_frenet_trajectories_matches_reference(name):
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils.frenet_functions import calc_frenet_trajectories

    step, ref = load_golden(name)
    fps = calc_frenet_trajectories(step.c_s, step.c_s_d, step.c_s_dd, step.c_d, step.c_d_d, step.c_d_dd, step.d_list,
                                   step.t_list, step.v_list, step.dt, step.csp, step.v_thr)
    assert len(fps) == len(ref["nT"])
    for li, fp in enumerate(fps):
        T = int(ref["nT"][li])
        assert len(fp.t) == len(fp.x) == T
        for i, n in enumerate(FIELDS):
            want = ref["traj"][i, li, :T]
            err = np.abs(getattr(fp, n) - want) / np.maximum(np.abs(want), 1e-3)
            assert err.max() <= 2e-5, (n, li, err.max())


@pytest.mark.parametrize("name", ["cfg1_ethical", "harm_lr12a", "short_pred", "mahalanobis"])
def test_risk_assessment_entry_points(name):
    from ethicaltrajectoryplanning_b200.risk_assessment.collision_probability import (get_collision_probability_fast,
                                                                                      get_inv_mahalanobis_dist)
    from ethicaltrajectoryplanning_b200.risk_assessment.harm_estimation import get_harm
    from ethicaltrajectoryplanning_b200.risk_assessment.risk_costs import (calc_risk, get_bayesian_costs, get_ego_costs,
                                                                           get_equality_costs, get_maximin_costs,
                                                                           get_responsibility_cost)

    step, ref = load_golden(name)
    trajs = _ref_trajectories(step, ref)
    scenario = _Scenario(step.obstacle_types)
    oids = list(step.predictions.keys())
    mahal = bool(step.params["modes"]["fast_prob_mahalanobis"])
    for li in (0, len(trajs) // 2, len(trajs) - 1):
        tr = trajs[li]
        prob = (get_inv_mahalanobis_dist if mahal else get_collision_probability_fast)(tr, step.predictions, step.vehicle)
        assert list(prob.keys()) == oids
        for j, oid in enumerate(oids):
            assert len(prob[oid]) == len(tr.x) - 1
            assert _close(prob[oid], ref["prob"][li, j, :len(prob[oid])]), (li, oid)
        er, orr, eh, oh, bd = calc_risk(tr, _ego(step), step.predictions, scenario, 1, step.vehicle, step.params, None)
        assert bd == 0
        for j, oid in enumerate(oids):
            assert _close([er[oid], orr[oid], eh[oid], oh[oid]], ref["red"][:, li, j]), (li, oid)
        ego_h, obst_h = get_harm(scenario, tr, step.predictions, 1, step.vehicle, step.params["modes"], step.params["harm"], None)
        for j, oid in enumerate(oids):
            n = min(len(tr.x) - 1, len(step.predictions[oid]["pos_list"]))
            assert len(ego_h[oid]) == len(obst_h[oid]) == n
            assert _close([ego_h[oid].max(), obst_h[oid].max()], ref["red"][2:, li, j]), (li, oid)
        # principle costs from the reference's own per-obstacle maxima
        rer = {oid: ref["red"][0, li, j] for j, oid in enumerate(oids)}
        ror = {oid: ref["red"][1, li, j] for j, oid in enumerate(oids)}
        reh = {oid: ref["red"][2, li, j] for j, oid in enumerate(oids)}
        roh = {oid: ref["red"][3, li, j] for j, oid in enumerate(oids)}
        got = [get_bayesian_costs(rer, ror, 0), get_equality_costs(rer, ror), get_maximin_costs(rer, ror, reh, roh, 0),
               get_responsibility_cost(scenario, tr, _ego(step), ror, step.predictions, None)[0], get_ego_costs(rer, 0)]
        if ref["has_cost"][li]:
            assert _close(got, ref["risk_terms"][:5, li], rtol=1e-12, atol=1e-15), (li, got, ref["risk_terms"][:5, li])


def test_principle_costs_of_empty_dicts_are_zero():
    from ethicaltrajectoryplanning_b200.risk_assessment import risk_costs as rc

    assert rc.get_bayesian_costs({}, {}, 0.3) == 0
    assert rc.get_equality_costs({}, {}) == 0
    assert rc.get_maximin_costs({}, {}, {}, {}, 0.3) == 0
    assert rc.get_ego_costs({}, 0.3) == 0


@pytest.mark.parametrize("generated", [False, True])
@pytest.mark.parametrize("name", CASES)
def test_sort_frenet_trajectories(name, generated):
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils.frenet_functions import (calc_frenet_trajectories,
                                                                                      sort_frenet_trajectories)

    step, ref = load_golden(name)
    if generated:
        fps = calc_frenet_trajectories(step.c_s, step.c_s_d, step.c_s_dd, step.c_d, step.c_d_d, step.c_d_dd, step.d_list,
                                       step.t_list, step.v_list, step.dt, step.csp, step.v_thr)
    else:
        fps = _ref_trajectories(step, ref)
    for i, fp in enumerate(fps):
        fp._gen = i
    valid, invalid, vdict = sort_frenet_trajectories(
        ego_state=_ego(step), fp_list=fps, global_path=None, predictions=step.predictions, mode=step.mode,
        params=step.params, planning_problem=_problem(step), scenario=_Scenario(step.obstacle_types),
        vehicle_params=step.vehicle, ego_id=1, dt=step.dt, sensor_radius=50.0, road_boundary=None,
        collision_checker=None, goal_area=None)
    np.testing.assert_array_equal([fp.valid_level for fp in fps], ref["level"])
    reasons = ("", "velocity", "acceleration", "curvature (lvc)", "curvature (hvc)", "collision", "boundaries", "max_risk")
    assert [fp.reason_invalid for fp in fps] == [reasons[r] for r in ref["reason"]]
    # the two lists hold the reference's members in the reference's order (before / after the sort by cost)
    assert [fp._gen for fp in invalid] == list(ref["invalid_order"])
    valid.sort(key=lambda fp: fp.cost)
    assert valid[0]._gen == int(ref["best"])
    hc = ref["has_cost"]
    got = np.array([fp.cost for fp in fps])
    assert _close(got[hc], ref["cost"][hc], rtol=1e-3, atol=1e-6)
    assert sorted(fp._gen for fp in valid) == sorted(np.flatnonzero(hc))
    best = valid[0]
    assert set(best.cost_dict) == {"risk_cost_dict", "weights", "factor", "unweighted_cost", "total_cost"}
    if step.predictions is not None and step.mode == "risk" and len(step.predictions):
        li = best._gen
        rd = best.cost_dict["risk_cost_dict"]
        got = [rd[k] for k in ("bayes", "equality", "maximin", "responsibility", "ego", "total")]
        assert _close(got, ref["risk_terms"][:, li], rtol=1e-3, atol=1e-7)
        oids = list(step.predictions.keys())
        assert list(best.ego_risk_dict.keys()) == [o for o in oids]
        assert _close([best.obst_harm_dict[o] for o in oids], ref["red"][3, li])


@pytest.mark.parametrize("name", ["cfg1_ethical", "max_risk_level3", "lowspeed_2p9"])
def test_check_validity_and_calc_trajectory_costs(name):
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils.calc_trajectory_cost import calc_trajectory_costs
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils.validity_checks import VALIDITY_LEVELS, check_validity

    assert list(VALIDITY_LEVELS) == [0, 1, 2, 3, 10]
    step, ref = load_golden(name)
    trajs = _ref_trajectories(step, ref)
    scenario = _Scenario(step.obstacle_types)
    reasons = ("", "velocity", "acceleration", "curvature (lvc)", "curvature (hvc)", "collision", "boundaries", "max_risk")
    list_level = int(max(ref["level"]))
    for li in (0, len(trajs) // 2, len(trajs) - 1, int(ref["best"])):
        lvl, why = check_validity(trajs[li], _ego(step), scenario, step.vehicle, step.params["modes"], step.mode, None, None,
                                  predictions=step.predictions)
        assert (lvl, why) == (int(ref["level"][li]), reasons[int(ref["reason"][li])])
        if ref["has_cost"][li]:
            cost, cd = calc_trajectory_costs(trajs[li], None, _problem(step), _ego(step), list_level, scenario, step.params,
                                             1, step.dt, step.predictions, None, step.vehicle, mode=step.mode)
            assert _close(cost, ref["cost"][li], rtol=1e-4, atol=1e-9)
            assert _close(cd["total_cost"], ref["cost"][li], rtol=1e-4, atol=1e-9)


def test_frenet_planner_step_and_log_line(tmp_path):
    from ethicaltrajectoryplanning_b200.planner.Frenet.frenet_planner import FrenetPlanner
    from ethicaltrajectoryplanning_b200.planner.Frenet.utils.logging import get_data_from_line

    step, ref = load_golden("cfg1_ethical")
    log = os.path.join(tmp_path, "logs", "scenario.csv")
    fpar = {"t_list": list(step.t_list), "v_list_generation_mode": "linspace", "n_v_samples": len(step.v_list),
            "d_list": step.d_list, "dt": step.dt, "v_thr": step.v_thr}
    ego = _ego(step)
    results = {}
    for materialize in (True, False):
        pl = FrenetPlanner(_Scenario(step.obstacle_types), _problem(step), 1, step.vehicle, "risk", frenet_parameters=fpar,
                           weights=step.params["weights"], settings={"risk_dict": step.params["modes"]},
                           reference_spline=step.csp, predictor=lambda ts: step.predictions, ego_state=ego,
                           log_path=log if materialize else None, materialize=materialize)
        # start state of the fixture instead of the zero-initialised trajectory (planning.py:251-263)
        pl._trajectory["s_loc_m"][1], pl._trajectory["d_loc_m"][1] = step.c_s, step.c_d
        pl._trajectory["d_d_loc_mps"][1], pl._trajectory["d_dd_loc_mps2"][1] = step.c_d_d, step.c_d_dd
        pl.step(pl.scenario, None, 0, ego)
        results[materialize] = pl.trajectory
        li = int(ref["best"])
        T = int(ref["nT"][li])
        np.testing.assert_allclose(pl.trajectory["x_m"], ref["traj"][8, li, :T], rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(pl.trajectory["v_mps"], ref["traj"][5, li, :T], rtol=2e-5, atol=1e-6)  # Q7: v := s_d
        assert set(pl.trajectory) == {"s_loc_m", "d_loc_m", "d_d_loc_mps", "d_dd_loc_mps2", "x_m", "y_m", "psi_rad",
                                      "kappa_radpm", "v_mps", "ax_mps2", "time_s"}
        nxt = pl.ideal_tracking_state()
        assert nxt.time_step == 1 and nxt.velocity == pl.trajectory["v_mps"][1]
    np.testing.assert_allclose(results[True]["y_m"], results[False]["y_m"], rtol=2e-5, atol=1e-6)
    # the log line: reference header, seven ';'-separated JSON fields, reference key order per trajectory
    line = get_data_from_line(log, 1)
    assert list(line) == "simulation_step;time;valid_trajectories;invalid_trajectories;predictions;calc_time_avg;reach_set".split(";")
    assert len(line["valid_trajectories"]) + len(line["invalid_trajectories"]) == len(ref["nT"])
    first = line["valid_trajectories"][0]
    assert list(first)[:19] == ["t", "d", "d_d", "d_dd", "d_ddd", "s", "s_d", "s_dd", "s_ddd", "x", "y", "yaw", "v", "curv",
                                "valid_level", "reason_invalid", "cost", "ego_risk_dict", "obst_risk_dict"]
    assert {"ego_harm_dict", "obst_harm_dict", "bd_harm", "risk_dict", "cost_dict"} <= set(first)
    assert abs(first["cost"] - ref["cost"][int(ref["best"])]) <= 1e-4 * abs(ref["cost"][int(ref["best"])])
    assert list(line["predictions"]) == [str(o) for o in step.predictions]  # ids become JSON strings


def test_exec_timer_labels_of_a_planning_step():
    """the reference's label names stay at the boundary (frenet_planner.py:270-561,
    planner/utils/timers.py); the fused launches are timed on the device and filed below them."""
    from ethicaltrajectoryplanning_b200.planner.Frenet.frenet_planner import FrenetPlanner
    from ethicaltrajectoryplanning_b200.planner.utils.timers import ExecTimer

    step, ref = load_golden("cfg1_ethical")
    fpar = {"t_list": list(step.t_list), "v_list_generation_mode": "linspace", "n_v_samples": len(step.v_list),
            "d_list": step.d_list, "dt": step.dt, "v_thr": step.v_thr}
    for materialize in (True, False):
        timer = ExecTimer()
        pl = FrenetPlanner(_Scenario(step.obstacle_types), _problem(step), 1, step.vehicle, "risk", exec_timer=timer,
                           frenet_parameters=fpar, weights=step.params["weights"], settings={"risk_dict": step.params["modes"]},
                           reference_spline=step.csp, predictor=lambda ts: step.predictions, ego_state=_ego(step),
                           materialize=materialize)
        for k in range(2):
            pl.step(pl.scenario, None, k, _ego(step))
        d = timer.get_timing_dict()
        for label in ("simulation/total", "simulation/get v list", "simulation/prediction",
                      "simulation/calculate trajectories/total", "simulation/sort trajectories/total",
                      "simulation/sort trajectories/device/profile stage", "simulation/sort trajectories/device/fused scoring kernel",
                      "simulation/sort trajectories/device/selection"):
            assert label in d and len(d[label]) == 2, (materialize, label, sorted(d))
        assert 0.0 < d["simulation/sort trajectories/device/fused scoring kernel"][1] < d["simulation/sort trajectories/total"][1]
        assert d["simulation/total"][1] >= d["simulation/sort trajectories/total"][1]
        if materialize:
            assert "simulation/calculate trajectories/device/fused scoring kernel" in d
            assert "simulation/sort trajectories/sort list by costs" in d
        assert timer.running == {}
    # no timer given: nothing is recorded, nothing is synchronised for timing
    pl = FrenetPlanner(_Scenario(step.obstacle_types), _problem(step), 1, step.vehicle, "risk", frenet_parameters=fpar,
                       weights=step.params["weights"], settings={"risk_dict": step.params["modes"]},
                       reference_spline=step.csp, predictor=lambda ts: step.predictions, ego_state=_ego(step), materialize=False)
    pl.step(pl.scenario, None, 0, _ego(step))
    assert pl.exec_timer.get_timing_dict() == {}
