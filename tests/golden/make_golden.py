"""Generate the golden fixtures in this directory by running the UNMODIFIED reference
functions (``/root/reference``) under the shims of ``oracle/ref_shim.py``.

Run in the build container only (the reference tree does not travel):

    python tests/golden/make_golden.py

Each ``case_*.npz`` stores the plain-array inputs of one planning step and the outputs of the
reference's own ``calc_frenet_trajectories`` -> ``sort_frenet_trajectories`` (``calc_risk``,
``check_validity``, ``calc_trajectory_costs``) -> stable sort, so the tests can replay the case
through the oracle port, the vectorised oracle and the CUDA path without the reference.

Host-only pieces the reference computes with un-vendored geometry libraries are pinned to the
defaults of SURVEY.md section 8b: road boundary never left, no OBB collision with predictions,
``get_cost_factor`` -> ``step.cost_factor``, ``velocity_costs`` -> the mode/target form.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from ethicaltrajectoryplanning_b200 import synthetic  # noqa: E402
from oracle import ref_shim  # noqa: E402

TRAJ_FIELDS = ("d", "d_d", "d_dd", "d_ddd", "s", "s_d", "s_dd", "s_ddd", "x", "y", "yaw", "v", "curv")
REASONS = ("", "velocity", "acceleration", "curvature (lvc)", "curvature (hvc)", "collision",
           "boundaries", "max_risk")


def run_reference(step, road_boundary=None, time_step=0, collision_checker=None):
    """One planning step through the reference's own functions; returns dense arrays."""
    ns = ref_shim.load_reference()
    ff = ns.frenet_functions
    ctc = ns.calc_trajectory_cost
    # host-side scalars (goal logic needs commonroad/pycrcc): pin to the step's cost context
    ctc.get_cost_factor = lambda **k: step.cost_factor

    def _vel_cost(traj, **k):
        m = step.velocity_cost_mode
        if m == 0:
            return abs(step.velocity_target - np.mean(traj.v))
        if m == 1:
            return 30.0 - np.mean(traj.v)
        if m == 2:
            return np.mean(traj.v)
        return 0.0

    ctc.velocity_costs = _vel_cost
    vp = ns.vehicleparams.VehicleParameters("bmw_320i")
    csp = ref_shim.RefSpline(step.csp.s, step.csp.coef_x, step.csp.coef_y)
    fp_list = ff.calc_frenet_trajectories(
        c_s=step.c_s, c_s_d=step.c_s_d, c_s_dd=step.c_s_dd, c_d=step.c_d, c_d_d=step.c_d_d,
        c_d_dd=step.c_d_dd, d_list=step.d_list, t_list=step.t_list, v_list=step.v_list, dt=step.dt,
        csp=csp, v_thr=step.v_thr)
    for i, fp in enumerate(fp_list):
        fp._gen_index = i
    scenario = ref_shim.ScenarioStub(ns, step.obstacle_types)
    ego_state = ref_shim.EgoStateStub(step.ego_position, step.ego_orientation, time_step, step.c_s_d, step.c_s_dd)
    # the probability arrays are not kept by calc_risk: capture them
    prob_log = {}
    cp = ns.collision_probability
    rc = ns.risk_costs
    orig = cp.get_collision_probability_fast

    def _capture(traj, predictions, vehicle_params, safety_margin=1.0):
        out = orig(traj=traj, predictions=predictions, vehicle_params=vehicle_params)
        prob_log[traj._gen_index] = out
        return out

    orig_m = cp.get_inv_mahalanobis_dist

    def _capture_m(traj, predictions, vehicle_params, safety_margin=1.0):
        out = orig_m(traj=traj, predictions=predictions, vehicle_params=vehicle_params)
        prob_log[traj._gen_index] = out
        return out

    rc.get_collision_probability_fast = _capture
    rc.get_inv_mahalanobis_dist = _capture_m
    try:
        valid, invalid, _ = ff.sort_frenet_trajectories(
            ego_state=ego_state, fp_list=fp_list, global_path=None, predictions=step.predictions,
            mode=step.mode, params=step.params, planning_problem=None, scenario=scenario,
            vehicle_params=vp, ego_id=1, dt=step.dt, sensor_radius=50.0, road_boundary=road_boundary,
            collision_checker=collision_checker, goal_area=None)
    finally:
        rc.get_collision_probability_fast = orig
        rc.get_inv_mahalanobis_dist = orig_m
    valid.sort(key=lambda fp: fp.cost, reverse=False)
    best = valid[0] if len(valid) > 0 else invalid[0]
    return fp_list, valid, invalid, best, prob_log


def densify(step, fp_list, valid, invalid, best, prob_log):
    C = len(fp_list)
    Tmax = max(len(fp.t) for fp in fp_list)
    oids = list(step.predictions.keys()) if step.predictions is not None else []
    O = len(oids)
    traj = np.full((13, C, Tmax), np.nan)
    nT = np.zeros(C, dtype=np.int64)
    prob = np.zeros((C, O, Tmax - 1))
    red = np.zeros((4, C, O))
    level = np.zeros(C, dtype=np.int64)
    reason = np.zeros(C, dtype=np.int64)
    cost = np.zeros(C)
    has_cost = np.zeros(C, dtype=bool)
    risk_terms = np.full((6, C), np.nan)  # bayes equality maximin responsibility ego total
    other_terms = np.full((2, C), np.nan)  # velocity, dist_to_global_path
    for i, fp in enumerate(fp_list):
        T = len(fp.t)
        nT[i] = T
        for f, name in enumerate(TRAJ_FIELDS):
            traj[f, i, :T] = getattr(fp, name)
        level[i] = fp.valid_level
        reason[i] = REASONS.index(fp.reason_invalid)
        cost[i] = fp.cost
        for j, oid in enumerate(oids):
            if i in prob_log:
                p = np.asarray(prob_log[i][oid], dtype=float)
                prob[i, j, :len(p)] = p
            if oid in fp.ego_risk_dict:
                red[0, i, j] = fp.ego_risk_dict[oid]
                red[1, i, j] = fp.obst_risk_dict[oid]
                red[2, i, j] = fp.ego_harm_dict[oid]
                red[3, i, j] = fp.obst_harm_dict[oid]
        if hasattr(fp, "cost_dict"):
            has_cost[i] = True
            rd = fp.cost_dict["risk_cost_dict"]
            for k, name in enumerate(("bayes", "equality", "maximin", "responsibility", "ego", "total")):
                if name in rd:
                    risk_terms[k, i] = rd[name]
            uw = fp.cost_dict["unweighted_cost"]
            if "velocity" in uw:
                other_terms[0, i] = uw["velocity"]
            if "dist_to_global_path" in uw:
                other_terms[1, i] = uw["dist_to_global_path"]
    return dict(ref_traj=traj, ref_nT=nT, ref_prob=prob, ref_red=red, ref_level=level, ref_reason=reason,
                ref_cost=cost, ref_has_cost=has_cost, ref_risk_terms=risk_terms, ref_other_terms=other_terms,
                ref_best=np.int64(best._gen_index),
                ref_valid_order=np.array([fp._gen_index for fp in valid], dtype=np.int64),
                ref_invalid_order=np.array([fp._gen_index for fp in invalid], dtype=np.int64))


def pack_inputs(step):
    oids = list(step.predictions.keys()) if step.predictions is not None else []
    d = dict(
        start=np.array([step.c_s, step.c_s_d, step.c_s_dd, step.c_d, step.c_d_d, step.c_d_dd]),
        v_list=np.asarray(step.v_list), t_list=np.asarray(step.t_list), d_list=np.asarray(step.d_list),
        dt=np.float64(step.dt), v_thr=np.float64(step.v_thr),
        knots=step.csp.s, coef_x=step.csp.coef_x, coef_y=step.csp.coef_y,
        ego_position=step.ego_position, ego_orientation=np.float64(step.ego_orientation),
        obstacle_ids=np.array(oids, dtype=np.int64),
        obstacle_types=np.array([step.obstacle_types[o] for o in oids]),
        params_json=np.array(json.dumps(step.params)),
        mode=np.array(step.mode), velocity_cost_mode=np.int64(step.velocity_cost_mode),
        velocity_target=np.float64(step.velocity_target), cost_factor=np.float64(step.cost_factor),
        vehicle=np.array("bmw_320i"),
    )
    for j, oid in enumerate(oids):
        p = step.predictions[oid]
        d[f"pos_{j}"] = np.asarray(p["pos_list"], dtype=float)
        d[f"cov_{j}"] = np.asarray(p["cov_list"], dtype=float)
        d[f"yaw_{j}"] = np.asarray(p["orientation_list"], dtype=float)
        d[f"vel_{j}"] = np.asarray(p["v_list"], dtype=float)
        d[f"shape_{j}"] = np.array([p["shape"]["length"], p["shape"]["width"]])
        d[f"resp_{j}"] = np.int64(p.get("responsibility", 0))
    return d


def cases():
    mk = synthetic.make_step
    out = {}
    out["cfg1_ethical"] = mk("cfg1", seed=1234)
    out["cfg1_ego"] = mk("cfg1", seed=7, weights="weights_ego.json")
    out["cfg1_standard"] = mk("cfg1", seed=8, weights="weights_standard.json")
    out["lowspeed_0p5"] = mk("cfg1", seed=11, v0=0.5)            # low-velocity lateral mode, skip rule
    out["lowspeed_2p9"] = mk("cfg1", seed=12, v0=2.9, c_d_d=0.3, c_d_dd=-0.2, c_s_dd=0.5)
    out["fast_30"] = mk("cfg1", seed=13, v0=30.0, c_d_d=-0.4)
    out["vmax_clamped"] = mk("cfg1", seed=14, v0=49.0)            # v_list hits v_max; duplicates v_cur
    out["planning_15x15"] = mk("planning", seed=21, lateral_spread=2.0)
    out["multi_t"] = mk("cfg1", seed=31, t_list=[2.0, 3.0], n_obstacles=4)
    out["short_pred"] = mk("cfg1", seed=41, n_pred=9, n_obstacles=4)   # Tp < T
    out["long_pred"] = mk("cfg1", seed=42, n_pred=35, n_obstacles=4)   # Tp > T
    out["no_obstacles"] = mk("cfg1", seed=51, n_obstacles=0)
    s = mk("cfg1", seed=61, lateral_spread=1.0)
    s.params["modes"]["max_acceptable_risk"] = 0.002             # forces level 3 on some candidates
    out["max_risk_level3"] = s
    for name, (ign, sym, red) in {"lr1s": (True, True, True), "lr12s": (False, True, False),
                                  "lr12a": (False, False, False), "lr4a": (False, False, True)}.items():
        s = mk("cfg1", seed=71, lateral_spread=1.5)
        s.params["modes"].update(ignore_angle=ign, sym_angle=sym, reduced_angle_areas=red)
        out[f"harm_{name}"] = s
    s = mk("cfg1", seed=81)
    s.params["modes"]["fast_prob_mahalanobis"] = True
    # the reference inverts every covariance: the all-zero one (Q14) would be singular there
    for oid, p in s.predictions.items():
        if not np.any(p["cov_list"]):
            p["cov_list"][:, 0, 0] = p["cov_list"][:, 1, 1] = 0.1
    out["mahalanobis"] = s
    s = mk("cfg1", seed=91, lateral_spread=1.0)
    s.params["modes"]["multiple_cost_functions"] = True
    s.params["modes"]["max_acceptable_risk"] = 1e-9              # everything level 3 -> risk-only weights
    out["multi_cost_fn"] = s
    s = mk("cfg1", seed=101, weights="weights_ethical.json")
    s.params["weights"].update(lon_jerk=0.3, lat_jerk=0.2, travelled_dist=0.1)
    out["jerk_travelled"] = s
    s = mk("cfg1", seed=111)
    s.velocity_cost_mode, s.cost_factor = 1, 0.01
    out["velmode1_factor"] = s
    # scenes of tests/test_gpu_random_scenes.py on a thinned grid: reference paths in every direction -- seeds 0, 3, 6, 9,
    # 12 run along the +-pi cut of the heading, so the ego-side impact angle leaves (-pi, pi] (quirk Q4) --, obstacles
    # with arbitrary headings (static ones keep an initial orientation outside (-pi, pi]), correlations up to +-0.9, all
    # five harm models (seed % 5), ragged prediction lengths, low- and high-speed lateral modes
    import importlib.util

    spec = importlib.util.spec_from_file_location("_scenes", os.path.join(os.path.dirname(HERE), "test_gpu_random_scenes.py"))
    scenes = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(scenes)
    for seed in (0, 1, 2, 3, 5, 6, 9, 12):
        s = scenes.random_scene(seed)
        s.v_list = np.ascontiguousarray(s.v_list[::2])
        s.d_list = np.linspace(-3.0, 3.0, 13)
        out[f"scene_{seed:02d}"] = s
    return out


def main():
    if not ref_shim.available():
        raise SystemExit("reference tree not available: fixtures can only be generated in the build container")
    only_missing = "--only-missing" in sys.argv
    for name, step in cases().items():
        if only_missing and os.path.exists(os.path.join(HERE, f"case_{name}.npz")):
            continue
        fp_list, valid, invalid, best, prob_log = run_reference(step)
        data = pack_inputs(step)
        data.update(densify(step, fp_list, valid, invalid, best, prob_log))
        path = os.path.join(HERE, f"case_{name}.npz")
        np.savez_compressed(path, **data)
        lv, cnt = np.unique(data["ref_level"], return_counts=True)
        ung = float((data["ref_prob"] != 0).mean()) if data["ref_prob"].size else 0.0
        print(f"{name:18s} C={len(fp_list):4d} best={int(data['ref_best']):4d} levels={dict(zip(lv.tolist(), cnt.tolist()))} "
              f"nonzero-prob={ung:.3f} size={os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
