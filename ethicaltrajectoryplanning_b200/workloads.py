"""Synthetic drivers of the two sequential / batched configurations of BASELINE.json:

* cfg4 -- ``evaluatefrenet``-style sweep: many independent scenarios, one planning step each, scenario-sharded
  (``planner/Frenet/plannertools/evaluatefrenet.py:16-55``, ``planner/plannertools/evaluate.py:160-169``);
* cfg5 -- closed loop: replanning every 0.1 s with ideal tracking (``planner/planning.py:117-131``), obstacles
  advancing one sample per cycle, per-step latency.
"""
from __future__ import annotations

import time
from types import SimpleNamespace

import numpy as np

from .configs import default_params
from .planner.Frenet.frenet_planner import FrenetPlanner
from .synthetic import make_obstacles, make_step, synthetic_path
from .vehicleparams import VehicleParameters


def sweep_steps(n_scenarios, first=0, weights="weights_ethical.json"):
    """Scenario ``i`` of the sweep: the ``planning.json`` grid (15 x 15, T = 20), O ~ U{2..12} obstacles, seed i."""
    for i in range(first, first + n_scenarios):
        rng = np.random.default_rng(10_000 + i)
        yield i, make_step("planning", seed=10_000 + i, n_obstacles=int(rng.integers(2, 13)),
                           v0=float(rng.uniform(5.0, 20.0)), weights=weights)


def run_sweep(ctx, scenario_ids, weights="weights_ethical.json"):
    """One planning step per scenario (arg-min only).  Returns ([(scenario, dense index, level, cost)], candidates, seconds
    of the scored part: configure + launch + 40-byte readback; scenario synthesis is outside the timed region)."""
    steps = [st for i in scenario_ids for _, st in sweep_steps(1, i, weights)]
    out, n_cand = [], 0
    ctx.stream.synchronize()
    t0 = time.perf_counter()
    for sid, step in zip(scenario_ids, steps):
        res = ctx.plan(step, materialize=False)
        out.append((sid, res.best["index"], res.best["level"], res.best["cost"]))
        n_cand += res.best["n_scored"]
    return out, n_cand, time.perf_counter() - t0


def run_sweep_batched(ctx, scenario_ids, weights="weights_ethical.json", chunk=256):
    """The same sweep through ``etp_plan_batch`` (one C call per ``chunk`` scenarios, 8 scenarios in flight on the
    device).  Returns (results, candidates, seconds of packing, seconds of the batched calls)."""
    steps = [st for i in scenario_ids for _, st in sweep_steps(1, i, weights)]
    if not steps:
        return [], 0, 0.0, 0.0
    ctx.set_params(steps[0].params, steps[0].vehicle, steps[0].mode)
    t0 = time.perf_counter()
    packs = [ctx.pack_scenarios(steps[i:i + chunk]) for i in range(0, len(steps), chunk)]
    t_pack = time.perf_counter() - t0
    t0 = time.perf_counter()
    bests = []
    for pk in packs:
        bests += ctx.plan_batch(packed=pk)
    t_run = time.perf_counter() - t0
    out = [(sid, b["index"], b["level"], b["cost"]) for sid, b in zip(scenario_ids, bests)]
    return out, sum(b["n_scored"] for b in bests), t_pack, t_run


def run_sweep_pipelined(ctx, scenario_ids, weights="weights_ethical.json", chunk=500):
    """The sweep from prediction dicts to winners with packing and scoring overlapped (``plan_batch_pipelined``).  Returns
    (results, candidates, wall-clock seconds from the first dict to the last winner)."""
    steps = [st for i in scenario_ids for _, st in sweep_steps(1, i, weights)]
    if not steps:
        return [], 0, 0.0
    t0 = time.perf_counter()
    bests = ctx.plan_batch_pipelined(steps, chunk=chunk)
    t = time.perf_counter() - t0
    out = [(sid, b["index"], b["level"], b["cost"]) for sid, b in zip(scenario_ids, bests)]
    return out, sum(b["n_scored"] for b in bests), t


class MovingObstacles:
    """Predictor stand-in of the closed loop: constant-velocity Gaussian predictions that advance one sample per
    cycle (``predictor.step(time_step=...)`` contract of frenet_planner.py:386-390).  Like a real predictor it returns
    RAW predictions ({pos_list, cov_list}); orientation, velocity, shape margins and the responsibility flag are
    derived per cycle (on the device, by the planner)."""

    def __init__(self, csp, n_obstacles, horizon, n_steps, dt=0.1, v0=12.0, seed=1234):
        rng = np.random.default_rng(seed)
        self.horizon = horizon
        self.full, self.types = make_obstacles(csp, rng, n_obstacles, horizon + n_steps, dt, v0, horizon * dt)
        # raw obstacle shapes / initial states as a CommonRoad scenario holds them
        self.shapes = {oid: (p["shape"]["length"] - 1.0, p["shape"]["width"] - 0.5) for oid, p in self.full.items()}
        self.ori0 = {oid: float(np.asarray(p["orientation_list"])[0]) for oid, p in self.full.items()}
        self.vel0 = {oid: float(np.asarray(p["v_list"])[0]) for oid, p in self.full.items()}

    def obstacle_by_id(self, oid):
        return SimpleNamespace(obstacle_type=SimpleNamespace(name=self.types[oid]),
                               obstacle_shape=SimpleNamespace(length=self.shapes[oid][0], width=self.shapes[oid][1]),
                               initial_state=SimpleNamespace(orientation=self.ori0[oid], velocity=self.vel0[oid]))

    def step(self, time_step, obstacle_id_list=None, scenario=None):
        h = self.horizon
        return {oid: {"pos_list": p["pos_list"][time_step:time_step + h], "cov_list": p["cov_list"][:h]}
                for oid, p in self.full.items()}


def closed_loop(n_steps=200, weights="weights_ethical.json", grid="cfg2", v0=12.0, n_obstacles=8, materialize=False,
                precision=None, seed=1234, columnar_log_path=None, device_checks=False):
    """Replan every 0.1 s over ``n_steps`` cycles with ideal tracking.  Returns per-step wall-clock latencies [s] of
    ``FrenetPlanner.step`` (host packing + launches + readback of the winner) and the driven states.

    ``device_checks``: everything the reference asks pycrcc / CommonRoad for runs on the device as well -- collision with
    the predicted boxes (swept), a road boundary 3.6 m either side of the path, and a goal (lanelet-like polygon 150-190 m
    down the path, velocity / orientation / time intervals) driving cost factor and velocity cost."""
    from .synthetic import GRIDS

    nv, nd, dmax, t_list, _ = GRIDS[grid]
    veh = VehicleParameters("bmw_320i")
    csp = synthetic_path()
    T = len(np.arange(0.0, max(t_list), 0.1))
    pred = MovingObstacles(csp, n_obstacles, T, n_steps, v0=v0, seed=seed)
    params = default_params(weights)
    fpar = {"t_list": list(t_list), "v_list_generation_mode": "linspace", "n_v_samples": nv,
            "d_list": np.linspace(-dmax, dmax, nd), "dt": 0.1, "v_thr": 3.0}
    scenario = SimpleNamespace(dt=0.1, obstacle_by_id=pred.obstacle_by_id)
    problem = SimpleNamespace(velocity_cost_mode=0, velocity_target=v0, cost_factor=1.0)
    road_boundary = goal_area = None
    if device_checks:
        import copy

        from .planner.Frenet.utils.calc_trajectory_cost import GoalArea
        from .planner.Frenet.utils.validity_checks import RoadBoundary
        from .synthetic import road_boundary_triangles

        params = copy.deepcopy(params)
        params["modes"]["collision_check_device"] = "swept"
        road_boundary = RoadBoundary(road_boundary_triangles(csp), "swept")
        ring = []
        for s_, side in [(150.0, 1), (170.0, 1), (190.0, 1), (190.0, -1), (170.0, -1), (150.0, -1)]:
            px, py = csp.calc_position(s_)
            yaw = csp.calc_yaw(s_)
            ring.append((px - side * 3.6 * np.sin(yaw), py + side * 3.6 * np.cos(yaw)))
        goal_area = GoalArea([ring])
        cx, cy = csp.calc_position(170.0)
        state = SimpleNamespace(time_step=SimpleNamespace(start=100, end=n_steps + 150), velocity=SimpleNamespace(start=5.0, end=20.0),
                                orientation=SimpleNamespace(start=-1.0, end=1.0), position=SimpleNamespace(center=np.array([cx, cy])))
        problem = SimpleNamespace(goal=SimpleNamespace(state_list=[state]))
    x0, y0 = csp.calc_position(0.0)
    ego = SimpleNamespace(position=np.array([x0, y0]), orientation=float(csp.calc_yaw(0.0)), velocity=v0, acceleration=0.0,
                          time_step=0)
    pl = FrenetPlanner(scenario, problem, 1, veh, "risk", frenet_parameters=fpar, weights=params["weights"],
                       settings={"risk_dict": params["modes"]}, reference_spline=csp, predictor=pred, ego_state=ego,
                       road_boundary=road_boundary, goal_area=goal_area, materialize=materialize, precision=precision, columnar_log_path=columnar_log_path)
    lat, states = [], []
    for k in range(n_steps):
        t0 = time.perf_counter()
        pl.step(scenario, None, k, ego)
        lat.append(time.perf_counter() - t0)
        ego = pl.ideal_tracking_state()
        states.append((float(pl.trajectory["s_loc_m"][1]), float(pl.trajectory["d_loc_m"][1]), ego.velocity,
                       int(pl.last_best["index"]) if isinstance(pl.last_best, dict) else -1))
    return np.array(lat), states
