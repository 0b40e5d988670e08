"""Engine behind the reference-named entry points (``planner/…`` and ``risk_assessment/…`` mirrors).

The reference scores one ``FrenetTrajectory`` object at a time (``calc_risk``, ``check_validity``,
``calc_trajectory_costs``) from a Python loop (``frenet_functions.py:529-593``).  Here every one of those entry points
packs its trajectories into the candidate-minor field table of ``etp_score_trajectories`` (``include/etp_b200.h``) and
runs the same fused CUDA kernel as the planner step; lists produced by this package's ``calc_frenet_trajectories``
keep their sampling grid and are re-generated on the device instead of being uploaded.

There is no CPU path in here: without the CUDA library / a GPU every function raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _abi
from .configs import default_params
from .scoring import NoCandidateError, ScoringContext, _dptr, _iptr  # noqa: F401
from .vehicleparams import VehicleParameters

FIELD_NAMES = _abi.FIELD_NAMES
VALIDITY_LEVELS = {0: "Physically invalid", 1: "Collision", 2: "Leaving road boundaries",
                   3: "Maximum acceptable risk", 10: "Valid"}  # validity_checks.py:22-28

_CTX = {}


def default_context(precision=None, device=None) -> ScoringContext:
    """Process-wide scoring context (one per (device, precision)); ``ETP_B200_PRECISION=fp64`` selects the check mode."""
    precision = precision or os.environ.get("ETP_B200_PRECISION", "fp32")
    device = int(os.environ.get("ETP_B200_DEVICE", "0")) if device is None else device
    key = (device, precision)
    if key not in _CTX:
        _CTX[key] = ScoringContext(device, precision)
    return _CTX[key]


def close_contexts():
    for c in _CTX.values():
        c.close()
    _CTX.clear()


def obstacle_type_names(scenario, predictions, obstacle_types=None):
    """``scenario.obstacle_by_id(id).obstacle_type`` -> ObstacleType name, as harm_estimation.py:110-113,258-260,356."""
    if obstacle_types is not None:
        return {oid: getattr(obstacle_types[oid], "name", obstacle_types[oid]) for oid in predictions}
    out = {}
    for oid in predictions:
        t = scenario.obstacle_by_id(oid).obstacle_type
        out[oid] = str(getattr(t, "name", t)).upper()
    return out


def trajectory_fields(trajs):
    """FrenetTrajectory objects -> ([13][t_max][C] float64, n_samples[C]).  Missing attributes are zero."""
    C_ = len(trajs)
    n = np.array([len(t.x) for t in trajs], dtype=np.int32)
    t_max = int(n.max())
    f = np.zeros((_abi.NUM_FIELDS, t_max, C_), dtype=np.float64)
    for j, tr in enumerate(trajs):
        for i, name in enumerate(FIELD_NAMES):
            v = getattr(tr, name, None)
            if v is not None:
                v = np.asarray(v, dtype=np.float64)
                f[i, :min(len(v), t_max), j] = v[:t_max]
    return f, n


class GivenResult:
    """Host copies of one ``etp_score_trajectories`` call in logical layout."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def score_trajectories(trajs, predictions, obstacle_types, vehicle_params, params, mode="risk", ext_level=None,
                       bd_harm=None, velocity_cost_mode=3, velocity_target=0.0, cost_factor=1.0, want_harm=False,
                       masses=None, protections=None, ctx=None, goal=None, dt=0.1, road_boundary=None) -> GivenResult:
    """Score caller-provided trajectories on the device (fused kernel, generation replaced by loads).

    Obstacle classes come from ``obstacle_types`` ({id: ObstacleType name}) or explicit ``masses`` / ``protections``."""
    ctx = ctx or default_context()
    torch = ctx.torch
    ctx.set_params(params, vehicle_params, mode)
    ctx.set_predictions(predictions, obstacle_types, masses, protections)
    # (triangles, check): boundary test on the device.  A shared default context does not keep a boundary between calls;
    # a context the caller owns keeps what the caller set.
    if road_boundary is not None or (getattr(ctx, "_boundary_key", None) is not None and any(ctx is c for c in _CTX.values())):
        ctx.ensure_road_boundary(road_boundary)
    f, n = trajectory_fields(trajs)
    Cn, Tm, O = len(trajs), f.shape[1], ctx.O
    bufs = ctx.allocate_outputs(Cn, Tm, O)
    if want_harm:
        bufs["harm"] = torch.empty((2, Tm - 1, O, bufs["c_stride"]), device=ctx.device, dtype=ctx.dtype)[..., :Cn]
    g = _abi.Trajectories()
    g.n_traj, g.t_max = Cn, Tm
    g.n_samples, g.fields = _iptr(n), _dptr(f)
    g.velocity_cost_mode, g.velocity_target = int(velocity_cost_mode), float(velocity_target)
    keep = [f, n]
    if np.ndim(cost_factor) == 0:
        g.cost_factor = float(cost_factor)
    else:
        cf = np.ascontiguousarray(cost_factor, dtype=np.float64)
        g.cost_factor, g.cost_factor_list = 1.0, _dptr(cf)
        keep.append(cf)
    if goal is not None:  # GoalContext: velocity cost and cost factor per trajectory on the device
        from .scoring import make_goal

        gs = make_goal(goal, keep)
        g.goal, g.dt = C.pointer(gs), float(dt)
        keep.append(gs)
    if ext_level is not None:
        e = np.ascontiguousarray(ext_level, dtype=np.uint8)
        g.ext_level = e.ctypes.data
        keep.append(e)
    if bd_harm is not None:
        b = np.ascontiguousarray(bd_harm, dtype=np.float64)
        g.bd_harm = _dptr(b)
        keep.append(b)
    o = _abi.Out()
    o.precision, o.c_stride = ctx.precision, int(bufs["c_stride"])
    for k in ("traj", "prob", "red", "cost", "level", "reason", "harm", "bd_harm"):
        t = bufs.get(k)
        setattr(o, k, None if t is None else C.c_void_p(t.data_ptr()))
    ctx._check(ctx.lib.etp_score_trajectories(ctx.h, C.byref(g), C.byref(o)))
    best = ctx.best()
    host = {k: bufs[k].detach().cpu().numpy() for k in ("prob", "red", "cost", "level", "reason")}
    harm = bufs["harm"].detach().cpu().numpy() if want_harm else None
    return GivenResult(
        n_samples=n, obstacle_ids=list(ctx.obstacle_ids), best=best,
        prob=host["prob"].transpose(2, 0, 1).astype(np.float64),        # [C][T-1][O]
        red=host["red"].transpose(0, 2, 1).astype(np.float64),          # [4][C][O]
        cost=host["cost"].astype(np.float64),                           # [13][C]
        level=host["level"].astype(np.int64), reason=host["reason"].astype(np.int64),
        harm=None if harm is None else harm.transpose(0, 3, 1, 2).astype(np.float64),  # [2][C][T-1][O]
        bd_harm=bufs["bd_harm"].detach().cpu().numpy(),                  # [C] fp64
    )


def principle_costs(ego_risk_max, obst_risk_max, ego_harm_max, obst_harm_max, boundary_harm, responsibility=None,
                    eps=10e-10, scale_factor=10, ctx=None):
    """(bayes, equality, maximin, responsibility, ego) of one trajectory from its per-obstacle maxima
    (risk_costs.py:128-255) through ``etp_principle_costs``."""
    ctx = ctx or default_context()
    keys = list(ego_risk_max.keys())
    n = len(keys)
    arr = lambda d: np.ascontiguousarray([float(d[k]) for k in keys], dtype=np.float64)  # noqa: E731
    er, orr = arr(ego_risk_max), arr(obst_risk_max)
    eh = arr(ego_harm_max) if ego_harm_max is not None else np.zeros(n)
    oh = arr(obst_harm_max) if obst_harm_max is not None else np.zeros(n)
    rs = np.ascontiguousarray([int(responsibility[k]) for k in keys] if responsibility else [0] * n, dtype=np.int32)
    out = np.zeros(5)
    ctx._check(ctx.lib.etp_principle_costs(ctx.h, n, _dptr(er), _dptr(orr), _dptr(eh), _dptr(oh), _iptr(rs),
                                           float(boundary_harm), float(eps), float(scale_factor), _dptr(out)))
    return out


def params_with(modes=None, coeffs=None, weights=None):
    """params_dict = {'weights', 'modes', 'harm'} (frenet_planner.py:149-153), defaults for the missing parts."""
    p = default_params("weights_ethical.json")
    if weights is not None:
        p["weights"] = weights
    if modes is not None:
        p["modes"] = modes
    if coeffs is not None:
        p["harm"] = coeffs
    return p


def default_vehicle():
    return VehicleParameters("bmw_320i")
