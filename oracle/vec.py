"""ORACLE (test infrastructure, not product code): vectorised fp64 restatement of the scoring path.

Same semantics as ``oracle/port.py`` (which follows the reference loop by loop), vectorised over
the lateral targets of one (v,t) profile, over time and over the un-gated (candidate, timestep)
pairs of an obstacle, so that 10^4..10^5 candidates finish in seconds.  Output is dense, in the
layout of the CUDA path (dense candidate index c = (iv*nt + it)*nd + id, skipped slots kept).

Pinned by ``tests/test_oracle_vec.py`` against the golden fixtures of the unmodified reference
(``tests/golden``) and against ``oracle/port.py``.  Reference lines are cited in ``oracle/port.py``;
only the differences in evaluation strategy are commented here.
"""
from __future__ import annotations

import numpy as np

from . import boxes as _boxes
from . import bvn as _bvn
from . import goal as _goal
from . import port as _port

LEVEL_SKIPPED = 255
COST_ROWS = _port.COST_ROWS
FIELDS = _port.Traj.FIELDS


def _quintic_batch(xs, vxs, axs, xe, T):
    """Coefficients for many end values ``xe`` sharing the start state and T (one LU, many RHS)."""
    a0 = np.full_like(xe, xs, dtype=np.float64)
    a1 = np.full_like(xe, vxs, dtype=np.float64)
    a2 = np.full_like(xe, axs / 2.0, dtype=np.float64)
    A = np.array([[T ** 3, T ** 4, T ** 5], [3 * T ** 2, 4 * T ** 3, 5 * T ** 4], [6 * T, 12 * T ** 2, 20 * T ** 3]])
    B = np.stack([xe - a0 - a1 * T - a2 * T ** 2, 0.0 - a1 - 2 * a2 * T, 0.0 - 2 * a2])
    x = np.stack([np.linalg.solve(A, B[:, j]) for j in range(B.shape[1])], axis=1) if B.shape[1] else np.zeros((3, 0))
    return a0, a1, a2, x[0], x[1], x[2]


def _poly(co, tau, order):
    a0, a1, a2, a3, a4, a5 = [c[:, None] for c in co]
    t = tau
    if order == 0:
        return a0 + a1 * t + a2 * np.power(t, 2) + a3 * np.power(t, 3) + a4 * np.power(t, 4) + a5 * np.power(t, 5)
    if order == 1:
        return a1 + 2 * a2 * t + 3 * a3 * np.power(t, 2) + 4 * a4 * np.power(t, 3) + 5 * a5 * np.power(t, 4)
    if order == 2:
        return 2 * a2 + 6 * a3 * t + 12 * a4 * np.power(t, 2) + 20 * a5 * np.power(t, 3)
    return 6 * a3 + 24 * a4 * t + 60 * a5 * np.power(t, 2)


def _angle_coef(angle, modes, coeff):
    """Vectorised impact-area coefficient; same band rules as port.angle_coefficient."""
    lr = coeff["log_reg"]
    if modes["ignore_angle"]:
        p = lr["ignore_angle"]
        return np.zeros_like(angle), p["const"], p["speed"]
    sym, red = modes["sym_angle"], modes["reduced_angle_areas"]
    deg = np.pi / 180
    if red:
        p = lr["reduced_sym_angle_areas"] if sym else lr["reduced_angle_areas"]
        thr = [45 / 180 * np.pi, (3 * (45 / 180 * np.pi)) if sym else 135 / 180 * np.pi]
        pos = [0.0, p["side"] if sym else p["driver_side"], p["rear"]]
        neg = [0.0, p["side"] if sym else p["right_side"], p["rear"]]
    else:
        p = lr["complete_sym_angle_areas"] if sym else lr["complete_angle_areas"]
        thr = [k / 180 * np.pi for k in (15, 45, 75, 105, 135, 165)]
        if sym:
            pos = neg = [0.0, p["Imp_1_11"], p["Imp_2_10"], p["Imp_3_9"], p["Imp_4_8"], p["Imp_5_7"], p["Imp_6"]]
        else:
            pos = [0.0, p["Imp_11"], p["Imp_10"], p["Imp_9"], p["Imp_8"], p["Imp_7"], p["Imp_6"]]
            neg = [0.0, p["Imp_1"], p["Imp_2"], p["Imp_3"], p["Imp_4"], p["Imp_5"], p["Imp_6"]]
    a = np.abs(angle)
    band = np.zeros(angle.shape, dtype=np.int64)
    for i, th in enumerate(thr):
        band += (~(a < th)) & (band == i)
    pos, neg = np.asarray(pos, dtype=np.float64), np.asarray(neg, dtype=np.float64)
    return np.where(angle < 0, neg[band], pos[band]), p["const"], p["speed"]


def _prob_obstacle(x, y, yaw, T, pred, vp):
    """Collision probability [n, T-1] of n candidates against one obstacle (collision_probability.py:136-256)."""
    n = x.shape[0]
    mean = np.asarray(pred["pos_list"], dtype=np.float64)
    cov = np.asarray(pred["cov_list"], dtype=np.float64)
    yaw_o = np.asarray(pred["orientation_list"], dtype=np.float64)
    length = pred["shape"]["length"]
    Tp = len(mean)
    prob = np.zeros((n, T - 1))
    min_len = min(T, Tp)
    if min_len < 2:
        return prob
    k = min_len - 1  # probability samples i = 1 .. min_len-1, index t = i-1
    dev = np.stack((np.cos(yaw_o[1:min_len]), np.sin(yaw_o[1:min_len])), axis=-1) * length / 2
    m0 = mean[:k]
    means = np.array([m0, m0 + dev, m0 - dev])  # [3, k, 2]
    ex, ey = x[:, 1:min_len], y[:, 1:min_len]   # [n, k]
    dist = np.sqrt((means[:, None, :, 0] - ex[None]) ** 2 + (means[:, None, :, 1] - ey[None]) ** 2)
    ungated = ~(dist.min(axis=0) > 5.0)
    ci, ti = np.nonzero(ungated)
    if len(ci) == 0:
        return prob
    c = cov[ti].copy()
    zero = np.all(c.reshape(-1, 4) == 0, axis=1)
    c[zero] = np.array([[0.1, 0.0], [0.0, 0.1]])
    s0, s1 = np.sqrt(c[:, 0, 0]), np.sqrt(c[:, 1, 1])
    rho = c[:, 1, 0] / s1 / s0
    px, py = ex[ci, ti], ey[ci, ti]
    eyaw = yaw[:, 1:min_len][ci, ti]
    r_x = vp.l / 2
    ax, ay = np.cos(eyaw), np.sin(eyaw)
    off = r_x * (2 / 3)
    cx = np.stack([px, px + off * ax, px - off * ax])  # [3, m]
    cy = np.stack([py, py + off * ay, py - off * ay])
    hx, hy = vp.l / 6, vp.w / 2
    total = np.zeros(len(ci))
    for km in range(3):
        mux, muy = means[km, ti, 0], means[km, ti, 1]
        for jb in range(3):
            xl = ((cx[jb] - hx) - mux) / s0
            xu = ((cx[jb] + hx) - mux) / s0
            yl = ((cy[jb] - hy) - muy) / s1
            yu = ((cy[jb] + hy) - muy) / s1
            total = total + _bvn.rect_prob_std(xl, xu, yl, yu, rho)
    prob[ci, ti] = total / 3
    return prob


def _mahal_obstacle(x, y, T, pred):
    n = x.shape[0]
    mean = np.asarray(pred["pos_list"], dtype=np.float64)
    inv = np.linalg.inv(np.asarray(pred["cov_list"], dtype=np.float64))
    Tp = len(mean)
    out = np.zeros((n, T - 1))
    k = min(T, Tp) - 1
    if k < 1:
        return out
    dx = x[:, 1:k + 1] - mean[:k, 0]
    dy = y[:, 1:k + 1] - mean[:k, 1]
    m2 = dx * dx * inv[:k, 0, 0] + dx * dy * (inv[:k, 0, 1] + inv[:k, 1, 0]) + dy * dy * inv[:k, 1, 1]
    out[:, :k] = 1e-4 / np.sqrt(m2)
    return out


def score_dense(step, profiles=None):
    """Score every candidate of the selected profiles (default: all).  Returns a dict of dense arrays
    over the candidates of those profiles, in dense order, plus ``dense_index``."""
    nv, nt, nd = len(step.v_list), len(step.t_list), len(step.d_list)
    if profiles is None:
        profiles = range(nv * nt)
    profiles = list(profiles)
    nsamp = [len(np.arange(0.0, tT, step.dt)) for tT in step.t_list]
    Tmax = max(nsamp)
    preds = step.predictions
    oids = list(preds.keys()) if preds is not None else []
    O = len(oids)
    C = len(profiles) * nd
    out = dict(
        dense_index=np.zeros(C, dtype=np.int64), T=np.zeros(C, dtype=np.int64),
        traj=np.zeros((13, C, Tmax)), prob=np.zeros((C, Tmax - 1, O)), red=np.zeros((4, C, O)),
        cost=np.zeros((len(COST_ROWS), C)), level=np.full(C, LEVEL_SKIPPED, dtype=np.int64),
        reason=np.zeros(C, dtype=np.int64),
    )
    vp, params = step.vehicle, step.params
    weights, modes, coeffs = params["weights"], params["modes"], params["harm"]
    if modes["harm_mode"] != "log_reg" or modes["crash_angle_simplified"] is False:
        raise ValueError("outside the restated path")
    radius = np.sqrt((vp.l ** 2 / np.tan(vp.steering.max) ** 2) + (vp.l_r ** 2))
    thr_v = np.sqrt(vp.lateral_a_max * radius)
    d_list = np.asarray(step.d_list, dtype=np.float64)
    for ip, p in enumerate(profiles):
        iv, it = divmod(p, nt)
        vT, tT = step.v_list[iv], step.t_list[it]
        T = nsamp[it]
        rows = slice(ip * nd, (ip + 1) * nd)
        out["dense_index"][rows] = p * nd + np.arange(nd)
        out["T"][rows] = T
        low = abs(step.c_s_d) < step.v_thr or abs(vT) < step.v_thr
        qp = _port.Quartic(xs=step.c_s, vxs=step.c_s_d, axs=step.c_s_dd, vxe=vT, axe=0.0, T=tT - step.dt)
        t = np.arange(0.0, tT, step.dt)
        s, s_d, s_dd, s_ddd = qp.p(t), qp.d1(t), qp.d2(t), qp.d3(t)
        gx, gy, gyaw, gcurv, gcurv_d = _port.reference_path_profile(step.csp, s, _port.relaxed_quirks(step))
        ds = s[-1] - s[0]
        keep = ~(ds <= np.abs(d_list))
        idx = np.flatnonzero(keep)
        if len(idx) == 0:
            continue
        dT = d_list[idx]
        n = len(idx)
        if not low:
            co = _quintic_batch(step.c_d, step.c_d_d, step.c_d_dd, dT, tT - step.dt)
            tau = t[None, :]
            d, d_d, d_dd, d_ddd = _poly(co, tau, 0), _poly(co, tau, 1), _poly(co, tau, 2), _poly(co, tau, 3)
            d_d_time, d_dd_time = d_d, d_dd
            d_d = d_d / s_d
            d_dd = (d_dd - d_d * s_dd) / np.power(s_d, 2)
        else:
            if step.c_s_d != 0.0:
                a = step.c_d_d / step.c_s_d
                b = (step.c_d_dd - step.c_s_dd * a) / (step.c_s_d ** 2)
            else:
                a = b = 0.0
            co = _quintic_batch(step.c_d, a, b, dT, ds)
            tau = (s - s[0])[None, :]
            d, d_d, d_dd, d_ddd = _poly(co, tau, 0), _poly(co, tau, 1), _poly(co, tau, 2), _poly(co, tau, 3)
            d_d_time = s_d * d_d
            d_dd_time = s_dd * d_d + np.power(s_d, 2) * d_dd
        yaw_diff = np.arctan(d_d / (1 - gcurv * d))
        yaw = yaw_diff + gyaw
        x = gx - d * np.sin(gyaw)
        y = gy + d * np.cos(gyaw)
        v = (s_d * (1 - gcurv * d)) / np.cos(yaw_diff)
        curv = (((d_dd + (gcurv_d * d + gcurv * d_d) * np.tan(yaw_diff))
                 * (np.power(np.cos(yaw_diff), 2) / (1 - gcurv * d))) + gcurv) * (np.cos(yaw_diff) / (1 - gcurv * d))
        ones = np.ones((n, 1))
        fields = [d, d_d_time, d_dd_time, d_ddd, s * ones, s_d * ones, s_dd * ones, s_ddd * ones, x, y, yaw, v, curv]
        gl = ip * nd + idx
        for f, arr in enumerate(fields):
            out["traj"][f, gl, :T] = arr
        # ---- risk -------------------------------------------------------------------------------
        er = np.zeros((n, O)); orr = np.zeros((n, O)); eh = np.zeros((n, O)); oh = np.zeros((n, O))
        for j, oid in enumerate(oids):
            pred = preds[oid]
            if modes["fast_prob_mahalanobis"]:
                pr = _mahal_obstacle(x, y, T, pred)
            else:
                pr = _prob_obstacle(x, y, yaw, T, pred, vp)
            out["prob"][gl, :T - 1, j] = pr
            pos = np.asarray(pred["pos_list"], dtype=np.float64)
            L = min(T - 1, len(pos))
            pv = np.asarray(pred["v_list"], dtype=np.float64)[:L]
            pyaw = np.asarray(pred["orientation_list"], dtype=np.float64)[:L]
            tname = step.obstacle_types[oid]
            m_obs = _port.obstacle_mass(tname, pred["shape"]["length"] * pred["shape"]["width"])
            pdof = pyaw - yaw[:, :L] + np.pi
            rel = np.arctan2(pos[:L, 1] - y[:, :L], pos[:L, 0] - x[:, :L])
            a_e = rel - yaw[:, :L]
            a_o = np.pi + rel - pyaw
            dv = np.sqrt(np.power(v[:, :L], 2) + np.power(pv, 2) + 2 * v[:, :L] * pv * np.cos(pdof))
            dve = m_obs / (vp.m + m_obs) * dv
            dvo = vp.m / (vp.m + m_obs) * dv
            if tname in _port.PROTECTED:
                ce, const, speed = _angle_coef(a_e, modes, coeffs)
                co_, _, _ = _angle_coef(a_o, modes, coeffs)
                he = 1 / (1 + np.exp(-const - speed * dve - ce))
                ho = 1 / (1 + np.exp(-const - speed * dvo - co_))
            elif tname in _port.UNPROTECTED:
                ia = coeffs["log_reg"]["ignore_angle"]
                he = 1 / (1 + np.exp(-ia["const"] - ia["speed"] * dve))
                ho = 1 / (1 + np.exp(coeffs["pedestrian"]["const"] - coeffs["pedestrian"]["speed"] * dvo))
            else:
                raise ValueError("unsupported obstacle type")
            er[:, j] = (he * pr[:, :L]).max(axis=1)
            orr[:, j] = (ho * pr[:, :L]).max(axis=1)
            eh[:, j] = he.max(axis=1)
            oh[:, j] = ho.max(axis=1)
        out["red"][0, gl], out["red"][1, gl], out["red"][2, gl], out["red"][3, gl] = er, orr, eh, oh
        # ---- validity -----------------------------------------------------------------------------
        vel_ok = bool(np.all(np.abs(s_d) <= vp.longitudinal.v_max))
        cmax = np.where(v < thr_v, 1.0 / radius, vp.lateral_a_max / (v ** 2))
        fail = np.abs(curv) > cmax
        first = np.where(fail.any(axis=1), fail.argmax(axis=1), -1)
        hv = np.where(first >= 0, ~(v[np.arange(n), np.maximum(first, 0)] < thr_v), False)
        bd = np.zeros(n) if step.bd_harm is None else np.asarray(step.bd_harm, dtype=np.float64)[p * nd + idx]
        ext = np.full(n, 10) if step.ext_level is None else np.asarray(step.ext_level)[p * nd + idx].astype(np.int64)
        rb = getattr(step, "road_boundary", None)
        if rb is not None:  # product extension, see oracle/boxes.py (risk_costs.py:104-123, validity_checks.py:108-115)
            leave = _boxes.first_boundary_collision_mask(x, y, yaw, rb[0], vp.l, vp.w, rb[1])
            ia = coeffs["log_reg"]["ignore_angle"]
            vl = v[np.arange(n), np.maximum(leave, 0)]
            bd = np.where(leave >= 0, 1 / (1 + np.exp(-ia["const"] - ia["speed"] * vl)), 0.0)
            ext = np.where(ext == 1, 1, np.where(leave >= 0, 2, 10))
        check = modes.get("collision_check_device", "host")  # product extension, see oracle/boxes.py
        if check != "host" and preds is not None and step.mode in ("WaleNet", "risk"):
            ext = np.where(_boxes.collision_mask(x, y, yaw, preds, vp.l, vp.w, check), 1, ext)
        # first failing test wins (validity_checks.py:63-122): apply in reverse order of precedence
        level = np.full(n, 10)
        reason = np.zeros(n, dtype=np.int64)
        if step.mode == "risk" and preds is not None and O > 0:
            m = (er.sum(axis=1) > modes["max_acceptable_risk"]) | (orr.max(axis=1) > modes["max_acceptable_risk"])
            level[m], reason[m] = 3, 7
        m = (bd != 0) if preds is not None else (ext == 2)
        level[m], reason[m] = 2, 6
        m = ext == 1
        level[m], reason[m] = 1, 5
        m = first >= 0
        level[m] = 0
        reason[m] = np.where(hv[m], 4, 3)
        if "acceleration_check" in _port.relaxed_quirks(step) and not bool(np.all(np.abs(s_dd) <= vp.longitudinal.a_max)):
            level[:], reason[:] = 0, 2  # relaxed Q2: the acceleration test tests the acceleration
        if not vel_ok:
            level[:], reason[:] = 0, 1
        out["level"][gl] = level
        out["reason"][gl] = reason
        # ---- costs ----------------------------------------------------------------------------------
        cst = np.zeros((len(COST_ROWS), n))
        if O > 0:
            cst[0] = (er.sum(axis=1) + orr.sum(axis=1) + bd) / (O * 2)
            cst[1] = np.abs(er - orr).sum(axis=1) / O
            gate = "maximin_gate" in _port.relaxed_quirks(step)  # relaxed Q3 keeps the harm where the risk is NOT below eps
            mm = np.maximum((eh * ((er < 10e-10) != gate)).max(axis=1), (oh * ((orr < 10e-10) != gate)).max(axis=1))
            cst[2] = np.maximum(mm, bd) ** 10
            resp = np.array([preds[o].get("responsibility", 0) for o in oids], dtype=np.float64)
            cst[3] = -(orr * resp).sum(axis=1)
            cst[4] = er.sum(axis=1) + bd
        risk_in_cost = weights["risk_cost"] > 0.0 and preds is not None and step.mode == "risk"
        if risk_in_cost:
            cst[5] = (weights["bayes"] * cst[0] + weights["equality"] * cst[1] + weights["maximin"] * cst[2]
                      + weights["responsibility"] * weights["bayes"] * cst[3] + weights["ego"] * cst[4])
        mv = v.mean(axis=1)
        cst[6] = {0: np.abs(step.velocity_target - mv), 1: 30.0 - mv, 2: mv}.get(step.velocity_cost_mode, np.zeros(n))
        cst[7] = np.abs(d).mean(axis=1)
        cst[8] = np.full(n, np.sum(np.power(s_ddd, 2)) / T)
        cst[9] = np.power(d_ddd, 2).sum(axis=1) / T
        cst[10] = np.sqrt(np.diff(x, axis=1) ** 2 + np.diff(y, axis=1) ** 2).sum(axis=1)
        cst[11] = step.cost_factor
        if getattr(step, "cost_factor_list", None) is not None:  # per-candidate get_cost_factor results (etp_step.cost_factor_list)
            cst[11] = np.asarray(step.cost_factor_list, dtype=np.float64)[p * nd + idx]
        if getattr(step, "goal", None) is not None:  # oracle/goal.py (calc_trajectory_cost.py:349-415, 438-519)
            cst[11], cst[6] = _goal.factor_and_velocity_cost(step.goal, x, y, v, yaw, t, step.dt)
        if weights.get("dist_to_goal_pos", 0.0) > 0.0:
            g = getattr(step, "goal", None)
            cst[13] = [_goal.calc_dist_to_goal_pos(g, x[i, -1], y[i, -1]) for i in range(n)]
        risk_only = (level < 10) & bool(modes["multiple_cost_functions"])
        total = np.zeros(n)
        if risk_in_cost:
            total += weights["risk_cost"] * cst[5]
        for row, key in ((8, "lon_jerk"), (9, "lat_jerk"), (6, "velocity"), (7, "dist_to_global_path"), (10, "travelled_dist"),
                         (13, "dist_to_goal_pos")):
            if weights[key] > 0.0:
                total += np.where(risk_only, 0.0, weights[key]) * cst[row]
        cst[12] = cst[11] * total
        out["cost"][:, gl] = cst
    # ---- selection: highest non-empty level, then cost, then generation order ---------------------------
    valid = out["level"] != LEVEL_SKIPPED
    if not valid.any():
        raise ValueError("max() arg is an empty sequence")  # what the reference raises (Q16)
    rank = np.where(out["level"] == 10, 4, out["level"])
    top = rank[valid].max()
    cand = np.flatnonzero(valid & (rank == top))
    best_local = cand[np.argmin(out["cost"][12, cand])]  # argmin keeps the first of equal costs
    out["best"] = dict(index=int(out["dense_index"][best_local]), level=int(out["level"][best_local]),
                       cost=float(out["cost"][12, best_local]),
                       list_index=int(np.count_nonzero(valid[:best_local])), n_scored=int(valid.sum()))
    out["Tmax"] = Tmax
    return out
