"""TEST INFRASTRUCTURE -- not part of the product path (only tests/, smoke() and bench.py's CPU legs import oracle/).

fp64 restatement of the collision test of a candidate trajectory against the predicted obstacle boxes (validity
level 1): ``collision_valid`` / ``create_collision_object`` (validity_checks.py:125-141, 216-274) and
``collision_checker_prediction`` (prediction_helpers.py:138-210).

PARITY: the geometry primitives are pycrcc's (commonroad-drivability-checker, not vendored in the reference and not
installable here) -- the box-overlap primitive and the ``swept`` hull are UNPINNED.  Everything the reference's own sources
decide is pinned by fixtures the reference's functions wrote under oracle/ref_shim.py with stand-ins for those primitives
only (tests/golden/make_golden_collision.py -> collision_cases.npz, make_golden_boundary.py -> boundary_cases.npz:
``mode="discrete"``) -- which boxes exist, their half extents, orientations and time steps:

* ego:  ``RectOBB(l / 2, w / 2, yaw[i], x[i], y[i])`` for every sample i of the candidate, object starts at
  ``ego_state.time_step`` (validity_checks.py:127-133);
* obstacle: ``RectOBB(length / 2, width / 2, orientation_list[j], pos_list[j])`` for ``j < min(len(ft.t), len(pos_list))``,
  object starts at ``time_step + 1`` (prediction_helpers.py:164-186)  =>  ego box j + 1 meets prediction box j;
* both objects go through ``trajectory_preprocess_obb_sum`` and are used un-preprocessed when that fails
  (validity_checks.py:135-141, prediction_helpers.py:189-194).

``mode="discrete"`` is the un-preprocessed branch (two oriented boxes overlap or not: no freedom left besides the
measure-zero touching case).  ``mode="swept"`` stands for the preprocessed branch with OUR hull: boxes k and k + 1 of an
object are replaced by the box in the frame of box k that encloses both; an object with a single box keeps it.
Two independent overlap tests are kept here (separating axes, and corner containment + edge crossing) so the tests can
check one against the other.
"""
from __future__ import annotations

import numpy as np


# ------------------------------------------------------------------------------------------------
# scalar forms
# ------------------------------------------------------------------------------------------------
def box(x, y, yaw, half_l, half_w):
    return (float(x), float(y), float(np.cos(yaw)), float(np.sin(yaw)), float(half_l), float(half_w))


def sat_margin(a, b):
    """Largest separating-axis margin over the four face normals; > 0 <=> disjoint."""
    ax, ay, ac, as_, al, aw = a
    bx, by, bc, bs, bl, bw = b
    dx, dy = bx - ax, by - ay
    c, s = abs(ac * bc + as_ * bs), abs(as_ * bc - ac * bs)
    return max(abs(dx * ac + dy * as_) - (al + bl * c + bw * s), abs(dy * ac - dx * as_) - (aw + bl * s + bw * c),
               abs(dx * bc + dy * bs) - (bl + al * c + aw * s), abs(dy * bc - dx * bs) - (bw + al * s + aw * c))


def corners(b):
    x, y, c, s, l, w = b
    return [(x + sx * l * c - sy * w * s, y + sx * l * s + sy * w * c) for sx, sy in ((1, 1), (-1, 1), (-1, -1), (1, -1))]


def _inside(b, pt):
    x, y, c, s, l, w = b
    dx, dy = pt[0] - x, pt[1] - y
    return abs(dx * c + dy * s) <= l and abs(dy * c - dx * s) <= w


def _segments_cross(p, q, r, t):
    def orient(a, b, c):
        return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])
    return orient(p, q, r) * orient(p, q, t) < 0 and orient(r, t, p) * orient(r, t, q) < 0


def overlap_by_polygon(a, b):
    """Independent of the separating-axis form: a corner of one box inside the other, or two edges crossing."""
    ca, cb = corners(a), corners(b)
    if any(_inside(b, p) for p in ca) or any(_inside(a, p) for p in cb):
        return True
    return any(_segments_cross(ca[i], ca[(i + 1) % 4], cb[j], cb[(j + 1) % 4]) for i in range(4) for j in range(4))


def hull_box(a, b):
    """Box in the frame of ``a`` enclosing ``a`` and ``b``."""
    ax, ay, ac, as_, al, aw = a
    bx, by, bc, bs, bl, bw = b
    dx, dy = bx - ax, by - ay
    ux, uy = dx * ac + dy * as_, dy * ac - dx * as_
    c, s = abs(ac * bc + as_ * bs), abs(as_ * bc - ac * bs)
    ex, ey = bl * c + bw * s, bl * s + bw * c
    xlo, xhi = min(-al, ux - ex), max(al, ux + ex)
    ylo, yhi = min(-aw, uy - ey), max(aw, uy + ey)
    mx, my = 0.5 * (xlo + xhi), 0.5 * (ylo + yhi)
    return (ax + mx * ac - my * as_, ay + mx * as_ + my * ac, ac, as_, 0.5 * (xhi - xlo), 0.5 * (yhi - ylo))


def _object(boxes, swept):
    if swept and len(boxes) >= 2:
        return [hull_box(boxes[k], boxes[k + 1]) for k in range(len(boxes) - 1)]
    return list(boxes)


def collides_with_predictions(x, y, yaw, predictions, vehicle_l, vehicle_w, mode="discrete"):
    """One candidate (arrays of its T samples) against every prediction; loop form of the reference's call sequence."""
    swept = mode == "swept"
    T = len(x)
    ego = _object([box(x[i], y[i], yaw[i], vehicle_l / 2, vehicle_w / 2) for i in range(T)], swept)
    for oid in predictions:  # prediction_helpers.py:156
        p = predictions[oid]
        pos = np.asarray(p["pos_list"], dtype=np.float64)
        ori = np.asarray(p["orientation_list"], dtype=np.float64)
        pred_length = min(T, len(pos))  # :168
        if pred_length == 0:
            continue
        obs = _object([box(pos[j, 0], pos[j, 1], ori[j], p["shape"]["length"] / 2, p["shape"]["width"] / 2)
                       for j in range(pred_length)], swept)
        for j, b in enumerate(obs):  # entry j lives at time_step + 1 + j, ego entry i at time_step + i
            if j + 1 < len(ego) and not sat_margin(ego[j + 1], b) > 0.0:
                return True
    return False


# ------------------------------------------------------------------------------------------------
# vectorised over candidates (oracle/vec.py)
# ------------------------------------------------------------------------------------------------
def _sat_margin_v(a, b):
    ax, ay, ac, as_, al, aw = a
    bx, by, bc, bs, bl, bw = b
    dx, dy = bx - ax, by - ay
    c, s = np.abs(ac * bc + as_ * bs), np.abs(as_ * bc - ac * bs)
    m = np.maximum(np.abs(dx * ac + dy * as_) - (al + bl * c + bw * s), np.abs(dy * ac - dx * as_) - (aw + bl * s + bw * c))
    m = np.maximum(m, np.abs(dx * bc + dy * bs) - (bl + al * c + aw * s))
    return np.maximum(m, np.abs(dy * bc - dx * bs) - (bw + al * s + aw * c))


def _hull_v(a, b):
    ax, ay, ac, as_, al, aw = a
    bx, by, bc, bs, bl, bw = b
    dx, dy = bx - ax, by - ay
    ux, uy = dx * ac + dy * as_, dy * ac - dx * as_
    c, s = np.abs(ac * bc + as_ * bs), np.abs(as_ * bc - ac * bs)
    ex, ey = bl * c + bw * s, bl * s + bw * c
    xlo, xhi = np.minimum(-al, ux - ex), np.maximum(al, ux + ex)
    ylo, yhi = np.minimum(-aw, uy - ey), np.maximum(aw, uy + ey)
    mx, my = 0.5 * (xlo + xhi), 0.5 * (ylo + yhi)
    return (ax + mx * ac - my * as_, ay + mx * as_ + my * ac, ac, as_, 0.5 * (xhi - xlo), 0.5 * (yhi - ylo))


def _cut(b, sl):
    return tuple(v[..., sl] if isinstance(v, np.ndarray) else v for v in b)


def collision_mask(x, y, yaw, predictions, vehicle_l, vehicle_w, mode="discrete"):
    """x, y, yaw: [n, T] samples of n candidates of equal length -> bool[n]."""
    swept = mode == "swept"
    n, T = x.shape
    ego = (x, y, np.cos(yaw), np.sin(yaw), vehicle_l / 2, vehicle_w / 2)
    if swept and T >= 2:
        ego = _hull_v(_cut(ego, slice(0, T - 1)), _cut(ego, slice(1, T)))
    n_ego = ego[0].shape[1]
    hit = np.zeros(n, dtype=bool)
    for oid in predictions:
        p = predictions[oid]
        pos = np.asarray(p["pos_list"], dtype=np.float64)
        ori = np.asarray(p["orientation_list"], dtype=np.float64)
        pred_length = min(T, len(pos))
        if pred_length == 0:
            continue
        obs = (pos[:pred_length, 0], pos[:pred_length, 1], np.cos(ori[:pred_length]), np.sin(ori[:pred_length]),
               p["shape"]["length"] / 2, p["shape"]["width"] / 2)
        if swept and pred_length >= 2:
            obs = _hull_v(_cut(obs, slice(0, pred_length - 1)), _cut(obs, slice(1, pred_length)))
        k = min(len(obs[0]), n_ego - 1)
        if k <= 0:
            continue
        m = _sat_margin_v(_cut(ego, slice(1, k + 1)), tuple(v[None, :k] if isinstance(v, np.ndarray) else v for v in obs))
        hit |= (~(m > 0.0)).any(axis=1)
    return hit


# ------------------------------------------------------------------------------------------------
# road boundary: ego boxes against the triangles outside the road (validity_checks.py:216-245, risk_costs.py:104-123)
# PARITY UNPINNED like the rest of this file: pycrcc's static-obstacle query is not vendored.  Restated from the
# reference's own sources: which boxes exist (create_collision_object), that the FIRST colliding box i is reported
# (leaving_road_at[0] - ego_state.time_step) and that the harm is LR1S(traj.v[i]).  Brute force over all triangles
# (the product uses a uniform grid: the comparison checks the grid as well).
# ------------------------------------------------------------------------------------------------
def sat_margin_triangle(b, tri):
    """Largest separating-axis margin of box ``b`` and triangle ``tri`` [3, 2]: two box axes, three edge normals."""
    x, y, c, s, l, w = b
    u = np.array([[(p[0] - x) * c + (p[1] - y) * s, (p[1] - y) * c - (p[0] - x) * s] for p in tri])
    m = max(u[:, 0].min() - l, -u[:, 0].max() - l, u[:, 1].min() - w, -u[:, 1].max() - w)
    for k in range(3):
        j, i = (k + 1) % 3, (k + 2) % 3
        n = np.array([-(u[j, 1] - u[k, 1]), u[j, 0] - u[k, 0]])
        ln = np.hypot(n[0], n[1])
        if not ln > 0.0:
            continue
        n = n / ln
        a, bb = float(n @ u[k]), float(n @ u[i])
        r = l * abs(n[0]) + w * abs(n[1])
        m = max(m, min(a, bb) - r, -max(a, bb) - r)
    return m


def box_triangle_overlap_by_polygon(b, tri):
    """Independent of the separating-axis form: a vertex of one inside the other, or two edges crossing."""
    cb = corners(b)
    t = [tuple(p) for p in tri]

    def in_tri(p):
        d = [(t[(k + 1) % 3][0] - t[k][0]) * (p[1] - t[k][1]) - (t[(k + 1) % 3][1] - t[k][1]) * (p[0] - t[k][0]) for k in range(3)]
        return all(v >= 0 for v in d) or all(v <= 0 for v in d)

    if any(_inside(b, p) for p in t) or any(in_tri(p) for p in cb):
        return True
    return any(_segments_cross(cb[i], cb[(i + 1) % 4], t[j], t[(j + 1) % 3]) for i in range(4) for j in range(3))


def first_boundary_collision(x, y, yaw, triangles, vehicle_l, vehicle_w, mode="swept"):
    """Index of the first ego box that touches a triangle, or -1 (trajectories_collision_static_obstacles)."""
    T = len(x)
    ego = _object([box(x[i], y[i], yaw[i], vehicle_l / 2, vehicle_w / 2) for i in range(T)], mode == "swept")
    tris = np.asarray(triangles, dtype=np.float64).reshape(-1, 3, 2)
    for i, e in enumerate(ego):
        for tri in tris:
            if not sat_margin_triangle(e, tri) > 0.0:
                return i
    return -1


def first_boundary_collision_mask(x, y, yaw, triangles, vehicle_l, vehicle_w, mode="swept"):
    """x, y, yaw: [n, T] -> int[n] first colliding box per candidate or -1 (vectorised over candidates and boxes)."""
    n, T = x.shape
    ego = (x, y, np.cos(yaw), np.sin(yaw), np.full_like(x, vehicle_l / 2), np.full_like(x, vehicle_w / 2))
    if mode == "swept" and T >= 2:
        ego = _hull_v(_cut(ego, slice(0, T - 1)), _cut(ego, slice(1, T)))
    ex, ey, ec, es, el, ew = ego
    hit = np.zeros(ex.shape, dtype=bool)
    for tri in np.asarray(triangles, dtype=np.float64).reshape(-1, 3, 2):
        ux = [(p[0] - ex) * ec + (p[1] - ey) * es for p in tri]
        uy = [(p[1] - ey) * ec - (p[0] - ex) * es for p in tri]
        m = np.maximum(np.maximum(np.minimum.reduce(ux) - el, -np.maximum.reduce(ux) - el),
                       np.maximum(np.minimum.reduce(uy) - ew, -np.maximum.reduce(uy) - ew))
        for k in range(3):
            j, i = (k + 1) % 3, (k + 2) % 3
            nx, ny = -(uy[j] - uy[k]), ux[j] - ux[k]
            ln = np.sqrt(nx * nx + ny * ny)
            ok = ln > 0.0
            ln = np.where(ok, ln, 1.0)
            nx, ny = nx / ln, ny / ln
            a, b = nx * ux[k] + ny * uy[k], nx * ux[i] + ny * uy[i]
            r = el * np.abs(nx) + ew * np.abs(ny)
            m = np.where(ok, np.maximum(m, np.maximum(np.minimum(a, b) - r, -np.maximum(a, b) - r)), m)
        hit |= ~(m > 0.0)
    return np.where(hit.any(axis=1), hit.argmax(axis=1), -1)
