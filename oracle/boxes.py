"""TEST INFRASTRUCTURE -- not part of the product path (only tests/, smoke() and bench.py's CPU legs import oracle/).

fp64 restatement of the collision test of a candidate trajectory against the predicted obstacle boxes (validity
level 1): ``collision_valid`` / ``create_collision_object`` (validity_checks.py:125-141, 216-274) and
``collision_checker_prediction`` (prediction_helpers.py:138-210).

PARITY UNPINNED: the geometry itself is pycrcc's (commonroad-drivability-checker, not vendored in the reference and
not installable here).  What the reference's own sources fix is restated exactly -- which boxes exist, their half
extents, orientations and time steps:

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
