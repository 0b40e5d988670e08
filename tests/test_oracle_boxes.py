"""CPU checks of the box-collision restatement (oracle/boxes.py): the separating-axis test against an independent
corner / edge-crossing test, the hull construction, and the loop form against the vectorised form."""
import copy

import numpy as np

from oracle import boxes


def test_separating_axes_agree_with_corner_and_edge_test():
    rng = np.random.default_rng(5)
    n_hit = 0
    for _ in range(5000):
        a = boxes.box(rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-7, 7), rng.uniform(0.3, 3), rng.uniform(0.3, 1.5))
        b = boxes.box(rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-7, 7), rng.uniform(0.3, 3), rng.uniform(0.3, 1.5))
        m = boxes.sat_margin(a, b)
        if abs(m) < 1e-9:
            continue
        assert (m <= 0) == boxes.overlap_by_polygon(a, b)
        n_hit += m <= 0
    assert 1000 < n_hit < 4000


def test_known_boxes():
    a = boxes.box(0, 0, 0, 2, 1)
    assert boxes.sat_margin(a, boxes.box(3.9, 0, 0, 2, 1)) < 0          # overlapping along x
    assert boxes.sat_margin(a, boxes.box(4.1, 0, 0, 2, 1)) > 0
    assert boxes.sat_margin(a, boxes.box(0, 2.1, 0, 2, 1)) > 0
    # a diamond next to the corner: the face normals of the rotated box separate it, the axis-aligned ones do not
    d = boxes.box(2.9, 1.9, np.pi / 4, 0.8, 0.8)
    assert abs(2.9 - 0) < 2 + 0.8 * np.sqrt(2) and abs(1.9) < 1 + 0.8 * np.sqrt(2)
    assert boxes.sat_margin(a, d) > 0 and not boxes.overlap_by_polygon(a, d)


def test_hull_box_encloses_both_and_is_tight_along_its_axes():
    rng = np.random.default_rng(6)
    for _ in range(500):
        a = boxes.box(rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-7, 7), rng.uniform(0.3, 3), rng.uniform(0.3, 1.5))
        b = boxes.box(a[0] + rng.uniform(-2, 2), a[1] + rng.uniform(-2, 2), rng.uniform(-7, 7), a[4], a[5])
        h = boxes.hull_box(a, b)
        x, y, c, s, l, w = h
        assert (c, s) == (a[2], a[3])
        u = np.array([[(p[0] - x) * c + (p[1] - y) * s, (p[1] - y) * c - (p[0] - x) * s] for p in boxes.corners(a) + boxes.corners(b)])
        assert np.all(np.abs(u[:, 0]) <= l + 1e-12) and np.all(np.abs(u[:, 1]) <= w + 1e-12)
        np.testing.assert_allclose([u[:, 0].max(), -u[:, 0].min(), u[:, 1].max(), -u[:, 1].min()], [l, l, w, w], atol=1e-12)


def _step(mode):
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    step = make_step("cfg1", seed=3, n_obstacles=5, nv=7, nd=9)
    step.params = copy.deepcopy(step.params)
    step.params["modes"]["collision_check_device"] = mode
    return step


def test_loop_form_equals_vectorised_form_and_finds_collisions():
    from oracle import port, vec

    found = 0
    for mode in ("discrete", "swept"):
        step = _step(mode)
        # put one obstacle on the reference path ahead of the ego so that some candidates drive through it
        oid = next(iter(step.predictions))
        p = step.predictions[oid]
        n = len(p["pos_list"])
        px, py = step.csp.calc_position(step.c_s + 14.0)
        p["pos_list"] = np.tile(np.array([[px, py + 1.0]]), (n, 1))
        ref = vec.score_dense(step)
        _, valid, invalid, fps = port.plan_step(step)
        lv = np.array([fp.valid_level for fp in fps])
        live = ref["level"] != vec.LEVEL_SKIPPED
        np.testing.assert_array_equal(ref["level"][live], lv)
        found += int(np.sum(lv == 1))
        off = vec.score_dense(_step("host"))
        assert not np.any(off["level"] == 1)
    assert found > 0


def test_box_triangle_separating_axes_agree_with_corner_and_edge_test():
    rng = np.random.default_rng(9)
    n_hit = 0
    for _ in range(4000):
        b = boxes.box(rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-7, 7), rng.uniform(0.3, 3), rng.uniform(0.3, 1.5))
        tri = rng.uniform(-5, 5, size=(3, 2))
        m = boxes.sat_margin_triangle(b, tri)
        if abs(m) < 1e-9:
            continue
        assert (m <= 0) == boxes.box_triangle_overlap_by_polygon(b, tri)
        n_hit += m <= 0
    assert 1000 < n_hit < 3500


def test_road_boundary_loop_form_equals_vectorised_form():
    from ethicaltrajectoryplanning_b200.synthetic import make_step, road_boundary_triangles
    from oracle import port, vec

    for mode in ("discrete", "swept"):
        step = make_step("planning", seed=5, n_obstacles=4, nv=5, nd=11)
        step.road_boundary = (road_boundary_triangles(step.csp, s_end=140.0), mode)
        ref = vec.score_dense(step)
        _, _, _, fps = port.plan_step(step)
        lv = np.array([fp.valid_level for fp in fps])
        live = ref["level"] != vec.LEVEL_SKIPPED
        np.testing.assert_array_equal(ref["level"][live], lv)
        assert 0 < np.sum(lv == 2) < len(lv)
        bd = np.array([fp.bd_harm for fp in fps], dtype=np.float64)
        assert np.all((bd > 0) == (lv == 2) | ((bd > 0) & (lv < 2)))
