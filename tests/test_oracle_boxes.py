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


def _collision_cases():
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location(
        "_mgc", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden_collision.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "collision_cases.npz"))
    return [(name, step, z[name].astype(bool)) for name, step, _ in mod.collision_cases()]


def test_discrete_mode_matches_the_reference_s_own_call_sequence():
    """tests/golden/collision_cases.npz was written by the reference's create_collision_object +
    collision_checker_prediction (fallback branch) with stand-in pycrcc primitives: box set, half extents, orientations,
    time alignment and pred_length are the reference's."""
    from oracle import port, vec

    n_hit = n_all = 0
    for name, step, want in _collision_cases():
        fps = port.calc_frenet_trajectories(step.c_s, step.c_s_d, step.c_s_dd, step.c_d, step.c_d_d, step.c_d_dd, step.d_list,
                                            step.t_list, step.v_list, step.dt, step.csp, step.v_thr)
        got = np.array([boxes.collides_with_predictions(fp.x, fp.y, fp.yaw, step.predictions, step.vehicle.l, step.vehicle.w,
                                                        "discrete") for fp in fps])
        np.testing.assert_array_equal(got, want, err_msg=name)
        step.params = copy.deepcopy(step.params)
        step.params["modes"]["collision_check_device"] = "discrete"
        ref = vec.score_dense(step)
        live = ref["level"] != vec.LEVEL_SKIPPED
        lvl = ref["level"][live]
        # level 0 (velocity / curvature) is decided before the collision test
        np.testing.assert_array_equal(lvl[lvl > 0] == 1, want[lvl > 0], err_msg=name)
        n_hit += int(want.sum())
        n_all += len(want)
    assert 0.3 < n_hit / n_all < 0.8


def _boundary_cases():
    import importlib.util
    import os

    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("_mgb", os.path.join(here, "golden", "make_golden_boundary.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    z = np.load(os.path.join(here, "golden", "boundary_cases.npz"))
    return [(name, step, {k: z[f"{name}_{k}"] for k in ("bd", "level", "reason", "cost", "best")}) for name, step, _ in mod.boundary_cases()]


def test_road_boundary_matches_the_reference_s_own_sort():
    """tests/golden/boundary_cases.npz: the reference's whole sort_frenet_trajectories with a road boundary (stand-in
    pycrcc primitives): boundary harm, level 2, costs, selection."""
    from oracle import port, vec

    for name, step, want in _boundary_cases():
        best, valid, invalid, fps = port.plan_step(step)
        np.testing.assert_array_equal([fp.valid_level for fp in fps], want["level"], err_msg=name)
        np.testing.assert_array_equal([port.REASONS.index(fp.reason_invalid) for fp in fps], want["reason"], err_msg=name)
        if step.predictions is not None:
            np.testing.assert_allclose([float(fp.bd_harm) for fp in fps], want["bd"], rtol=1e-13, atol=0, err_msg=name)
        np.testing.assert_allclose([float(fp.cost) for fp in fps], want["cost"], rtol=1e-12, atol=1e-14, err_msg=name)
        assert fps.index(best) == int(want["best"])
        ref = vec.score_dense(step)
        live = ref["level"] != vec.LEVEL_SKIPPED
        np.testing.assert_array_equal(ref["level"][live], want["level"], err_msg=name)
        assert ref["best"]["list_index"] == int(want["best"])
