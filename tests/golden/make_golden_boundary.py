"""Golden vectors of the road-boundary handling (SURVEY.md 8f, N1 second half), written by the UNMODIFIED reference:
the whole ``sort_frenet_trajectories`` (frenet_functions.py:477-595) with a road boundary -- ``calc_risk``'s boundary
harm (risk_costs.py:104-123), ``check_validity`` level 2 (validity_checks.py:108-115, and ``boundary_valid`` :216-245
without predictions), the harm entering bayes / maximin / ego costs -- under oracle/ref_shim.py.

pycrcc is not installable here: ``RectOBB`` / ``TimeVariantCollisionObject`` are the records of
make_golden_collision.py, ``trajectories_collision_static_obstacles`` returns the first time step whose ego box overlaps
a triangle of the boundary (overlap = oracle/boxes.sat_margin_triangle), ``trajectory_preprocess_obb_sum`` reports failure
(the reference then uses the un-preprocessed object).  Pinned: the box set, ``leaving_road_at[0] - ego_state.time_step``,
``traj.v[coll_time_step]``, the LR1S formula and everything downstream of ``boundary_harm``.

    python tests/golden/make_golden_boundary.py        # build container only: /root/reference does not travel
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import boxes, ref_shim  # noqa: E402


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(HERE, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def boundary_cases():
    """(name, PlanningStep with .road_boundary, ego time step)."""
    from ethicaltrajectoryplanning_b200.synthetic import make_step, road_boundary_triangles

    out = []
    for name, kw, hw, ts in (("planning", dict(seed=31, n_obstacles=4, nv=7, nd=13), 3.6, 0),
                             ("narrow_road", dict(seed=32, n_obstacles=3, nv=6, nd=13, v0=20.0), 2.6, 17),
                             ("two_horizons", dict(seed=33, n_obstacles=3, nv=6, nd=11, t_list=(2.0, 3.0)), 3.2, 4),
                             ("no_predictions", dict(seed=34, n_obstacles=0, nv=6, nd=13), 3.0, 9)):
        s = make_step("planning", **kw)
        if name == "no_predictions":
            # without predictions the reference only runs in mode "ground_truth" (max_risk_valid reads risk dicts that
            # calc_risk never filled otherwise); collision_valid then asks the scenario's collision checker
            s.predictions = None
            s.mode = "ground_truth"
        s.road_boundary = (road_boundary_triangles(s.csp, half_width=hw, s_end=160.0), "discrete")
        out.append((name, s, ts))
    return out


def main():
    mg, mc = _load("make_golden"), _load("make_golden_collision")
    ns = ref_shim.load_reference()

    def static_query(trajectories, static_obstacles, method=None, num_cells=None, auto_orientation=None):
        res = []
        for ego in trajectories:
            first = -1
            for k, o in enumerate(ego.items):
                if any(not boxes.sat_margin_triangle(o.box, tri) > 0.0 for tri in static_obstacles):
                    first = ego.start + k
                    break
            res.append(first)
        return res

    for mod_name in list(sys.modules):
        mod = sys.modules[mod_name]
        if mod_name.endswith("helper_functions"):
            mod.pycrcc.RectOBB = mc.RectOBB
            mod.pycrcc.TimeVariantCollisionObject = mc.TimeVariantCollisionObject
        if mod_name.endswith("validity_checks") or mod_name.endswith("risk_costs"):
            mod.trajectory_queries.trajectory_preprocess_obb_sum = lambda obj: (obj, 1)
            mod.trajectory_queries.trajectories_collision_static_obstacles = static_query
    out = {}
    for name, step, time_step in boundary_cases():
        tris = np.asarray(step.road_boundary[0]).reshape(-1, 3, 2)
        from types import SimpleNamespace

        fp_list, valid, invalid, best, _ = mg.run_reference(step, road_boundary=tris, time_step=time_step,
                                                            collision_checker=SimpleNamespace(collide=lambda obj: False))
        out[f"{name}_bd"] = np.array([float(getattr(fp, "bd_harm", 0.0)) for fp in fp_list])  # not set without predictions
        out[f"{name}_level"] = np.array([fp.valid_level for fp in fp_list], dtype=np.int64)
        out[f"{name}_reason"] = np.array([mg.REASONS.index(fp.reason_invalid) for fp in fp_list], dtype=np.int64)
        out[f"{name}_cost"] = np.array([float(fp.cost) for fp in fp_list])
        out[f"{name}_best"] = np.int64(fp_list.index(best))
        print(name, len(fp_list), "levels", dict(zip(*np.unique(out[f"{name}_level"], return_counts=True))),
              "bd > 0:", int(np.sum(out[f"{name}_bd"] > 0)), "best", int(out[f"{name}_best"]))
    np.savez_compressed(os.path.join(HERE, "boundary_cases.npz"), **out)


if __name__ == "__main__":
    main()
