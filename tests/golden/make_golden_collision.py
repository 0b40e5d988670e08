"""Golden vectors of the collision test with the predicted obstacle boxes, written by the
UNMODIFIED reference functions ``create_collision_object`` (planner/Frenet/utils/validity_checks.py:125-141) and
``collision_checker_prediction`` (planner/Frenet/utils/prediction_helpers.py:138-210) under oracle/ref_shim.py.

pycrcc is not installable here, so its three primitives are replaced by stand-ins: ``RectOBB`` (a record),
``TimeVariantCollisionObject`` (start index + list of boxes), ``trajectories_collision_dynamic_obstacles`` (first time
step at which the boxes that BOTH objects hold for that step overlap; overlap = oracle/boxes.sat_margin).
``trajectory_preprocess_obb_sum`` reports failure, which sends the reference down its own fallback branch (the
un-preprocessed objects).  What the fixtures pin is therefore everything the reference's sources decide: which boxes
exist, their half extents and orientations, their time steps, ``pred_length``, the early exit over obstacles.

    python tests/golden/make_golden_collision.py        # build container only: /root/reference does not travel
"""
from __future__ import annotations

import importlib
import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import boxes, port, ref_shim  # noqa: E402


class RectOBB(ref_shim._RectOBB):
    """ref_shim's record (center / r_x / local_x_axis, what collision_probability.py:349-362 reads) plus the box tuple."""

    def __init__(self, r_x, r_y, orientation, x, y):
        super().__init__(r_x, r_y, orientation, x, y)
        self.box = boxes.box(x, y, orientation, r_x, r_y)


class TimeVariantCollisionObject:
    def __init__(self, time_start_idx):
        self.start, self.items = int(time_start_idx), []

    def append_obstacle(self, o):
        self.items.append(o)

    def at(self, t):
        k = t - self.start
        return self.items[k] if 0 <= k < len(self.items) else None


def trajectories_collision_dynamic_obstacles(trajectories, dynamic_obstacles, method=None, num_cells=None):
    out = []
    for ego in trajectories:
        first = -1
        for t in range(ego.start, ego.start + len(ego.items)):
            if any(o.at(t) is not None and not boxes.sat_margin(ego.at(t).box, o.at(t).box) > 0.0 for o in dynamic_obstacles):
                first = t
                break
        out.append(first)
    return out


def collision_cases():
    """(name, PlanningStep, ego time step): obstacles moved onto the path so that some candidates hit them."""
    import importlib.util

    from ethicaltrajectoryplanning_b200.synthetic import make_step

    spec = importlib.util.spec_from_file_location("_scenes", os.path.join(os.path.dirname(HERE), "test_gpu_random_scenes.py"))
    scenes = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(scenes)
    out = []
    for seed in (0, 1, 3, 6, 13, 22, 29):  # 0: every candidate collides; the others: between 30 and 68 of ~78
        s = scenes.random_scene(seed)
        s.v_list = np.ascontiguousarray(s.v_list[::2])
        s.d_list = np.linspace(-3.0, 3.0, 13)
        out.append((f"scene_{seed:02d}", s, 3 * seed))
    s = make_step("planning", seed=21, n_obstacles=6, nv=7, nd=13, t_list=(2.0, 3.0))
    out.append(("two_horizons", s, 0))
    s = make_step("planning", seed=22, n_obstacles=5, nv=7, nd=13, n_pred=9)   # predictions shorter than the candidates
    oid = list(s.predictions)[0]
    px, py = s.csp.calc_position(s.c_s + 9.0)
    s.predictions[oid]["pos_list"] = np.tile(np.array([[px, py + 1.2]]), (len(s.predictions[oid]["pos_list"]), 1))
    out.append(("short_predictions", s, 11))
    return out


def main():
    ns = ref_shim.load_reference()
    ph = importlib.import_module("EthicalTrajectoryPlanning.planner.Frenet.utils.prediction_helpers")
    vc = ns.validity_checks
    for mod_name in list(sys.modules):
        if mod_name.endswith("helper_functions"):
            hf = sys.modules[mod_name]
            hf.pycrcc.RectOBB = RectOBB
            hf.pycrcc.TimeVariantCollisionObject = TimeVariantCollisionObject
    failed = lambda obj: (obj, 1)  # noqa: E731  -> "if err: use the raw trajectory"
    for mod in (ph, vc):
        mod.trajectory_queries.trajectory_preprocess_obb_sum = failed
        mod.trajectory_queries.trajectories_collision_dynamic_obstacles = trajectories_collision_dynamic_obstacles
    out = {}
    for name, step, time_step in collision_cases():
        fps = port.calc_frenet_trajectories(step.c_s, step.c_s_d, step.c_s_dd, step.c_d, step.c_d_d, step.c_d_dd, step.d_list,
                                            step.t_list, step.v_list, step.dt, step.csp, step.v_thr)
        ego = SimpleNamespace(time_step=time_step)
        scenario = SimpleNamespace(obstacle_by_id=lambda oid: SimpleNamespace(obstacle_shape=SimpleNamespace(length=1.0)))
        hits = []
        for fp in fps:
            co = vc.create_collision_object(fp, step.vehicle, ego)
            hits.append(bool(ph.collision_checker_prediction(predictions=step.predictions, scenario=scenario, ego_co=co,
                                                             frenet_traj=fp, ego_state=ego)))
        out[name] = np.asarray(hits, dtype=np.uint8)
        print(name, len(fps), "colliding:", int(np.sum(hits)))
    np.savez_compressed(os.path.join(HERE, "collision_cases.npz"), **out)


if __name__ == "__main__":
    main()
