"""Golden vectors of the goal logic of the cost function, written by the UNMODIFIED reference
functions ``get_cost_factor`` (planner/Frenet/utils/calc_trajectory_cost.py:349-415) and ``velocity_costs`` (:438-519)
under the shims of oracle/ref_shim.py, for the planning steps of ``tests/helpers.py:goal_cases``.

pycrcc is not installable here: ``pycrcc.Point(x, y).collide(goal_area)`` is replaced by the even-odd containment test
of oracle/goal.py (goal areas are plain polygons / circles) and commonroad's ``Interval`` by a (start, end) record.  Everything else -- the branch structure, the interval
tests, the orientation folding, the time-step arithmetic -- is the reference's own code.

    python tests/golden/make_golden_goal.py        # build container only: /root/reference does not travel
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import goal as ogoal  # noqa: E402
from oracle import port, ref_shim  # noqa: E402


class _Point:
    def __init__(self, x, y):
        self.x, self.y = x, y

    def collide(self, area):
        return bool(ogoal.point_in_goal(area, self.x, self.y))


def planning_problem(g):
    """CommonRoad-shaped planning problem of a GoalContext: attributes exist only where the goal has them."""
    state = SimpleNamespace(time_step=SimpleNamespace(start=g.time_step[0], end=g.time_step[1]))
    if g.velocity is not None:
        state.velocity = SimpleNamespace(start=g.velocity[0], end=g.velocity[1])
    if g.orientation is not None:
        state.orientation = SimpleNamespace(start=g.orientation[0], end=g.orientation[1])
    states = [state]
    if g.goal_centers is not None:
        # velocity_costs reads only state_list[0].position.center (:485-488); further centres are lanelet centres in the
        # reference -- modelled here by a lanelet network of one-vertex centre lines
        if len(g.goal_centers) == 1:
            state.position = SimpleNamespace(center=np.asarray(g.goal_centers[0], dtype=np.float64))
            return SimpleNamespace(goal=SimpleNamespace(state_list=states)), None
        lanelets = {i: SimpleNamespace(center_vertices=[np.asarray(c, dtype=np.float64)]) for i, c in enumerate(g.goal_centers)}
        scenario = SimpleNamespace(lanelet_network=SimpleNamespace(find_lanelet_by_id=lambda i: lanelets[i]))
        return SimpleNamespace(goal=SimpleNamespace(state_list=states, lanelets_of_goal_position={0: list(lanelets)})), scenario
    return SimpleNamespace(goal=SimpleNamespace(state_list=states)), None


def main():
    from helpers import goal_cases

    ns = ref_shim.load_reference()
    ctc = ns.calc_trajectory_cost
    ctc.pycrcc.Point = _Point
    ctc.Interval = lambda start, end: SimpleNamespace(start=start, end=end)  # commonroad.common.util.Interval (mocked package)
    out = {}
    for name, step in goal_cases():
        g = step.goal
        fps = port.calc_frenet_trajectories(step.c_s, step.c_s_d, step.c_s_dd, step.c_d, step.c_d_d, step.c_d_dd, step.d_list,
                                            step.t_list, step.v_list, step.dt, step.csp, step.v_thr)
        problem, scenario = planning_problem(g)
        ego = SimpleNamespace(position=np.asarray(g.ego_position, dtype=np.float64), time_step=g.ego_time_step)
        area = g if g.has_area else None
        factor = [ctc.get_cost_factor(ego, fp, problem, step.dt, area) for fp in fps]
        vel = [ctc.velocity_costs(fp, ego, problem, scenario, step.dt, area) for fp in fps]
        out[f"{name}_factor"] = np.asarray(factor, dtype=np.float64)
        out[f"{name}_velocity"] = np.asarray(vel, dtype=np.float64)
        if scenario is None:
            # calc_dist_to_goal_pos (:760-803), goal-shape / survival branches; the lanelet branch measures against
            # shapely line strings (dist_to_nearest_point, :835-853), and shapely is not installable here
            dist = [ctc.calc_dist_to_goal_pos(fp, problem, None) for fp in fps]
            out[f"{name}_dist_goal"] = np.asarray(dist, dtype=np.float64)
        print(name, len(fps), {f: int(np.sum(np.asarray(factor) == f)) for f in (1.0, 0.01, 0.0001, 10.0)},
              "velocity cost %.3f .. %.3f" % (min(vel), max(vel)))
    np.savez_compressed(os.path.join(HERE, "goal_cases.npz"), **out)


if __name__ == "__main__":
    main()
