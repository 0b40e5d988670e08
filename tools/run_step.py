#!/usr/bin/env python
"""Run a few planning steps of a synthetic workload (used under ncu)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--nv", type=int, default=None)
    ap.add_argument("--nd", type=int, default=None)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--argmin-only", action="store_true")
    ap.add_argument("--lateral-spread", type=float, default=4.0)
    ap.add_argument("--collision", default="host", choices=["host", "discrete", "swept"])
    ap.add_argument("--boundary", default="host", choices=["host", "discrete", "swept"], help="synthetic road boundary on the device")
    ap.add_argument("--goal", action="store_true", help="attach a goal context (rectangle ahead on the path)")
    args = ap.parse_args()
    import torch

    from ethicaltrajectoryplanning_b200.scoring import ScoringContext
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    step = make_step(args.workload, nv=args.nv, nd=args.nd, lateral_spread=args.lateral_spread)
    if args.collision != "host":
        import copy
        step.params = copy.deepcopy(step.params)
        step.params["modes"]["collision_check_device"] = args.collision
    if args.boundary != "host":
        from ethicaltrajectoryplanning_b200.synthetic import road_boundary_triangles
        step.road_boundary = (road_boundary_triangles(step.csp), args.boundary)
    if args.goal:
        from ethicaltrajectoryplanning_b200.synthetic import GoalContext
        step.goal = GoalContext(polygons=[[(24.0, 0.5), (60.0, 0.5), (60.0, 4.2), (24.0, 4.2)]], velocity=(5.0, 30.0),
                                orientation=(-0.5, 0.5), time_step=(10, 60), ego_time_step=0,
                                ego_position=tuple(step.ego_position), goal_centers=[(42.0, 2.3)])
    ctx = ScoringContext(0, args.precision)
    ctx.configure(step)
    s, keep, nsamp = ctx.make_step_struct(step)
    bufs = None if args.argmin_only else ctx.allocate_outputs(step.n_dense, int(nsamp.max()), ctx.O)
    ctx.enable_timing(True)
    for i in range(args.steps):
        ctx.launch(s, bufs)
        print("step", i, "profile_ms %.4f score_ms %.4f select_ms %.4f" % ctx.stage_timing(), ctx.best())
    if bufs is not None:
        print("nonzero prob fraction", float((bufs["prob"] != 0).float().mean()))
        print("level histogram", torch.bincount(bufs["level"].to(torch.int64), minlength=11)[[0, 1, 2, 3, 10]].tolist())
    torch.cuda.synchronize()
    print("debug counters", ctx.debug_counters())
    ctx.close()


if __name__ == "__main__":
    main()
