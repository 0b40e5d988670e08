#!/usr/bin/env python
"""Host-side cost of packing a scenario sweep for etp_plan_batch (cProfile)."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from ethicaltrajectoryplanning_b200 import workloads
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext

    ctx = ScoringContext(0, "fp32")
    ids = list(range(400))
    steps = [st for i in ids for _, st in workloads.sweep_steps(1, i)]
    ctx.set_params(steps[0].params, steps[0].vehicle, steps[0].mode)
    ctx.plan_batch(packed=ctx.pack_scenarios(steps[:32]))
    pr = cProfile.Profile()
    pr.enable()
    packs = [ctx.pack_scenarios(steps[i:i + 256]) for i in range(0, len(steps), 256)]
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(22)


if __name__ == "__main__":
    main()
