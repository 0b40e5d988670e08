#!/usr/bin/env python
"""cfg4 (scenario sweep) and cfg5 (closed loop) on one GPU: prints scenarios/s and step-latency percentiles."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402


def main():
    from ethicaltrajectoryplanning_b200 import dropin, workloads

    ctx = dropin.default_context()
    ids = list(range(int(sys.argv[1]) if len(sys.argv) > 1 else 200))
    workloads.run_sweep(ctx, ids[:10])
    res, n_cand, sec = workloads.run_sweep(ctx, ids)
    print(f"cfg4: {len(ids)} scenarios, {n_cand} candidates in {sec:.3f} s -> {len(ids) / sec:.1f} scenarios/s, "
          f"{n_cand / sec:.0f} candidates/s, {1e3 * sec / len(ids):.3f} ms/scenario")
    for grid in ("cfg1", "cfg2"):
        for w in ("weights_ethical.json", "weights_ego.json", "weights_standard.json"):
            lat, states = workloads.closed_loop(200, w, grid)
            lat = lat[5:] * 1e3
            print(f"cfg5 {grid} {w}: p50 {np.percentile(lat, 50):.3f} ms p99 {np.percentile(lat, 99):.3f} ms  "
                  f"final s={states[-1][0]:.1f} d={states[-1][1]:.2f} v={states[-1][2]:.2f}")
    # where the time of one closed-loop step goes
    import cProfile
    import pstats

    pr = cProfile.Profile()
    pr.enable()
    workloads.closed_loop(100, "weights_ethical.json", "cfg2")
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)


if __name__ == "__main__":
    main()
