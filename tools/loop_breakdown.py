#!/usr/bin/env python
"""Where a closed-loop FrenetPlanner.step spends its wall clock (cProfile over 200 cycles)."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from ethicaltrajectoryplanning_b200.workloads import closed_loop

    closed_loop(20, grid="cfg2")
    pr = cProfile.Profile()
    pr.enable()
    lat, _ = closed_loop(200, grid="cfg2")
    pr.disable()
    import numpy as np
    print("p50 ms (under cProfile)", float(np.percentile(lat[5:], 50)) * 1e3)
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)


if __name__ == "__main__":
    main()
