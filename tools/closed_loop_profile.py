"""Where a closed-loop cycle's wall time goes: FrenetPlanner.step in total, ScoringContext.plan_cycle inside it, and the
etp_plan_cycle library call inside that (ETP_CYCLE_PROFILE=1 adds the library's own split on stderr).  GPU box only.

    python tools/closed_loop_profile.py [cfg1|cfg2] [n_cycles]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ethicaltrajectoryplanning_b200 import scoring, workloads  # noqa: E402


class _TimedLib:
    def __init__(self, lib, sink):
        self._lib, self._sink = lib, sink

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if name != "etp_plan_cycle":
            return fn

        def timed(*a):
            t0 = time.perf_counter()
            r = fn(*a)
            self._sink.append(time.perf_counter() - t0)
            return r

        return timed


def main():
    grid = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    t_cycle, t_lib, ctxs = [], [], []
    orig = scoring.ScoringContext.plan_cycle

    def plan_cycle(self, step, raw=None):
        if not ctxs:
            ctxs.append(self)
            self.lib = _TimedLib(self.lib, t_lib)
        t0 = time.perf_counter()
        r = orig(self, step, raw)
        t_cycle.append(time.perf_counter() - t0)
        return r

    scoring.ScoringContext.plan_cycle = plan_cycle
    lat, _ = workloads.closed_loop(n, "weights_ethical.json", grid)
    p50 = lambda a: np.percentile(np.asarray(a[5:]) * 1e6, 50) if len(a) > 5 else float("nan")  # noqa: E731
    print(f"{grid}: step p50 {p50(lat):.1f} us, plan_cycle p50 {p50(t_cycle):.1f} us, etp_plan_cycle p50 {p50(t_lib):.1f} us, "
          f"cycles {len(t_cycle)}, counters {ctxs[0].debug_counters() if ctxs else None}")


if __name__ == "__main__":
    main()
