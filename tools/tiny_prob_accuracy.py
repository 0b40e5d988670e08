#!/usr/bin/env python
"""Relative accuracy of the fp32 path's small probabilities and risks against the fp64 check mode, per decade.

The maximin gate (risk_costs.py:204-206) compares a per-obstacle risk maximum with eps = 1e-9; the selection flags a
candidate as "fp32 cannot decide" when a risk lies within kTolGate (relative) of eps.  This tool measures what fp32
actually does down there: the rho-integral form of the rectangle probability has no cancellation, so its relative
error should not grow as the value shrinks."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402


def main():
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    steps = [("cfg2", make_step("cfg2")), ("cfg3 100x200", make_step("cfg3", nv=100, nd=200)),
             ("cfg2 spread 1.5", make_step("cfg2", lateral_spread=1.5, seed=5)), ("cfg2 seed 9", make_step("cfg2", seed=9, lateral_spread=6.0))]
    edges = 10.0 ** np.arange(-16, 1)
    worst_p = np.zeros(len(edges) - 1)
    worst_r = np.zeros(len(edges) - 1)
    cnt_p = np.zeros(len(edges) - 1, dtype=np.int64)
    cnt_r = np.zeros(len(edges) - 1, dtype=np.int64)
    for name, step in steps:
        outs = {}
        for prec in ("fp32", "fp64"):
            ctx = ScoringContext(0, prec)
            outs[prec] = ctx.plan(step).numpy()
            ctx.close()
        for key, worst, cnt in (("prob", worst_p, cnt_p), ("red", worst_r, cnt_r)):
            a = outs["fp32"][key].astype(np.float64)
            b = outs["fp64"][key]
            if key == "red":
                a, b = a[:2], b[:2]
            m = b > 0
            rel = np.abs(a[m] - b[m]) / b[m]
            k = np.digitize(b[m], edges) - 1
            for i in range(len(edges) - 1):
                sel = k == i
                if sel.any():
                    worst[i] = max(worst[i], float(rel[sel].max()))
                    cnt[i] += int(sel.sum())
        print(name, "done", flush=True)
    print("decade            n(prob)  worst rel(prob)      n(risk)  worst rel(risk max)")
    for i in range(len(edges) - 1):
        print(f"[{edges[i]:.0e}, {edges[i+1]:.0e})  {cnt_p[i]:9d}  {worst_p[i]:.3e}      {cnt_r[i]:9d}  {worst_r[i]:.3e}")


if __name__ == "__main__":
    main()
