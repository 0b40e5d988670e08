import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ethicaltrajectoryplanning_b200 import scoring, workloads
hist = collections.Counter()
orig = scoring.ScoringContext.plan_cycle
last = [0]
rows = []
def pc(self, step, raw=None):
    r = orig(self, step, raw)
    n = self.debug_counters()["rescored"]
    hist[n - last[0]] += 1
    rows.append((n - last[0], r[0]["index"], r[0]["level"], r[0]["cost"]))
    last[0] = n
    return r
scoring.ScoringContext.plan_cycle = pc
workloads.closed_loop(200, "weights_ethical.json", sys.argv[1] if len(sys.argv) > 1 else "cfg1")
print(sorted(hist.items()))
print(rows[:12])
