import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ethicaltrajectoryplanning_b200 import scoring, workloads
from ethicaltrajectoryplanning_b200.planner.Frenet import frenet_planner as fp
keep = {}
orig = fp.FrenetPlanner._step_fused
def sf(self, *a):
    r = orig(self, *a)
    keep["pl"] = self
    return r
fp.FrenetPlanner._step_fused = sf
workloads.closed_loop(150, "weights_ethical.json", "cfg1")
pl = keep["pl"]
step = pl.last_step
print("v_list", repr(step.v_list), "d_list", step.d_list, "t", step.t_list)
import ctypes as C
pl.ctx._check(pl.ctx.lib.etp_set_raw_predictions(pl.ctx.h, C.byref(pl._raw_record), pl.ctx.precision))
res = pl.ctx.plan(step, configure=False)
out = res.numpy()
cost = out["cost"][12]; lvl = out["level"]
order = np.argsort(cost)
for c in order[:6]:
    print(c, lvl[c], repr(float(cost[c])), "v", step.v_list[c // (len(step.t_list) * len(step.d_list))], "d", step.d_list[c % len(step.d_list)])
print("best", res.best)
