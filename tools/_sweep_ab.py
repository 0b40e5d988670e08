import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ethicaltrajectoryplanning_b200 import workloads
from ethicaltrajectoryplanning_b200.scoring import ScoringContext
ctx = ScoringContext(0, "fp32")
ids = list(range(2000))
workloads.run_sweep_batched(ctx, ids, chunk=2048)
dev, wall = [], []
for _ in range(5):
    ctx.enable_timing(True)
    res, n, tp, tr = workloads.run_sweep_batched(ctx, ids, chunk=2048)
    dev.append(sum(ctx.stage_timing())); wall.append(tr * 1e3)
    ctx.enable_timing(False)
print(os.environ.get("ETP_B200_LIB", "default")[-24:], "device ms", np.round(dev, 3), "wall ms", np.round(wall, 3))
if hasattr(workloads, "run_sweep_pipelined") and hasattr(ctx, "plan_batch_pipelined"):
    for ch in (250, 500, 1000):
        t = [workloads.run_sweep_pipelined(ctx, ids, chunk=ch)[2] for _ in range(4)]
        print("pipelined chunk", ch, "scen/s", np.round(2000 / np.array(t)))
