import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ethicaltrajectoryplanning_b200.scoring import ScoringContext
from ethicaltrajectoryplanning_b200.synthetic import make_step
from oracle import vec
step = make_step("cfg3")
ctx = ScoringContext(0, "fp32")
res = ctx.plan(step)
rng = np.random.default_rng(5)
profiles = sorted(rng.choice(len(step.v_list) * len(step.t_list), size=3, replace=False).tolist())
ref = vec.score_dense(step, profiles=profiles)
sel = torch.as_tensor(ref["dense_index"], device=res.level.device)
prob = res.prob[:, :, sel].cpu().numpy().transpose(2, 0, 1).astype(np.float64)
r = ref["prob"]
bad = ~(np.abs(prob - r) <= 1e-4 * np.abs(r) + 1e-9)
print("profiles", profiles, "bad", bad.sum(), "of", bad.size, "nonzero", (r != 0).sum())
idx = np.argwhere(bad)
rel = np.abs(prob - r)[bad] / np.maximum(np.abs(r[bad]), 1e-300)
order = np.argsort(-rel)[:25]
oids = list(step.predictions.keys())
for k in order:
    c, t, o = idx[k]
    cov = step.predictions[oids[o]]["cov_list"][t]
    print(f"c={c} t={t} o={o} got={prob[c,t,o]:.6e} ref={r[c,t,o]:.6e} rel={rel[k]:.2e} cov=({cov[0,0]:.3f},{cov[0,1]:.3f},{cov[1,1]:.3f})")
print("by obstacle:", np.bincount(idx[:, 2], minlength=32))
print("by t:", np.bincount(idx[:, 1], minlength=49))
gz = (prob == 0) != (r == 0)
print("gate mismatches", gz.sum())
