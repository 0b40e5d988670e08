import sys, time
sys.path.insert(0, '/root/repo')
from ethicaltrajectoryplanning_b200 import dropin, workloads
ctx = dropin.default_context()
ids = list(range(1000))
workloads.run_sweep_batched(ctx, ids[:32])
res, n, tp, tr = workloads.run_sweep_batched(ctx, ids)
print(f"batched: {len(ids)} scenarios {n} cand: pack {tp:.3f}s run {tr:.3f}s -> {len(ids)/tr:.0f} scen/s (device+C), {len(ids)/(tp+tr):.0f} scen/s incl. python packing")
one, n1, sec = workloads.run_sweep(ctx, ids)
print(f"loop: {len(ids)/sec:.0f} scen/s")
print("equal", all(a[:3]==b[:3] for a,b in zip(one,res)))
