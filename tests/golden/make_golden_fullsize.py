"""Full-size selection fixtures of BASELINE.json configs[1] / configs[2] (cfg2: 10^4, cfg3: 10^6 candidates).

Runs the vectorised fp64 oracle (oracle/vec.py, pinned to the reference's own fixtures by tests/test_oracle_vec.py)
over EVERY (v,t) profile of the synthetic step, in parallel over profiles, and keeps what a selection test needs:

  winner            dense index, list index, level, cost of the selected trajectory
                    (frenet_functions.py:562-593, frenet_planner.py:457,555-558)
  level_hist        [P][6] candidates of each profile at level 0 / 1 / 2 / 3 / 10 / skipped
  level_digest      [P] sum over the profile's lateral targets of (id + 1) * (level + 1): catches a single flipped mask
  min_cost, arg_min [P][5] smallest total cost and its dense index per profile and level rank (inf / -1 if empty)
  top_index, top_cost   the 4096 cheapest candidates of the highest non-empty level, in (cost, index) order

    python tests/golden/make_golden_fullsize.py cfg3 [--procs 8] [--dump gpurun_out/cfg3_oracle_costs.npz]

cfg3 takes about five minutes on eight cores.  Needs only numpy / scipy (the oracle), not /root/reference.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RANKS = {0: 0, 1: 1, 2: 2, 3: 3, 10: 4}
_STEP = None


def _init(name):
    global _STEP
    from ethicaltrajectoryplanning_b200.synthetic import make_step

    _STEP = make_step(name)


def _score(profiles):
    from oracle import vec

    step = _STEP
    nd = len(step.d_list)
    out = []
    for p in profiles:
        try:
            r = vec.score_dense(step, profiles=[p])
            level, cost, red = r["level"], r["cost"], r["red"]
        except ValueError:  # every lateral target of the profile skipped
            level = np.full(nd, vec.LEVEL_SKIPPED, dtype=np.int64)
            cost = np.zeros((len(vec.COST_ROWS), nd))
            red = np.zeros((4, nd, 0))
        out.append((p, level.astype(np.uint8), cost[12].copy(), cost[5].copy(), red[0].sum(axis=1), red[1].sum(axis=1)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=["cfg2", "cfg3"])
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--dump", default=None, help="also write every candidate's level / cost (scratch, not committed)")
    ap.add_argument("--top", type=int, default=4096)
    args = ap.parse_args()
    import multiprocessing as mp

    _init(args.workload)
    step = _STEP
    nv, nt, nd = len(step.v_list), len(step.t_list), len(step.d_list)
    P = nv * nt
    t0 = time.time()
    chunks = [list(range(i, min(i + 4, P))) for i in range(0, P, 4)]
    level = np.zeros((P, nd), dtype=np.uint8)
    total = np.zeros((P, nd))
    risk = np.zeros((P, nd))
    sum_e = np.zeros((P, nd))
    sum_o = np.zeros((P, nd))
    with mp.get_context("fork").Pool(args.procs, initializer=_init, initargs=(args.workload,)) as pool:
        done = 0
        for res in pool.imap_unordered(_score, chunks):
            for p, lv, c, rk, se, so in res:
                level[p], total[p], risk[p], sum_e[p], sum_o[p] = lv, c, rk, se, so
            done += len(res)
            if done % 100 < 4:
                print(f"{done}/{P} profiles, {time.time() - t0:.0f} s", flush=True)
    lv = level.astype(np.int64)
    skipped = lv == 255
    hist = np.stack([(lv == k).sum(axis=1) for k in (0, 1, 2, 3, 10, 255)], axis=1).astype(np.int32)
    digest = (np.where(skipped, 0, lv + 1) * (np.arange(nd)[None, :] + 1)).sum(axis=1).astype(np.int64)
    min_cost = np.full((P, 5), np.inf)
    arg_min = np.full((P, 5), -1, dtype=np.int64)
    for k, r in RANKS.items():
        m = lv == k
        c = np.where(m, total, np.inf)
        j = c.argmin(axis=1)  # first of equal costs = generation order
        has = m.any(axis=1)
        min_cost[has, r] = c[has, j[has]]
        arg_min[has, r] = (np.arange(P) * nd + j)[has]
    rank = np.full(lv.shape, -1)
    for k, r in RANKS.items():
        rank[lv == k] = r
    top_rank = rank.max()
    flat_cost = np.where(rank == top_rank, total, np.inf).ravel()
    order = np.lexsort((np.arange(flat_cost.size), flat_cost))[:args.top]
    order = order[np.isfinite(flat_cost[order])]
    win = int(order[0])
    out = dict(
        workload=args.workload, nv=nv, nt=nt, nd=nd,
        winner_index=win, winner_cost=float(flat_cost[win]), winner_level=int(lv.ravel()[win]),
        winner_list_index=int(np.count_nonzero(~skipped.ravel()[:win])), n_scored=int((~skipped).sum()),
        level_hist=hist, level_digest=digest, min_cost=min_cost, arg_min=arg_min,
        top_index=order.astype(np.int64), top_cost=flat_cost[order],
    )
    path = os.path.join(HERE, f"fullsize_{args.workload}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "winner", win, flat_cost[win], "runner-up", int(order[1]), flat_cost[order[1]],
          f"{time.time() - t0:.0f} s")
    if args.dump:
        np.savez_compressed(args.dump, level=level, total=total, risk_total=risk, sum_e=sum_e, sum_o=sum_o)


if __name__ == "__main__":
    main()
