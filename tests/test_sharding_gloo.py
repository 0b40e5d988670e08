"""N > 1 host logic on CPU: profile-aligned candidate shards, exchange of the 40-byte best records
over a world_size-2 gloo group, lexicographic combine == unsharded selection (oracle as the scorer)."""
import ctypes
import os

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from ethicaltrajectoryplanning_b200 import _abi
from ethicaltrajectoryplanning_b200 import distributed as D
from helpers import load_golden


def test_profile_shards_partition_the_dense_range():
    for n_prof, nd, world in [(1000, 1000, 8), (7, 5, 2), (5, 3, 4), (3, 9, 8)]:
        edges = [D.profile_shard(n_prof, nd, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n_prof * nd
        for (a0, a1), (b0, b1) in zip(edges[:-1], edges[1:]):
            assert a1 == b0 and a0 % nd == 0 and a1 % nd == 0
    assert D.scenario_shard(10, 1, 4) == [1, 5, 9]


def test_key_orders_by_level_then_cost_then_index():
    k = D.key_hi_from
    assert k(10, 5.0) < k(3, 0.0) < k(2, -1.0) < k(1, 0.0) < k(0, -100.0)
    assert k(10, -2.0) < k(10, -1.0) < k(10, 0.0) < k(10, 1e-300) < k(10, 3.0) < k(10, float("inf"))
    assert k(10, float("nan")) > k(10, float("inf"))


def _rec(index, level, cost, n_scored, list_index):
    b = _abi.Best()
    b.key_hi, b.key_lo = D.key_from(level, cost, max(index, 0))
    b.cost, b.index, b.level, b.n_scored, b.list_index = cost, index, level, n_scored, list_index
    if index < 0:
        b.key_hi = b.key_lo = (1 << 64) - 1
    return b


def test_c_combine_matches_python_combine():
    _abi.build_library()
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(1, 9))
        recs, base = [], 0
        for r in range(n):
            cnt = int(rng.integers(0, 6))
            if cnt == 0:
                recs.append(_rec(-1, -1, 0.0, 0, -1))
            else:
                li = int(rng.integers(0, cnt))
                recs.append(_rec(base + li, int(rng.choice([0, 1, 2, 3, 10])), float(rng.choice([0.0, 1.5, -2.0, 7.25])), cnt, li))
            base += 10
        rc, out = D.combine_best_records(recs, consecutive_ranges=True)
        py = D.python_combine([dict(index=b.index, level=b.level, cost=b.cost, n_scored=b.n_scored,
                                    list_index=b.list_index, key_hi=b.key_hi, key_lo=b.key_lo) for b in recs], consecutive_ranges=True)
        if py is None:
            assert rc == _abi.ERR_NO_CANDIDATE
        else:
            assert rc == 0 and (out.index, out.level, out.n_scored, out.list_index) == (
                py["index"], py["level"], py["n_scored"], py["list_index"])


def _worker(rank, world, port, case, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import bvn, port as oracle_port

        step, ref = load_golden(case)
        n_prof, nd = len(step.v_list) * len(step.t_list), len(step.d_list)
        c0, c1 = D.profile_shard(n_prof, nd, rank, world)
        # score only this rank's profiles with the oracle
        import copy

        sub = copy.copy(step)
        nt = len(step.t_list)
        sub.v_list = step.v_list[c0 // nd // nt:c1 // nd // nt]
        try:
            best, valid, invalid, allfp = oracle_port.plan_step(sub, mvnun=bvn.mvnun_fast)
            li = allfp.index(best)
            rec = _rec(c0 + best.dense_index, best.valid_level, float(best.cost), len(allfp), li)
        except ValueError:
            rec = _rec(-1, -1, 0.0, 0, -1)
        recs = D.gather_best_records(np.frombuffer(bytes(rec), dtype=np.uint8), dist)
        rc, out = D.combine_best_records(recs, consecutive_ranges=True)
        q.put((rank, rc, out.index, out.level, out.list_index, out.n_scored))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["cfg1_ethical", "lowspeed_0p5", "max_risk_level3"])
def test_world_size_2_gloo_argmin_equals_unsharded(case):
    _abi.build_library()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    step, ref = load_golden(case)
    for rank, rc, index, level, list_index, n_scored in res:
        assert rc == 0
        assert list_index == int(ref["best"]), (rank, list_index, int(ref["best"]))
        assert level == int(ref["level"][int(ref["best"])])
        assert n_scored == len(ref["nT"])
