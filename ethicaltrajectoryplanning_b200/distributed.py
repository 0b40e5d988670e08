"""Multi-GPU partitioning of the scoring path (one process per GPU, torch.distributed).

Two shardings, as the path offers them:

* candidate sharding inside one planning step: whole (v,t) profiles of the dense (v,t,d) index
  space are dealt to the ranks round-robin (profile p on rank ``p % world``: neighbouring
  profiles cost about the same, so the shards do too); every rank scores its profiles with the
  fused kernel
  and the shard-local arg-min records (40 bytes, level|cost|index key) are exchanged with one
  all-gather and reduced lexicographically -- the "min-reduce of (cost, index)".  The validity
  level is part of the key, so the reference's "highest non-empty level" rule
  (frenet_functions.py:562-564) holds globally;
* scenario sharding for sweeps: scenario i goes to rank ``i % world`` (the reference's
  ``Pool(10).imap_unordered`` over scenarios, planner/plannertools/evaluate.py:160-169), no
  data-path collective.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi


def profile_shard(n_profiles: int, nd: int, rank: int, world: int):
    """Dense candidate range [c_begin, c_end) of ``rank`` when the profiles are cut into contiguous blocks instead (the
    C ABI's ``c_begin`` / ``c_end``; ``ShardedPlanner`` interleaves, see ``etp_step.shard_index``)."""
    base, rem = divmod(n_profiles, world)
    p0 = rank * base + min(rank, rem)
    p1 = p0 + base + (1 if rank < rem else 0)
    return p0 * nd, p1 * nd


def scenario_shard(n_scenarios: int, rank: int, world: int):
    """Round-robin scenario ids of ``rank``."""
    return list(range(rank, n_scenarios, world))


def combine_best_records(records, consecutive_ranges=False):
    """Lexicographic min over shard-local etp_best records (interleaved shards of one range, or
    consecutive candidate ranges given in dense order)."""
    n = len(records)
    arr = (_abi.Best * n)(*records)
    out = _abi.Best()
    lib = _abi.load_library()
    rc = lib.etp_combine_best(arr, n, int(bool(consecutive_ranges)), C.byref(out))
    return rc, out


def records_from_bytes(buf: np.ndarray, n: int):
    size = C.sizeof(_abi.Best)
    raw = np.ascontiguousarray(buf, dtype=np.uint8).tobytes()
    return [_abi.Best.from_buffer_copy(raw[i * size:(i + 1) * size]) for i in range(n)]


def key_from(level: int, cost: float, index: int):
    """Host twin of the device key (etp_device.cuh: make_key_hi / make_key_lo): (key_hi, key_lo) orders like
    (highest level first, cost, index) with the full 64-bit image of the cost; smaller wins."""
    import struct

    rank = 4 if level == 10 else level
    if cost != cost:
        mono = (1 << 64) - 1
    else:
        b = struct.unpack("<Q", struct.pack("<d", float(cost)))[0]
        mono = (~b & ((1 << 64) - 1)) if (b >> 63) else (b | (1 << 63))
    return ((4 - rank) << 61) | (mono >> 3), ((mono & 7) << 61) | int(index)


def key_hi_from(level: int, cost: float) -> int:
    return key_from(level, cost, 0)[0]


def gather_best_records(local: "np.ndarray", dist, group=None, device=None):
    """all_gather of one 40-byte etp_best record per rank (works on NCCL and gloo tensors)."""
    import torch

    world = dist.get_world_size(group)
    t = torch.as_tensor(np.ascontiguousarray(local, dtype=np.uint8))
    if device is not None:
        t = t.to(device)
    out = torch.empty((world * t.numel(),), dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return records_from_bytes(out.cpu().numpy(), world)


def python_combine(records, consecutive_ranges=False):
    """Pure-Python twin of etp_combine_best for hosts without the CUDA library (gloo tests)."""
    win, scored, before = -1, 0, 0
    for i, b in enumerate(records):
        if b["index"] >= 0 and (win < 0 or (b["key_hi"], b["key_lo"]) < (records[win]["key_hi"], records[win]["key_lo"])):
            win, before = i, scored
        scored += b["n_scored"]
    if win < 0:
        return None
    out = dict(records[win])
    out["n_scored"] = scored
    out["list_index"] = (before if consecutive_ranges else 0) + records[win]["list_index"]
    return out


class ShardedPlanner:
    """Candidate-sharded planning step over the ranks of a process group."""

    def __init__(self, ctx, group=None):
        import torch.distributed as dist

        self.ctx = ctx
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._gather = None

    def make_step_struct(self, step):
        """Step struct of this rank: whole dense range, profiles interleaved over the ranks."""
        return self.ctx.make_step_struct(step, shard_index=self.rank, shard_count=self.world)

    def launch(self, s, bufs=None):
        """Local fused launch, all-gather of the 40-byte best records and their combination on the device: everything is
        asynchronous on the context's stream; the global selection replaces the context's best record."""
        torch = self.ctx.torch
        self.ctx.launch(s, bufs)
        local = self.ctx.best_device_tensor()
        if self._gather is None:
            self._gather = torch.empty((self.world * local.numel(),), dtype=torch.uint8, device=local.device)
        with torch.cuda.stream(self.ctx.stream):  # ordered after the selection kernels on the context's stream
            self.dist.all_gather_into_tensor(self._gather, local, group=self.group)
        self.ctx._check(self.ctx.lib.etp_combine_best_device(self.ctx.h, C.c_void_p(self._gather.data_ptr()), self.world, 0))
        return self._gather

    def best(self):
        """Synchronise and read the global selection (same on every rank): one 40-byte read-back."""
        return self.ctx.best()

    def best_and_trajectory(self):
        """Global selection and the winner's 13 fp64 rows, regenerated on this rank (every rank holds the whole profile
        table): one launch, one copy, one synchronisation."""
        return self.ctx.best_and_trajectory()

    def best_from_records(self):
        """The same selection from the gathered records on the host (etp_combine_best): cross-check of the device path."""
        with self.ctx.torch.cuda.stream(self.ctx.stream):
            host = self._gather.cpu()
        recs = records_from_bytes(host.numpy(), self.world)
        rc, out = combine_best_records(recs)
        if rc == _abi.ERR_NO_CANDIDATE:
            from .scoring import NoCandidateError

            raise NoCandidateError("every candidate of every shard was skipped")
        return dict(index=out.index, level=out.level, cost=out.cost, n_scored=out.n_scored,
                    list_index=out.list_index, key_hi=out.key_hi, key_lo=out.key_lo)
