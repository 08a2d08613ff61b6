"""Topic sharding across GPUs (one process per GPU) and the result gather.

Topics are independent in the reference — totals and counts are created afresh
for each one (LagBasedPartitionAssignor.java:215-225) and the batch driver loops
over them sequentially (Main:177-184) — so they shard with no data-path
exchange. The only collective is the one the path needs at the end: an
all-gather so every rank holds the whole assignment map (NCCL over NVLink on
GPUs). What crosses NVLink is the NARROW form of each rank's results
(exchange_format(): partition ids in 1/2/4 bytes, consumers in 0/2/4 bytes —
0 when every topic has P <= C, because the consumer of sorted partition k is
then the k-th member of the topic's set and every rank writes it itself);
lagassign_exchange_pack / _unpack (include/lagassign.h) narrow and widen on the
device. The reference has no counterpart (single JVM thread).

Shards are CONTIGUOUS topic ranges balanced by partition count, so the global
result is simply the concatenation of the ranks' results. A single huge topic
is never split: its round fill is one sequential chain.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_bounds(topic_off: np.ndarray, world_size: int) -> List[Tuple[int, int]]:
    """Split topics [0, T) into world_size contiguous ranges with ~equal partition
    counts. Returns [(t_lo, t_hi)] per rank; ranges may be empty."""
    topic_off = np.asarray(topic_off, dtype=np.int64)
    T = len(topic_off) - 1
    N = int(topic_off[-1]) if T >= 0 else 0
    bounds = [0]
    for r in range(1, world_size):
        target = (N * r) // world_size
        # first topic boundary whose prefix count reaches the target, never moving backwards
        t = int(np.searchsorted(topic_off, target, side="left"))
        t = min(max(t, bounds[-1]), T)
        bounds.append(t)
    bounds.append(T)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def shard_batch(batch: dict, t_lo: int, t_hi: int) -> dict:
    """The sub-batch of topics [t_lo, t_hi) in the flat layout (consumer sets are replicated)."""
    off = np.asarray(batch["topic_off"], dtype=np.int64)
    lo, hi = int(off[t_lo]), int(off[t_hi])
    part = batch.get("partition")
    return dict(topic_off=(off[t_lo:t_hi + 1] - lo), lag=batch["lag"][lo:hi],
                partition=None if part is None else part[lo:hi],
                topic_cset=batch["topic_cset"][t_lo:t_hi], cset_off=batch["cset_off"],
                cset_rank=batch["cset_rank"])


def partition_counts(topic_off: np.ndarray, bounds: Sequence[Tuple[int, int]]) -> List[int]:
    off = np.asarray(topic_off, dtype=np.int64)
    return [int(off[hi] - off[lo]) for lo, hi in bounds]


def exchange_format(batch: dict) -> Tuple[int, int]:
    """(part_width, cons_width) in bytes of the narrow result form for this batch, see
    include/lagassign.h "Result exchange". Decided from the WHOLE batch, identically on every rank."""
    off = np.asarray(batch["topic_off"], dtype=np.int64)
    P = np.diff(off)
    part = batch.get("partition")
    if part is None:
        max_id = int(P.max()) - 1 if len(P) else 0
        min_id = 0
    else:
        part = np.asarray(part)
        max_id = int(part.max()) if part.size else 0
        min_id = int(part.min()) if part.size else 0
    pw = 4 if min_id < 0 or max_id >= 65536 else (2 if max_id >= 256 else 1)
    cset_off = np.asarray(batch["cset_off"], dtype=np.int64)
    csize = np.diff(cset_off)
    C = csize[np.asarray(batch["topic_cset"], dtype=np.int64)] if len(P) else np.zeros(0, np.int64)
    if bool(np.all((P <= C) | (C == 0))):
        cw = 0                                  # consumers are derivable everywhere
    else:
        ranks = np.asarray(batch["cset_rank"])
        n_members = int(ranks.max()) + 1 if ranks.size else 0
        cw = 2 if n_members < 65535 else 4      # 0xFFFF stands for "no subscriber" (-1)
    return pw, cw


def all_gather_results(local_part, local_cons, counts: Sequence[int], group=None):
    """All-gather the ranks' (out_partition, out_consumer_rank) torch tensors
    (int32, 1-D, on the backend's device) into the full arrays of sum(counts)
    entries, in rank order. Equal counts use one all_gather_into_tensor per array;
    ragged shards are padded to the largest one."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert len(counts) == world and local_part.numel() == counts[rank] == local_cons.numel()
    total = int(sum(counts))
    full_p = torch.empty(total, dtype=torch.int32, device=local_part.device)
    full_c = torch.empty(total, dtype=torch.int32, device=local_part.device)
    if len(set(counts)) == 1:
        dist.all_gather_into_tensor(full_p, local_part.contiguous(), group=group)
        dist.all_gather_into_tensor(full_c, local_cons.contiguous(), group=group)
        return full_p, full_c
    m = max(counts)
    pad = torch.empty(2 * m, dtype=torch.int32, device=local_part.device)
    pad[:counts[rank]] = local_part
    pad[m:m + counts[rank]] = local_cons
    gathered = torch.empty(world * 2 * m, dtype=torch.int32, device=local_part.device)
    dist.all_gather_into_tensor(gathered, pad, group=group)
    o = 0
    for r, c in enumerate(counts):
        full_p[o:o + c] = gathered[r * 2 * m: r * 2 * m + c]
        full_c[o:o + c] = gathered[r * 2 * m + m: r * 2 * m + m + c]
        o += c
    return full_p, full_c


class ShardedAssignor:
    """One rank's view of a topic-sharded assign(): upload this rank's contiguous
    shard, run the kernels, narrow the results into this rank's slot of the gather
    buffer, all-gather the slots, widen them into the full int32 arrays.
    torch (CUDA + NCCL) is plumbing only."""

    def __init__(self, flat_assignor, rank: int, world_size: int, group=None):
        self.f = flat_assignor
        self.rank, self.world, self.group = rank, world_size, group
        self.counts = None
        self.full_part = self.full_cons = None

    def upload(self, batch: dict):
        import torch
        from . import load_library
        self.L = load_library()
        self.bounds = shard_bounds(batch["topic_off"], self.world)
        self.counts = partition_counts(batch["topic_off"], self.bounds)
        lo, hi = self.bounds[self.rank]
        self.local = shard_batch(batch, lo, hi)
        b = self.local
        self.f.upload(b["topic_off"], b["lag"], b["partition"], b["topic_cset"], b["cset_off"], b["cset_rank"])
        total = int(sum(self.counts))
        dev = torch.device("cuda", torch.cuda.current_device())
        self.start = int(sum(self.counts[:self.rank]))
        n = self.counts[self.rank]
        self.full_part = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        self.full_cons = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        if self.world == 1:
            self.f.bind_device_outputs(self.full_part.data_ptr(), self.full_cons.data_ptr())
            return
        self.n_topics = len(batch["topic_off"]) - 1
        self.pw, self.cw = exchange_format(batch)
        self.slot_bytes = int(self.L.lagassign_exchange_slot_bytes(max(self.counts), self.pw, self.cw))
        self.slots = torch.empty(self.world * max(self.slot_bytes, 16), dtype=torch.uint8, device=dev)
        self.loc_part = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        self.loc_cons = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        self.f.bind_device_outputs(self.loc_part.data_ptr(), self.loc_cons.data_ptr())
        # the whole batch's shape, for the widening kernel (small: 12 B per topic + the consumer sets)
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
        self.d_topic_off = t(batch["topic_off"], np.int64)
        self.d_topic_cset = t(batch["topic_cset"], np.int32)
        self.d_cset_off = t(batch["cset_off"], np.int64)
        self.d_cset_rank = t(batch["cset_rank"], np.int32) if len(batch["cset_rank"]) else torch.zeros(1, dtype=torch.int32, device=dev)
        self.d_first = t([lo_ for lo_, _ in self.bounds] + [self.n_topics], np.int64)

    def execute(self, gather: bool = True):
        self.f.execute()
        if gather:
            self.gather()

    def gather(self, events=None):
        """Narrow, all-gather, widen (no-op on one rank). `events`: four torch.cuda.Event recorded before the narrowing,
        before and after the collective and after the widening (bench.py's breakdown of the exchange)."""
        import ctypes
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return
        cur = torch.cuda.current_stream()
        st = ctypes.c_void_p(cur.cuda_stream)
        if events: events[0].record(cur)
        vp = ctypes.c_void_p
        my = self.slots[self.rank * self.slot_bytes:(self.rank + 1) * self.slot_bytes]
        rc = self.L.lagassign_exchange_pack(st, self.counts[self.rank], vp(self.loc_part.data_ptr()), vp(self.loc_cons.data_ptr()),
                                            self.pw, self.cw, vp(my.data_ptr()))
        assert rc == 0, rc
        if events: events[1].record(cur)
        dist.all_gather_into_tensor(self.slots[:self.world * self.slot_bytes], my, group=self.group)
        if events: events[2].record(cur)
        rc = self.L.lagassign_exchange_unpack(st, self.n_topics, vp(self.d_topic_off.data_ptr()), vp(self.d_topic_cset.data_ptr()),
                                              vp(self.d_cset_off.data_ptr()), vp(self.d_cset_rank.data_ptr()),
                                              vp(self.d_first.data_ptr()), self.world, vp(self.slots.data_ptr()),
                                              self.slot_bytes, self.pw, self.cw,
                                              vp(self.full_part.data_ptr()), vp(self.full_cons.data_ptr()))
        assert rc == 0, rc
        if events: events[3].record(cur)

    def results(self):
        """(out_partition, out_consumer_rank) of the whole batch as contiguous 1-D tensors."""
        total = int(sum(self.counts))
        return self.full_part[:total], self.full_cons[:total]
