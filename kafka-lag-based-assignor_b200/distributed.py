"""Topic sharding across GPUs (one process per GPU) and the result gather.

Topics are independent in the reference — totals and counts are created afresh
for each one (LagBasedPartitionAssignor.java:215-225) and the batch driver loops
over them sequentially (Main:177-184) — so they shard with no data-path
exchange. The only collective is the one the path needs at the end: an
all-gather of (out_partition, out_consumer_rank) so every rank holds the whole
assignment map (NCCL over NVLink on GPUs; the same code runs on gloo/CPU
tensors in the tests). The reference has no counterpart (single JVM thread).

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
    shard, run the kernels writing straight into this rank's slice of the gather
    buffers, all-gather. torch (CUDA + NCCL) is plumbing only."""

    def __init__(self, flat_assignor, rank: int, world_size: int, group=None):
        self.f = flat_assignor
        self.rank, self.world, self.group = rank, world_size, group
        self.counts = None
        self.full_part = self.full_cons = None

    def upload(self, batch: dict):
        import torch
        self.bounds = shard_bounds(batch["topic_off"], self.world)
        self.counts = partition_counts(batch["topic_off"], self.bounds)
        lo, hi = self.bounds[self.rank]
        self.local = shard_batch(batch, lo, hi)
        b = self.local
        self.f.upload(b["topic_off"], b["lag"], b["partition"], b["topic_cset"], b["cset_off"], b["cset_rank"])
        total = int(sum(self.counts))
        dev = torch.device("cuda", torch.cuda.current_device())
        self.equal = len(set(self.counts)) == 1
        self.start = int(sum(self.counts[:self.rank]))
        n = self.counts[self.rank]
        if self.equal:
            # One buffer [rank][partition | consumer][n]: the kernels write into this rank's block and
            # ONE in-place all-gather moves both result arrays (a second collective launch costs more
            # than the extra bytes).
            self.buf = torch.empty((self.world, 2, max(n, 1)), dtype=torch.int32, device=dev)
            self.f.bind_device_outputs(self.buf[self.rank, 0].data_ptr(), self.buf[self.rank, 1].data_ptr())
            self.full_part = self.full_cons = None
        else:
            self.full_part = torch.empty(total, dtype=torch.int32, device=dev)
            self.full_cons = torch.empty(total, dtype=torch.int32, device=dev)
            self.loc_part = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
            self.loc_cons = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
            self.f.bind_device_outputs(self.loc_part.data_ptr(), self.loc_cons.data_ptr())

    def execute(self, gather: bool = True):
        self.f.execute()
        if gather:
            self.gather()

    def gather(self):
        """All-gather the ranks' results (no-op on one rank)."""
        import torch.distributed as dist
        if self.world == 1:
            return
        n = self.counts[self.rank]
        if self.equal:
            dist.all_gather_into_tensor(self.buf.view(-1), self.buf[self.rank].view(-1), group=self.group)
        else:
            self.full_part, self.full_cons = all_gather_results(self.loc_part[:n], self.loc_cons[:n],
                                                                self.counts, self.group)

    def results(self):
        """(out_partition, out_consumer_rank) of the whole batch as contiguous 1-D tensors."""
        if self.equal:
            n = self.counts[self.rank]
            return self.buf[:, 0, :n].reshape(-1), self.buf[:, 1, :n].reshape(-1)
        return self.full_part, self.full_cons
