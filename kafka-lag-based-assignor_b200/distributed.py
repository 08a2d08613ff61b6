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


def chunk_ranges(bounds: Sequence[Tuple[int, int]], n_chunks: int) -> List[List[Tuple[int, int]]]:
    """Every rank's contiguous topic range cut into n_chunks contiguous pieces: [[(lo, hi) per rank] per chunk]."""
    return [[(lo + (hi - lo) * c // n_chunks, lo + (hi - lo) * (c + 1) // n_chunks) for lo, hi in bounds]
            for c in range(n_chunks)]


def pick_chunks(local_partitions: int, has_radix_topics: bool) -> int:
    """Chunks of the exchange pipeline (a chunk's all-gather runs while the next chunk's kernels do). Measured at
    BASELINE config 4 on 2 x B200 with the steps driven from Python: 4 chunks 0.227 ms per step, 1 chunk 0.191 ms — the
    collective does hide behind the kernels (the compute stream waits 16 us for it instead of 62), but four small kernel,
    narrow and widen launches each cost more than they hide. So: one chunk unless the caller asks (n_chunks)."""
    return 1


class ShardedAssignor:
    """One rank's view of a topic-sharded assign(): upload this rank's contiguous shard; then, chunk by chunk of the
    shard, run the kernels (they write this rank's int32 results straight into its place in the full arrays), narrow
    the chunk into this rank's slot and start its all-gather, which proceeds on NCCL's stream while the next chunk's
    kernels run; finally widen the other ranks' slots into the full arrays. torch (CUDA + NCCL) is plumbing only."""

    def __init__(self, flat_assignor, rank: int, world_size: int, group=None, n_chunks: int = 0):
        self.f = flat_assignor
        self.rank, self.world, self.group = rank, world_size, group
        self.counts = None
        self.full_part = self.full_cons = None
        self.force_chunks = n_chunks

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
        self.full_part = torch.empty(max(total, 1) + 4, dtype=torch.int32, device=dev)
        self.full_cons = torch.empty(max(total, 1) + 4, dtype=torch.int32, device=dev)
        # this rank's kernels write their results where they belong in the full arrays
        self.f.bind_device_outputs(self.full_part.data_ptr() + 4 * self.start, self.full_cons.data_ptr() + 4 * self.start)
        if self.world == 1:
            return
        off = np.asarray(batch["topic_off"], dtype=np.int64)
        self.n_topics = len(off) - 1
        self.pw, self.cw = exchange_format(batch)
        has_radix = bool(self.f.stats()["topics_device"] > 0)
        # the same chunking on every rank: decided from the largest shard
        self.n_chunks = self.force_chunks or pick_chunks(max(self.counts), False)
        if self.world > 1:
            import torch.distributed as dist
            flag = torch.tensor([1 if has_radix else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)      # any rank with radix-path topics: one chunk everywhere
            if int(flag.item()):
                self.n_chunks = 1
        self.use_ranges = self.n_chunks > 1          # one chunk: lagassign_execute() (all size bins, concurrent streams)
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
        self.chunks = []
        for ranges in chunk_ranges(self.bounds, self.n_chunks):
            n_rc = [int(off[h] - off[l]) for l, h in ranges]
            sb = int(self.L.lagassign_exchange_slot_bytes(max(n_rc), self.pw, self.cw))
            my_lo, my_hi = ranges[self.rank]
            self.chunks.append(dict(
                local=(my_lo - lo, my_hi - lo), e0=int(off[my_lo]), n=n_rc[self.rank], slot_bytes=sb,
                slots=torch.empty(self.world * max(sb, 16), dtype=torch.uint8, device=dev),
                d_lo=t([l for l, _ in ranges], np.int64), d_hi=t([h for _, h in ranges], np.int64),
                max_topics=max(h - l for l, h in ranges)))
        # the whole batch's shape, for the widening kernel (small: 12 B per topic + the consumer sets)
        self.d_topic_off = t(off, np.int64)
        self.d_topic_cset = t(batch["topic_cset"], np.int32)
        self.d_cset_off = t(batch["cset_off"], np.int64)
        self.d_cset_rank = t(batch["cset_rank"], np.int32) if len(batch["cset_rank"]) else torch.zeros(1, dtype=torch.int32, device=dev)

    def execute(self, gather: bool = True, events=None):
        """One step. `events` (bench.py): a list that receives (name, start event, end event) triples of the exchange legs."""
        if self.world == 1 or not gather:
            self.f.execute()
            return
        self._step(events)

    def _step(self, events=None):
        import ctypes
        import torch
        import torch.distributed as dist
        cur = torch.cuda.current_stream()
        st = ctypes.c_void_p(cur.cuda_stream)
        vp = ctypes.c_void_p

        def mark():
            e = torch.cuda.Event(enable_timing=True)
            e.record(cur)
            return e
        works = []
        for ch in self.chunks:
            if self.use_ranges:
                self.f.execute_range(*ch["local"])
            else:
                self.f.execute()
            e0 = mark() if events is not None else None
            my = ch["slots"][self.rank * ch["slot_bytes"]:(self.rank + 1) * ch["slot_bytes"]]
            rc = self.L.lagassign_exchange_pack(st, ch["n"], vp(self.full_part.data_ptr() + 4 * ch["e0"]),
                                                vp(self.full_cons.data_ptr() + 4 * ch["e0"]), self.pw, self.cw, vp(my.data_ptr()))
            assert rc == 0, rc
            e1 = mark() if events is not None else None
            works.append(dist.all_gather_into_tensor(ch["slots"][:self.world * ch["slot_bytes"]], my, group=self.group, async_op=True)
                         if ch["slot_bytes"] > 0 else None)      # (a chunk without partitions on any rank)
            if events is not None:
                events.append(("narrow", e0, e1))
        for ch, w in zip(self.chunks, works):
            e0 = mark() if events is not None else None
            if w is None:
                continue
            w.wait()                                   # the current stream waits for this chunk's collective
            e1 = mark() if events is not None else None
            rc = self.L.lagassign_exchange_unpack_ranges(
                st, vp(self.d_topic_off.data_ptr()), vp(self.d_topic_cset.data_ptr()), vp(self.d_cset_off.data_ptr()),
                vp(self.d_cset_rank.data_ptr()), vp(ch["d_lo"].data_ptr()), vp(ch["d_hi"].data_ptr()), self.world, self.rank,
                ch["max_topics"], vp(ch["slots"].data_ptr()), ch["slot_bytes"], self.pw, self.cw,
                vp(self.full_part.data_ptr()), vp(self.full_cons.data_ptr()))
            assert rc == 0, rc
            if events is not None:
                events.append(("all_gather_wait", e0, e1))
                events.append(("widen", e1, mark()))

    def results(self):
        """(out_partition, out_consumer_rank) of the whole batch as contiguous 1-D tensors."""
        total = int(sum(self.counts))
        return self.full_part[:total], self.full_cons[:total]
