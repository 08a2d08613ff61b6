"""Map-form <-> flat-array conversion for the ORACLE side of the tests.

TEST INFRASTRUCTURE ONLY (the product has its own flatten/unflatten in the C++
host mirror; this one is written separately so the two can disagree).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import oracle_py


def flatten(partition_lag_per_topic: Dict[str, List[Tuple[str, int, int]]],
            subscriptions: Sequence[Tuple[str, Sequence[str]]] | Dict[str, Sequence[str]]):
    """Returns (flat dict of numpy arrays, topics in HashMap order, members by rank)."""
    if isinstance(subscriptions, dict):
        subscriptions = list(subscriptions.items())
    topic_order, cpt = oracle_py.consumers_per_topic(subscriptions)
    rank = oracle_py.rank_members(m for m, _ in subscriptions)
    members_by_rank = sorted(rank, key=rank.get)
    topic_off = [0]
    lag: List[int] = []
    part: List[int] = []
    csets: Dict[Tuple[int, ...], int] = {}
    topic_cset: List[int] = []
    for t in topic_order:
        entries = partition_lag_per_topic.get(t, [])
        for (_tn, p, l) in entries:
            part.append(p)
            lag.append(l)
        topic_off.append(len(lag))
        # a member that lists the topic twice appears twice in cpt[t] but is ONE key of the
        # per-topic maps (Main:216-219): a consumer set is a set
        key = tuple(sorted({rank[m] for m in cpt[t]}))
        topic_cset.append(csets.setdefault(key, len(csets)))
    cset_off = [0]
    cset_rank: List[int] = []
    for key in csets:  # insertion order == set id order
        cset_rank.extend(key)
        cset_off.append(len(cset_rank))
    flat = dict(
        topic_off=np.array(topic_off, dtype=np.int64),
        lag=np.array(lag, dtype=np.int64),
        partition=np.array(part, dtype=np.int32),
        topic_cset=np.array(topic_cset, dtype=np.int32),
        cset_off=np.array(cset_off, dtype=np.int64),
        cset_rank=np.array(cset_rank, dtype=np.int32),
    )
    return flat, topic_order, members_by_rank


def unflatten(flat, topic_order, members_by_rank, out_partition, out_consumer_rank,
              all_members: Sequence[str]):
    """Rebuild Map<member, List<(topic, partition)>> (Main:264 append order)."""
    res = {m: [] for m in all_members}
    off = flat["topic_off"]
    for ti, t in enumerate(topic_order):
        for k in range(int(off[ti]), int(off[ti + 1])):
            r = int(out_consumer_rank[k])
            if r >= 0:
                res[members_by_rank[r]].append((t, int(out_partition[k])))
    return res
