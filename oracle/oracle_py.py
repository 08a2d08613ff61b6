"""Pure-Python restatement of the reference's assign() — TEST INFRASTRUCTURE ONLY.

Written independently of oracle.c (maps of objects like the Java original, not
flat arrays) so the two can be cross-checked on generated inputs. Small cases
only: it is O(P*C) Python loops.

"Main:n" = /root/reference/src/main/java/com/github/grantneale/kafka/
LagBasedPartitionAssignor.java line n.
"""
from __future__ import annotations

import functools
from typing import Dict, Iterable, List, Sequence, Tuple

_M64 = (1 << 64) - 1


def wrap64(x: int) -> int:
    """Java long arithmetic: wrap to signed 64 bits."""
    x &= _M64
    return x - (1 << 64) if x >> 63 else x


def utf16_units(s: str) -> List[int]:
    b = s.encode("utf-16-be", "surrogatepass")
    return [(b[i] << 8) | b[i + 1] for i in range(0, len(b), 2)]


def java_compare_to(a: str, b: str) -> int:
    """java.lang.String.compareTo: first differing UTF-16 code unit, else length."""
    ua, ub = utf16_units(a), utf16_units(b)
    for x, y in zip(ua, ub):
        if x != y:
            return x - y
    return len(ua) - len(ub)


def java_string_hash(s: str) -> int:
    """java.lang.String.hashCode as a signed 32-bit int."""
    h = 0
    for u in utf16_units(s):
        h = (31 * h + u) & 0xFFFFFFFF
    return h


class JavaHashMapOrder:
    """Iteration order of a java.util.HashMap<String, ?> (OpenJDK 8+) that only
    ever receives put() of new keys: enough for consumersPerTopic (Main:412-424),
    whose iteration at Main:177 fixes the order of every member's result list.

    Models: default capacity 16, load factor .75, power-of-two growth with
    order-preserving bucket splits, bucket index (h ^ h>>>16) & (n-1), append at
    the chain tail. Treeified bins (a 9th key in one bucket while the table has
    >= 64 slots) reorder the chain and are not modelled: raises.
    """

    def __init__(self, initial_capacity: int = 16):
        self.cap = initial_capacity
        self.buckets: List[List[str]] = [[] for _ in range(self.cap)]
        self.size = 0
        self.keys = set()

    @staticmethod
    def spread(h: int) -> int:
        return (h ^ (h >> 16)) & 0xFFFFFFFF

    def put_if_absent(self, key: str) -> bool:
        if key in self.keys:
            return False
        h = self.spread(java_string_hash(key))
        chain = self.buckets[h & (self.cap - 1)]
        chain.append(key)
        self.keys.add(key)
        if len(chain) >= 9:
            if self.cap < 64:
                self._resize()          # treeifyBin resizes small tables instead
            else:
                raise NotImplementedError("treeified HashMap bin: iteration order not modelled")
        self.size += 1
        if self.size > self.cap * 3 // 4:
            self._resize()
        return True

    def _resize(self):
        old = self.buckets
        self.cap *= 2
        self.buckets = [[] for _ in range(self.cap)]
        for chain in old:
            for k in chain:
                h = self.spread(java_string_hash(k))
                self.buckets[h & (self.cap - 1)].append(k)

    def order(self) -> List[str]:
        return [k for chain in self.buckets for k in chain]


TopicPartitionLag = Tuple[str, int, int]   # (topic, partition, lag)   Main:431-455
TopicPartition = Tuple[str, int]


def consumers_per_topic(subscriptions: Sequence[Tuple[str, Sequence[str]]]):
    """Main:410-426. ``subscriptions`` is the member -> topics map given as its
    entries IN ITERATION ORDER. Returns (topic order as a HashMap would iterate,
    topic -> [members in arrival order])."""
    order = JavaHashMapOrder()
    res: Dict[str, List[str]] = {}
    for member, topics in subscriptions:
        for topic in topics:
            order.put_if_absent(topic)
            res.setdefault(topic, []).append(member)
    return order.order(), res


def assign_topic(assignment: Dict[str, List[TopicPartition]], topic: str,
                 consumers: Sequence[str], partition_lags: List[TopicPartitionLag],
                 totals_out: Dict[str, Dict[str, Tuple[int, int]]] | None = None) -> None:
    """Main:204-277, statement by statement."""
    if not consumers:                                              # Main:211-213
        return
    total = {m: 0 for m in consumers}                              # Main:216-219
    count = {m: 0 for m in consumers}                              # Main:222-225

    def cmp_partition(p1, p2):                                     # Main:228-235
        if p1[2] == p2[2]:
            return (p1[1] > p2[1]) - (p1[1] < p2[1])
        return (p2[2] > p1[2]) - (p2[2] < p1[2])

    partition_lags.sort(key=functools.cmp_to_key(cmp_partition))  # stable, in place

    def cmp_consumer(c1, c2):                                      # Main:243-261
        if count[c1] != count[c2]:
            return (count[c1] > count[c2]) - (count[c1] < count[c2])
        if total[c1] != total[c2]:
            return (total[c1] > total[c2]) - (total[c1] < total[c2])
        return java_compare_to(c1, c2)

    for (_t, partition, lag) in partition_lags:                    # Main:237
        best = None
        for m in total:                                            # Collections.min
            if best is None or cmp_consumer(m, best) < 0:
                best = m
        assignment[best].append((topic, partition))                # Main:264
        total[best] = wrap64(total[best] + lag)                    # Main:265
        count[best] += 1                                           # Main:266
    if totals_out is not None:
        totals_out[topic] = {m: (total[m], count[m]) for m in consumers}


def assign(partition_lag_per_topic: Dict[str, List[TopicPartitionLag]],
           subscriptions: Sequence[Tuple[str, Sequence[str]]] | Dict[str, Sequence[str]],
           totals_out=None) -> Dict[str, List[TopicPartition]]:
    """Main:166-188. A dict argument is taken in insertion order (like Guava's
    ImmutableMap in the reference's tests)."""
    if isinstance(subscriptions, dict):
        subscriptions = list(subscriptions.items())
    assignment: Dict[str, List[TopicPartition]] = {m: [] for m, _ in subscriptions}  # Main:171-174
    topic_order, cpt = consumers_per_topic(subscriptions)                            # Main:176
    for topic in topic_order:                                                        # Main:177
        lags = partition_lag_per_topic.get(topic, [])                                # Main:182
        assign_topic(assignment, topic, cpt[topic], lags, totals_out)
    return assignment


def compute_partition_lag(committed, begin_offset: int, end_offset: int, reset_mode: str) -> int:
    """Main:376-404. ``committed`` is None when there is no committed offset."""
    if committed is not None:
        nxt = committed
    elif reset_mode.lower() == "latest":
        nxt = end_offset
    else:
        nxt = begin_offset
    return max(wrap64(end_offset - nxt), 0)


def rank_members(ids: Iterable[str]) -> Dict[str, int]:
    s = sorted(set(ids), key=functools.cmp_to_key(java_compare_to))
    return {m: i for i, m in enumerate(s)}
