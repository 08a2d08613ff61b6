"""Python view of the host mirror of the reference plugin class.

``LagBasedPartitionAssignor`` here has the reference's surface for the assign()
path (LagBasedPartitionAssignor.java): ``configure`` (Main:97-130), ``name``
(Main:132-135), ``assign(partitionLagPerTopic, subscriptions)`` (Main:166-188) and
``compute_partition_lag`` (Main:376-404). Every call goes through the C++ host
mirror (``liblagassign_host.so``, include/lagassign_host.h), which flattens the
maps, calls the CUDA library and unflattens the result. Nothing is computed in
Python and there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
from collections import namedtuple
from typing import Dict, List, Mapping, Optional, Sequence

import numpy as np

TopicPartition = namedtuple("TopicPartition", ["topic", "partition"])
TopicPartitionLag = namedtuple("TopicPartitionLag", ["topic", "partition", "lag"])   # Main:431-455

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_host = None


def _load_host():
    global _host
    if _host is not None:
        return _host
    from . import LagAssignError, load_library
    load_library()   # the CUDA library first: the host library links against it
    path = os.path.join(_PKG_DIR, "liblagassign_host.so")
    if not os.path.exists(path):
        raise LagAssignError(-5, f"{path} is missing: build it with __graft_entry__.build()")
    H = ctypes.CDLL(path)
    vp, i64, i32, cp = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_char_p
    H.lagassign_group_create.argtypes = [ctypes.POINTER(vp), i32]
    H.lagassign_group_destroy.argtypes = [vp]
    H.lagassign_group_destroy.restype = None
    H.lagassign_group_last_error.argtypes = [vp]
    H.lagassign_group_last_error.restype = cp
    H.lagassign_group_configure.argtypes = [vp, ctypes.POINTER(cp), ctypes.POINTER(cp), i32]
    H.lagassign_group_metadata_client_id.argtypes = [vp]
    H.lagassign_group_metadata_client_id.restype = cp
    H.lagassign_group_name.argtypes = [vp]
    H.lagassign_group_name.restype = cp
    H.lagassign_group_clear.argtypes = [vp]
    H.lagassign_group_add_subscription.argtypes = [vp, cp, ctypes.POINTER(cp), i32]
    H.lagassign_group_add_topic_lags.argtypes = [vp, cp, vp, vp, i64]
    H.lagassign_group_add_topic_offsets.argtypes = [vp, cp, vp, vp, vp, vp, vp, i64]
    H.lagassign_group_set_auto_offset_reset.argtypes = [vp, cp]
    H.lagassign_group_assign.argtypes = [vp]
    H.lagassign_group_flatten.argtypes = [vp, vp]
    H.lagassign_group_unflatten.argtypes = [vp, vp, vp]
    H.lagassign_group_member_count.argtypes = [vp]
    H.lagassign_group_member_count.restype = i64
    H.lagassign_group_member_at_rank.argtypes = [vp, i64]
    H.lagassign_group_member_at_rank.restype = cp
    H.lagassign_group_result_size.argtypes = [vp, cp]
    H.lagassign_group_result_size.restype = i64
    H.lagassign_group_result_get.argtypes = [vp, cp, i64, ctypes.POINTER(cp), ctypes.POINTER(ctypes.c_int32)]
    H.lagassign_group_topic_count.argtypes = [vp]
    H.lagassign_group_topic_count.restype = i64
    H.lagassign_group_topic_at.argtypes = [vp, i64]
    H.lagassign_group_topic_at.restype = cp
    H.lagassign_compute_partition_lag.argtypes = [i32, i64, i64, i64, cp]
    H.lagassign_compute_partition_lag.restype = i64
    _host = H
    return H


def _enc(s: str) -> bytes:
    return s.encode("utf-8", "surrogatepass")


def compute_partition_lag(committed: Optional[int], begin_offset: int, end_offset: int,
                          auto_offset_reset_mode: str) -> int:
    """computePartitionLag (Main:376-404); ``committed`` None = no committed offset."""
    H = _load_host()
    return int(H.lagassign_compute_partition_lag(0 if committed is None else 1,
                                                 0 if committed is None else committed,
                                                 begin_offset, end_offset, _enc(auto_offset_reset_mode)))


class LagBasedPartitionAssignor:
    """Host mirror of com.github.grantneale.kafka.LagBasedPartitionAssignor."""

    def __init__(self, device: int = 0):
        from . import LagAssignError
        self._H = _load_host()
        self._g = ctypes.c_void_p()
        rc = self._H.lagassign_group_create(ctypes.byref(self._g), device)
        if rc != 0:
            raise LagAssignError(rc, self._H.lagassign_group_last_error(None).decode())
        self.topic_order: List[str] = []

    def close(self):
        if getattr(self, "_g", None) is not None and self._g.value:
            self._H.lagassign_group_destroy(self._g)
            self._g = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            from . import LagAssignError
            raise LagAssignError(rc, self._H.lagassign_group_last_error(self._g).decode())

    # Configurable.configure, Main:97-130
    def configure(self, configs: Mapping[str, object]) -> None:
        keys = [_enc(str(k)) for k in configs]
        vals = [_enc(str(v)) for v in configs.values()]
        ka = (ctypes.c_char_p * len(keys))(*keys)
        va = (ctypes.c_char_p * len(vals))(*vals)
        rc = self._H.lagassign_group_configure(self._g, ka, va, len(keys))
        if rc == -1:
            # the reference throws IllegalArgumentException (Main:108-113)
            raise ValueError(self._H.lagassign_group_last_error(self._g).decode())
        self._check(rc)

    @property
    def metadata_client_id(self) -> str:
        return self._H.lagassign_group_metadata_client_id(self._g).decode()

    # ConsumerPartitionAssignor.name, Main:132-135
    def name(self) -> str:
        return self._H.lagassign_group_name(self._g).decode()

    def _load_maps(self, partition_lag_per_topic, subscriptions) -> List[str]:
        H = self._H
        self._check(H.lagassign_group_clear(self._g))
        members = list(subscriptions.keys())            # iteration order of the caller's map
        for m in members:
            topics = [_enc(t) for t in subscriptions[m]]
            ta = (ctypes.c_char_p * max(1, len(topics)))(*topics)
            self._check(H.lagassign_group_add_subscription(self._g, _enc(m), ta, len(topics)))
        for topic, entries in partition_lag_per_topic.items():
            parts = np.array([e[1] for e in entries], dtype=np.int32)
            lags = np.array([e[2] for e in entries], dtype=np.int64)
            self._check(H.lagassign_group_add_topic_lags(
                self._g, _enc(topic), parts.ctypes.data if len(parts) else None,
                lags.ctypes.data if len(lags) else None, len(lags)))
        return members

    # static assign(Map, Map), Main:166-188
    def assign(self, partition_lag_per_topic: Mapping[str, Sequence[TopicPartitionLag]],
               subscriptions: Mapping[str, Sequence[str]]) -> Dict[str, List[TopicPartition]]:
        H = self._H
        members = self._load_maps(partition_lag_per_topic, subscriptions)
        self._check(H.lagassign_group_assign(self._g))
        return self._collect(members)

    def assign_from_offsets(self, offsets_per_topic, subscriptions, auto_offset_reset: str = "latest"):
        """The same assignment from raw offsets: {topic: [(partition, begin, end, committed or None), ...]}.
        computePartitionLag (Main:376-404) runs on the device (lagassign_group_add_topic_offsets)."""
        H = self._H
        members = self._load_maps({}, subscriptions)
        self._check(H.lagassign_group_set_auto_offset_reset(self._g, _enc(auto_offset_reset)))
        for topic, entries in offsets_per_topic.items():
            n = len(entries)
            parts = np.array([e[0] for e in entries], dtype=np.int32)
            begin = np.array([e[1] for e in entries], dtype=np.int64)
            end = np.array([e[2] for e in entries], dtype=np.int64)
            com = np.array([0 if e[3] is None else e[3] for e in entries], dtype=np.int64)
            has = np.array([0 if e[3] is None else 1 for e in entries], dtype=np.uint8)
            ptr = lambda a: a.ctypes.data if n else None
            self._check(H.lagassign_group_add_topic_offsets(self._g, _enc(topic), ptr(parts), ptr(begin), ptr(end),
                                                            ptr(com), ptr(has), n))
        self._check(H.lagassign_group_assign(self._g))
        return self._collect(members)

    # The host-side halves of assign() (lagassign_group_flatten / _unflatten): what happens before
    # and after the device call. No GPU needed; the CPU tests drive them with the oracle in between.
    def flatten(self, partition_lag_per_topic, subscriptions):
        """Returns (flat dict of numpy arrays as in include/lagassign.h, topics in assignment
        order, members by String.compareTo rank). Arrays are copies."""
        from . import _Batch
        H = self._H
        self._flat_members = self._load_maps(partition_lag_per_topic, subscriptions)
        b = _Batch()
        self._check(H.lagassign_group_flatten(self._g, ctypes.byref(b)))
        T, S = int(b.n_topics), int(b.n_csets)

        def arr(ptr, n, dtype):
            if not ptr or n == 0:
                return np.zeros(0, dtype)
            ct = ctypes.c_int64 if dtype == np.int64 else ctypes.c_int32
            return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ct)), shape=(n,)).astype(dtype, copy=True)
        topic_off = arr(b.topic_off, T + 1, np.int64)
        N = int(topic_off[-1]) if T >= 0 and len(topic_off) else 0
        cset_off = arr(b.cset_off, S + 1, np.int64)
        R = int(cset_off[-1]) if len(cset_off) else 0
        flat = dict(topic_off=topic_off, lag=arr(b.lag, N, np.int64),
                    partition=arr(b.partition, N, np.int32) if b.partition else None,
                    topic_cset=arr(b.topic_cset, T, np.int32), cset_off=cset_off,
                    cset_rank=arr(b.cset_rank, R, np.int32))
        topics = [H.lagassign_group_topic_at(self._g, i).decode("utf-8", "surrogatepass")
                  for i in range(H.lagassign_group_topic_count(self._g))]
        by_rank = [H.lagassign_group_member_at_rank(self._g, r).decode("utf-8", "surrogatepass")
                   for r in range(H.lagassign_group_member_count(self._g))]
        self.topic_order = topics
        return flat, topics, by_rank

    def unflatten(self, out_partition, out_consumer_rank) -> Dict[str, List[TopicPartition]]:
        """Per-member lists from the device step's outputs for the batch of the last flatten()."""
        op = np.ascontiguousarray(out_partition, dtype=np.int32)
        oc = np.ascontiguousarray(out_consumer_rank, dtype=np.int32)
        self._check(self._H.lagassign_group_unflatten(self._g, op.ctypes.data if len(op) else None,
                                                       oc.ctypes.data if len(oc) else None))
        return self._collect(self._flat_members)

    def _collect(self, members) -> Dict[str, List[TopicPartition]]:
        H = self._H
        out: Dict[str, List[TopicPartition]] = {}
        t = ctypes.c_char_p()
        p = ctypes.c_int32()
        for m in members:
            n = H.lagassign_group_result_size(self._g, _enc(m))
            lst = []
            for i in range(max(0, n)):
                self._check(H.lagassign_group_result_get(self._g, _enc(m), i, ctypes.byref(t), ctypes.byref(p)))
                lst.append(TopicPartition(t.value.decode("utf-8", "surrogatepass"), int(p.value)))
            out[m] = lst
        self.topic_order = [H.lagassign_group_topic_at(self._g, i).decode("utf-8", "surrogatepass")
                            for i in range(H.lagassign_group_topic_count(self._g))]
        return out
