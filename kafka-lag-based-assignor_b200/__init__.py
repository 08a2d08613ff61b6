"""B200-native lag-based partition assignment — Python binding of the C ABI.

The product is ``liblagassign.so`` (hand-written sm_100a CUDA kernels behind the
plain-C interface of ``include/lagassign.h``), the drop-in for the pure function
``LagBasedPartitionAssignor.assign(Map, Map)`` of the reference
(LagBasedPartitionAssignor.java:166-188). This module is only a ctypes binding
for tests, benchmarks and Python callers; it contains no algorithm and there is
no CPU fallback: without the built library or without a CUDA device every call
raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

__all__ = ["LagAssignError", "FlatAssignor", "load_library", "library_path", "pinned_empty",
           "LagBasedPartitionAssignor", "TopicPartitionLag", "TopicPartition", "compute_partition_lag"]

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "liblagassign.so"
_lib: Optional[ctypes.CDLL] = None

ERR_NAMES = {0: "OK", -1: "ERR_ARG", -2: "ERR_CUDA", -3: "ERR_OOM", -4: "ERR_STATE",
             -5: "ERR_NO_DEVICE", -6: "ERR_UNSUPPORTED"}
OPT_TOTALS = 1
OPT_PROFILE = 2
OPT_GROUP = 3
OPT_PIPE_CHUNK = 4
OPT_PLAN_CACHE = 5
OPT_EXACT_RANK = 6


class LagAssignError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"lagassign {ERR_NAMES.get(code, code)}: {message}")
        self.code = code


class _Batch(ctypes.Structure):
    _fields_ = [("n_topics", ctypes.c_int64), ("topic_off", ctypes.c_void_p), ("lag", ctypes.c_void_p),
                ("partition", ctypes.c_void_p), ("topic_cset", ctypes.c_void_p), ("n_csets", ctypes.c_int64),
                ("cset_off", ctypes.c_void_p), ("cset_rank", ctypes.c_void_p)]


class _Result(ctypes.Structure):
    _fields_ = [("out_partition", ctypes.c_void_p), ("out_consumer_rank", ctypes.c_void_p),
                ("out_total", ctypes.c_void_p), ("out_count", ctypes.c_void_p)]


class _ResultNarrow(ctypes.Structure):
    _fields_ = [("out_partition", ctypes.c_void_p), ("out_consumer_rank", ctypes.c_void_p),
                ("part_width", ctypes.c_int32), ("cons_width", ctypes.c_int32)]


class _Subscriptions(ctypes.Structure):
    _fields_ = [("n_members", ctypes.c_int64), ("n_topics", ctypes.c_int64), ("sub_off", ctypes.c_void_p),
                ("sub_topic", ctypes.c_void_p)]


class _ConsumerSets(ctypes.Structure):
    _fields_ = [("n_csets", ctypes.c_int64), ("topic_cset", ctypes.c_void_p), ("cset_off", ctypes.c_void_p),
                ("cset_rank", ctypes.c_void_p)]


class _Offsets(ctypes.Structure):
    _fields_ = [("begin_offset", ctypes.c_void_p), ("end_offset", ctypes.c_void_p),
                ("committed", ctypes.c_void_p), ("has_committed", ctypes.c_void_p),
                ("reset_latest", ctypes.c_int32)]


class _Grouping(ctypes.Structure):
    _fields_ = [("member_off", ctypes.c_void_p), ("topic", ctypes.c_void_p), ("partition", ctypes.c_void_p)]


class Stats(ctypes.Structure):
    _fields_ = [("n_partitions", ctypes.c_int64), ("n_assigned", ctypes.c_int64),
                ("topics_warp", ctypes.c_int64), ("topics_cta", ctypes.c_int64),
                ("topics_device", ctypes.c_int64), ("topics_unassigned", ctypes.c_int64),
                ("device_sort_elems", ctypes.c_int64), ("radix_passes_run", ctypes.c_int32),
                ("kernel_launches", ctypes.c_int32), ("h2d_bytes", ctypes.c_int64),
                ("d2h_bytes", ctypes.c_int64)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def library_path() -> str:
    # LAGASSIGN_LIBRARY selects another build of the same library (tuning experiments)
    return os.environ.get("LAGASSIGN_LIBRARY") or os.path.join(_PKG_DIR, _LIB_NAME)


def load_library() -> ctypes.CDLL:
    """Load liblagassign.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise LagAssignError(-5, f"{path} is missing: build it with __graft_entry__.build() "
                                 "(nvcc, sm_100a). There is no CPU fallback.")
    L = ctypes.CDLL(path)
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    L.lagassign_abi_version.restype = i32
    L.lagassign_device_count.argtypes = [ctypes.POINTER(ctypes.c_int)]
    L.lagassign_create.argtypes = [ctypes.POINTER(vp), i32]
    L.lagassign_create_on_stream.argtypes = [ctypes.POINTER(vp), i32, vp]
    L.lagassign_set_stream.argtypes = [vp, vp]
    L.lagassign_create_multi.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_int), i32]
    L.lagassign_shard_count.argtypes = [vp, ctypes.POINTER(ctypes.c_int)]
    L.lagassign_destroy.argtypes = [vp]
    L.lagassign_destroy.restype = None
    L.lagassign_last_error.argtypes = [vp]
    L.lagassign_last_error.restype = ctypes.c_char_p
    L.lagassign_set_option.argtypes = [vp, i32, i64]
    L.lagassign_assign_batch.argtypes = [vp, ctypes.POINTER(_Batch), ctypes.POINTER(_Result)]
    L.lagassign_assign_batch_narrow.argtypes = [vp, ctypes.POINTER(_Batch), ctypes.POINTER(_ResultNarrow)]
    L.lagassign_upload.argtypes = [vp, ctypes.POINTER(_Batch)]
    L.lagassign_upload_offsets.argtypes = [vp, ctypes.POINTER(_Batch), ctypes.POINTER(_Offsets)]
    L.lagassign_execute.argtypes = [vp]
    L.lagassign_execute_range.argtypes = [vp, i64, i64]
    L.lagassign_exchange_unpack_ranges.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i64, vp, i64, i32, i32, vp, vp]
    L.lagassign_sync.argtypes = [vp]
    L.lagassign_download.argtypes = [vp, ctypes.POINTER(_Result)]
    L.lagassign_bind_device_outputs.argtypes = [vp, vp, vp]
    L.lagassign_download_lags.argtypes = [vp, vp]
    L.lagassign_get_stats.argtypes = [vp, ctypes.POINTER(Stats)]
    L.lagassign_group_sizes.argtypes = [vp, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
    L.lagassign_download_grouping.argtypes = [vp, ctypes.POINTER(_Grouping)]
    L.lagassign_profile_count.argtypes = [vp, ctypes.POINTER(ctypes.c_int)]
    L.lagassign_profile_entry.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_char_p),
                                          ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)]
    L.lagassign_exchange_slot_bytes.argtypes = [i64, i32, i32]
    L.lagassign_exchange_slot_bytes.restype = i64
    L.lagassign_exchange_pack.argtypes = [vp, i64, vp, vp, i32, i32, vp]
    L.lagassign_exchange_unpack.argtypes = [vp, i64, vp, vp, vp, vp, vp, i32, vp, i64, i32, i32, vp, vp]
    L.lagassign_invert_subscriptions.argtypes = [i32, ctypes.POINTER(_Subscriptions), ctypes.POINTER(_ConsumerSets)]
    L.lagassign_subscriptions_last_error.restype = ctypes.c_char_p
    L.lagassign_host_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_size_t]
    L.lagassign_host_free.argtypes = [vp]
    for name in ("lagassign_device_count", "lagassign_create", "lagassign_create_on_stream", "lagassign_set_stream", "lagassign_create_multi",
                 "lagassign_shard_count",
                 "lagassign_set_option", "lagassign_assign_batch", "lagassign_assign_batch_narrow", "lagassign_upload",
                 "lagassign_upload_offsets", "lagassign_execute", "lagassign_execute_range", "lagassign_sync", "lagassign_download",
                 "lagassign_exchange_unpack_ranges",
                 "lagassign_bind_device_outputs", "lagassign_download_lags", "lagassign_get_stats",
                 "lagassign_profile_count", "lagassign_profile_entry",
                 "lagassign_group_sizes", "lagassign_download_grouping",
                 "lagassign_exchange_pack", "lagassign_exchange_unpack", "lagassign_invert_subscriptions",
                 "lagassign_host_alloc", "lagassign_host_free"):
        getattr(L, name).restype = i32
    _lib = L
    return L


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def _arr(a, dtype, name):
    if a is None:
        return None
    out = np.ascontiguousarray(a, dtype=dtype)
    if out.ndim != 1:
        raise ValueError(f"{name} must be one-dimensional")
    return out


def _free_pinned(ptr):
    try:
        load_library().lagassign_host_free(ctypes.c_void_p(ptr))
    except Exception:
        pass


def pinned_empty(n: int, dtype) -> np.ndarray:
    """numpy array over pinned host memory from lagassign_host_alloc(); freed with the array."""
    import weakref
    dtype = np.dtype(dtype)
    L = load_library()
    p = ctypes.c_void_p()
    nbytes = max(1, n) * dtype.itemsize
    rc = L.lagassign_host_alloc(ctypes.byref(p), nbytes)
    if rc != 0:
        raise LagAssignError(rc, L.lagassign_last_error(None).decode())
    buf = (ctypes.c_char * nbytes).from_address(p.value)
    weakref.finalize(buf, _free_pinned, p.value)      # numpy arrays over `buf` hold a reference to it (their base)
    return np.frombuffer(buf, dtype=dtype, count=n)




def invert_subscriptions(n_topics: int, sub_off, sub_topic, device: int = 0) -> dict:
    """lagassign_invert_subscriptions(): consumersPerTopic (Main:410-426) + consumer-set de-duplication on the device.
    sub_off[M+1] / sub_topic: the topic indices every member (in rank order) subscribes to. Returns the consumer-set
    part of the flat batch: dict(topic_cset, cset_off, cset_rank)."""
    L = load_library()
    sub_off = _arr(sub_off, np.int64, "sub_off")
    sub_topic = _arr(sub_topic, np.int32, "sub_topic")
    M = len(sub_off) - 1
    R = int(sub_off[-1]) if M >= 0 and len(sub_off) else 0
    tc = np.empty(max(n_topics, 1), np.int32)
    co = np.empty(n_topics + 1, np.int64)
    cr = np.empty(max(R, 1), np.int32)
    s_in = _Subscriptions(max(M, 0), n_topics, _ptr(sub_off), _ptr(sub_topic))
    s_out = _ConsumerSets(0, _ptr(tc), _ptr(co), _ptr(cr))
    rc = L.lagassign_invert_subscriptions(device, ctypes.byref(s_in), ctypes.byref(s_out))
    if rc != 0:
        raise LagAssignError(rc, L.lagassign_subscriptions_last_error().decode())
    S = int(s_out.n_csets)
    return dict(topic_cset=tc[:n_topics].copy(), cset_off=co[:S + 1].copy(), cset_rank=cr[:int(co[S]) if n_topics else 0].copy())


class FlatAssignor:
    """One lagassign context (one GPU, one stream). Flat-array interface, see
    include/lagassign.h for the layout."""

    def __init__(self, device=0, stream: Optional[int] = None, want_totals: bool = False):
        """device: one device index, or a list of distinct indices for a multi-device context
        (lagassign_create_multi: topics are split over the devices, results land in the caller's arrays)."""
        self._L = load_library()
        self._ctx = ctypes.c_void_p()
        if isinstance(device, (list, tuple)):
            if stream is not None:
                raise ValueError("a multi-device context owns its streams")
            devs = (ctypes.c_int * len(device))(*[int(d) for d in device])
            rc = self._L.lagassign_create_multi(ctypes.byref(self._ctx), devs, len(device))
        elif stream is None:
            rc = self._L.lagassign_create(ctypes.byref(self._ctx), device)
        else:
            rc = self._L.lagassign_create_on_stream(ctypes.byref(self._ctx), device, ctypes.c_void_p(stream))
        if rc != 0:
            raise LagAssignError(rc, self._L.lagassign_last_error(None).decode())
        self._keep = None
        self._shape = None
        if want_totals:
            self.set_option(OPT_TOTALS, 1)
        self.want_totals = want_totals
        self.profiling = False

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._L.lagassign_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise LagAssignError(rc, self._L.lagassign_last_error(self._ctx).decode())

    def set_option(self, option: int, value: int):
        self._check(self._L.lagassign_set_option(self._ctx, option, value))
        if option == OPT_PROFILE:
            self.profiling = bool(value)

    def _batch(self, topic_off, lag, partition, topic_cset, cset_off, cset_rank):
        topic_off = _arr(topic_off, np.int64, "topic_off")
        lag = _arr(lag, np.int64, "lag")
        partition = _arr(partition, np.int32, "partition")
        topic_cset = _arr(topic_cset, np.int32, "topic_cset")
        cset_off = _arr(cset_off, np.int64, "cset_off")
        cset_rank = _arr(cset_rank, np.int32, "cset_rank")
        T = len(topic_off) - 1
        S = len(cset_off) - 1
        if T < 0 or S < 0:
            raise ValueError("topic_off and cset_off need at least one entry")
        N = int(topic_off[-1])
        if lag is not None and len(lag) != N:
            raise ValueError("len(lag) != topic_off[-1]")
        if partition is not None and len(partition) != N:
            raise ValueError("len(partition) != topic_off[-1]")
        if len(topic_cset) != T:
            raise ValueError("len(topic_cset) != number of topics")
        b = _Batch(T, _ptr(topic_off), _ptr(lag), _ptr(partition), _ptr(topic_cset), S,
                   _ptr(cset_off), _ptr(cset_rank))
        keep = (topic_off, lag, partition, topic_cset, cset_off, cset_rank)
        sizes = cset_off[1:] - cset_off[:-1]
        # invalid set ids are reported by the library's own validation, not here
        ok = T > 0 and S > 0 and topic_cset.min() >= 0 and topic_cset.max() < S
        total_tc = int(sizes[topic_cset].sum()) if ok else 0
        return b, keep, (T, N, total_tc)

    def upload(self, topic_off, lag, partition, topic_cset, cset_off, cset_rank):
        b, keep, shape = self._batch(topic_off, lag, partition, topic_cset, cset_off, cset_rank)
        self._check(self._L.lagassign_upload(self._ctx, ctypes.byref(b)))
        self._shape = shape

    def upload_offsets(self, topic_off, begin_offset, end_offset, committed, has_committed,
                       reset_latest, partition, topic_cset, cset_off, cset_rank):
        b, keep, shape = self._batch(topic_off, None, partition, topic_cset, cset_off, cset_rank)
        bo = _arr(begin_offset, np.int64, "begin_offset")
        eo = _arr(end_offset, np.int64, "end_offset")
        co = _arr(committed, np.int64, "committed")
        hc = _arr(has_committed, np.uint8, "has_committed")
        for a in (bo, eo, co, hc):
            if len(a) != shape[1]:
                raise ValueError("offset arrays must have topic_off[-1] entries")
        o = _Offsets(_ptr(bo), _ptr(eo), _ptr(co), _ptr(hc), 1 if reset_latest else 0)
        self._check(self._L.lagassign_upload_offsets(self._ctx, ctypes.byref(b), ctypes.byref(o)))
        self._shape = shape

    def execute(self):
        self._check(self._L.lagassign_execute(self._ctx))

    def set_stream(self, cuda_stream: int):
        """lagassign_set_stream(): later work is enqueued on this stream (contexts created on a caller's stream only)."""
        self._check(self._L.lagassign_set_stream(self._ctx, ctypes.c_void_p(cuda_stream)))

    def execute_range(self, t_lo: int, t_hi: int):
        """lagassign_execute_range(): the kernels of the topics [t_lo, t_hi) of the uploaded batch only."""
        self._check(self._L.lagassign_execute_range(self._ctx, t_lo, t_hi))

    def sync(self):
        self._check(self._L.lagassign_sync(self._ctx))

    def bind_device_outputs(self, d_out_partition: Optional[int], d_out_consumer_rank: Optional[int]):
        self._check(self._L.lagassign_bind_device_outputs(
            self._ctx, ctypes.c_void_p(d_out_partition), ctypes.c_void_p(d_out_consumer_rank)))

    def download(self, out_partition=None, out_consumer_rank=None):
        T, N, total_tc = self._shape
        op = np.empty(N, np.int32) if out_partition is None else out_partition
        oc = np.empty(N, np.int32) if out_consumer_rank is None else out_consumer_rank
        ot = on = None
        if self.want_totals:
            ot = np.empty(total_tc, np.int64)
            on = np.empty(total_tc, np.int32)
        r = _Result(_ptr(op), _ptr(oc), _ptr(ot), _ptr(on))
        self._check(self._L.lagassign_download(self._ctx, ctypes.byref(r)))
        return dict(out_partition=op, out_consumer_rank=oc, out_total=ot, out_count=on)

    def download_grouping(self):
        """Per-member CSR view of the result (OPT_GROUP enabled before upload):
        dict(member_off[M+1], topic[n_assigned], partition[n_assigned])."""
        m = ctypes.c_int64(0)
        a = ctypes.c_int64(0)
        self._check(self._L.lagassign_group_sizes(self._ctx, ctypes.byref(m), ctypes.byref(a)))
        off = np.empty(m.value + 1, np.int64)
        topic = np.empty(a.value, np.int32)
        part = np.empty(a.value, np.int32)
        g = _Grouping(_ptr(off), _ptr(topic), _ptr(part))
        self._check(self._L.lagassign_download_grouping(self._ctx, ctypes.byref(g)))
        return dict(member_off=off, topic=topic, partition=part)

    def download_lags(self):
        T, N, _ = self._shape
        out = np.empty(N, np.int64)
        self._check(self._L.lagassign_download_lags(self._ctx, _ptr(out)))
        return out

    def assign_batch(self, topic_off, lag, partition, topic_cset, cset_off, cset_rank,
                     out_partition=None, out_consumer_rank=None):
        """The whole hot path through lagassign_assign_batch() (host buffers in and out)."""
        b, keep, shape = self._batch(topic_off, lag, partition, topic_cset, cset_off, cset_rank)
        T, N, total_tc = shape
        op = np.empty(N, np.int32) if out_partition is None else out_partition
        oc = np.empty(N, np.int32) if out_consumer_rank is None else out_consumer_rank
        ot = on = None
        if self.want_totals:
            ot = np.empty(total_tc, np.int64)
            on = np.empty(total_tc, np.int32)
        r = _Result(_ptr(op), _ptr(oc), _ptr(ot), _ptr(on))
        self._check(self._L.lagassign_assign_batch(self._ctx, ctypes.byref(b), ctypes.byref(r)))
        self._shape = shape
        return dict(out_partition=op, out_consumer_rank=oc, out_total=ot, out_count=on)

    def assign_batch_narrow(self, topic_off, lag, partition, topic_cset, cset_off, cset_rank,
                            out_partition=None, out_consumer_rank=None):
        """lagassign_assign_batch_narrow(): results as arrays of part_width / cons_width bytes (see lagassign.h).
        Returns dict(out_partition=uint8/16/32 array, out_consumer_rank=uint16/uint32 array or None when cons_width
        is 0, part_width, cons_width). Buffers given by the caller must be uint8 arrays of 4 * N bytes."""
        b, keep, shape = self._batch(topic_off, lag, partition, topic_cset, cset_off, cset_rank)
        T, N, total_tc = shape
        op = np.empty(4 * N, np.uint8) if out_partition is None else out_partition
        oc = np.empty(4 * N, np.uint8) if out_consumer_rank is None else out_consumer_rank
        r = _ResultNarrow(_ptr(op), _ptr(oc), 0, 0)
        self._check(self._L.lagassign_assign_batch_narrow(self._ctx, ctypes.byref(b), ctypes.byref(r)))
        self._shape = shape
        dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}
        part = op[:N * r.part_width].view(dt[r.part_width])
        cons = None if r.cons_width == 0 else oc[:N * r.cons_width].view(dt[r.cons_width])
        return dict(out_partition=part, out_consumer_rank=cons, part_width=int(r.part_width), cons_width=int(r.cons_width))

    def profile(self) -> dict:
        """{kernel name: (milliseconds, launches)} of the last execute() (OPT_PROFILE on)."""
        n = ctypes.c_int(0)
        self._check(self._L.lagassign_profile_count(self._ctx, ctypes.byref(n)))
        out = {}
        for i in range(n.value):
            name = ctypes.c_char_p()
            ms = ctypes.c_float(0)
            k = ctypes.c_int(0)
            self._check(self._L.lagassign_profile_entry(self._ctx, i, ctypes.byref(name), ctypes.byref(ms),
                                                        ctypes.byref(k)))
            if k.value:
                out[name.value.decode()] = (float(ms.value), int(k.value))
        return out

    def stats(self) -> dict:
        s = Stats()
        self._check(self._L.lagassign_get_stats(self._ctx, ctypes.byref(s)))
        return s.as_dict()


def device_count() -> int:
    n = ctypes.c_int(0)
    load_library().lagassign_device_count(ctypes.byref(n))
    return n.value


from .assignor import (LagBasedPartitionAssignor, TopicPartition, TopicPartitionLag,  # noqa: E402
                       compute_partition_lag)
