"""CPU oracle for the lag-based assignment hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package; the product package never does.

``oracle.c`` is the C restatement of the reference (LagBasedPartitionAssignor
.java:166-188, :204-277, :376-404, :410-426); ``oracle_py.py`` is a second,
independently written pure-Python restatement used to cross-check it on small
cases. Both are pinned against the reference's own test vectors in
tests/golden/ (the reference is a JVM class and no JDK exists in this image, so
it cannot be run here).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

LITERAL = 0
ROUNDS = 1


def build(force: bool = False) -> str:
    """Compile oracle.c into oracle/liboracle.so (gcc, OpenMP if available)."""
    src = os.path.join(_HERE, "oracle.c")
    hdr = os.path.join(_HERE, "oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    base = ["-O3", "-march=x86-64-v3", "-fPIC", "-std=c11", "-D_GNU_SOURCE", "-shared",
            "-o", _LIB_PATH, src]
    last = None
    for cc in ("/usr/bin/gcc", "gcc", "cc"):
        for omp in (["-fopenmp"], []):
            try:
                subprocess.run([cc] + omp + base, check=True, capture_output=True, cwd=_HERE)
                return _LIB_PATH
            except (OSError, subprocess.CalledProcessError) as e:  # try the next variant
                last = e
    raise RuntimeError(f"could not build the oracle: {last}")


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        p = ctypes.c_void_p
        L.oracle_assign_topic.restype = ctypes.c_int
        L.oracle_assign_topic.argtypes = [ctypes.c_int, ctypes.c_int64, p, p, ctypes.c_int64, p, p, p, p, p]
        L.oracle_assign_batch.restype = ctypes.c_int64
        L.oracle_assign_batch.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, p, p, p, p,
                                          ctypes.c_int64, p, p, ctypes.c_int64, ctypes.c_int64, p, p, p, p]
        L.oracle_compute_partition_lag.restype = ctypes.c_int64
        L.oracle_compute_partition_lag.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                                   ctypes.c_int64, ctypes.c_char_p]
        L.oracle_compare_member_ids.restype = ctypes.c_int
        L.oracle_compare_member_ids.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        L.oracle_rank_members.restype = ctypes.c_int
        L.oracle_rank_members.argtypes = [ctypes.c_int64, p, p]
        L.oracle_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def assign_batch(topic_off, lag, partition, topic_cset, cset_off, cset_rank, *,
                 mode: int = ROUNDS, threads: int = 1, want_totals: bool = False,
                 topic_lo: int = 0, topic_hi: int | None = None):
    """Flat-array oracle with the layout of include/lagassign.h.

    Returns dict(out_partition, out_consumer_rank, out_total, out_count, assigned).
    """
    topic_off = _i64(topic_off)
    lag = _i64(lag)
    partition = _i32(partition)
    topic_cset = _i32(topic_cset)
    cset_off = _i64(cset_off)
    cset_rank = _i32(cset_rank)
    T = len(topic_off) - 1
    N = int(topic_off[-1]) if T >= 0 else 0
    S = len(cset_off) - 1
    out_p = np.full(N, -2, dtype=np.int32)
    out_c = np.full(N, -2, dtype=np.int32)
    out_t = out_n = None
    if want_totals:
        sizes = (cset_off[1:] - cset_off[:-1])[topic_cset] if T > 0 else np.zeros(0, np.int64)
        tot = int(sizes.sum())
        out_t = np.zeros(tot, dtype=np.int64)
        out_n = np.zeros(tot, dtype=np.int32)
    hi = T if topic_hi is None else topic_hi
    r = lib().oracle_assign_batch(mode, threads, T, _ptr(topic_off), _ptr(lag), _ptr(partition),
                                  _ptr(topic_cset), S, _ptr(cset_off), _ptr(cset_rank),
                                  topic_lo, hi, _ptr(out_p), _ptr(out_c), _ptr(out_t), _ptr(out_n))
    if r < 0:
        raise MemoryError("oracle_assign_batch failed")
    return dict(out_partition=out_p, out_consumer_rank=out_c, out_total=out_t, out_count=out_n,
                assigned=int(r))


def compute_partition_lag(committed, begin_offset: int, end_offset: int, reset_mode: str) -> int:
    """computePartitionLag (Main:376-404); ``committed`` None = no committed offset."""
    has = 0 if committed is None else 1
    return int(lib().oracle_compute_partition_lag(has, 0 if committed is None else committed,
                                                  begin_offset, end_offset, reset_mode.encode()))


def compare_member_ids(a: str, b: str) -> int:
    """String.compareTo sign for two member ids (UTF-16 code-unit order)."""
    r = lib().oracle_compare_member_ids(a.encode("utf-8", "surrogatepass"),
                                        b.encode("utf-8", "surrogatepass"))
    return (r > 0) - (r < 0)


def rank_members(ids):
    enc = [s.encode("utf-8") for s in ids]
    arr = (ctypes.c_char_p * len(enc))(*enc)
    out = np.zeros(len(enc), dtype=np.int32)
    rc = lib().oracle_rank_members(len(enc), ctypes.cast(arr, ctypes.c_void_p), _ptr(out))
    if rc != 0:
        raise MemoryError("oracle_rank_members failed")
    return out


def max_threads() -> int:
    return int(lib().oracle_max_threads())
