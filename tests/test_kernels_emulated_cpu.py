"""The on-chip CUDA kernels, executed on the CPU.

tests/emu/ compiles the product's own device templates (kernels_warp.cuh, kernels_cta.cuh,
fill.cuh, sort_warp.cuh, sort_cta.cuh: k_warp_topics and k_cta_topics<EPT>, unmodified) with g++
against a small stand-in for the CUDA execution model (one host thread per CUDA thread, real
rendezvous for shuffles / reductions / barriers) and this file checks them bit-exactly against
the oracle. It is a logic check that needs no GPU: it covers the sort networks, the adaptive
element widths, the optimistic truncated sort and its fallback, explicit-id handling, the round
fill in all its variants and the store paths. It says nothing about timing, memory-model
behaviour (cp.async is a plain copy here), which only the `-m gpu` tests exercise. The
device-wide path (radix.cuh) runs the same way through tests/emu/emu_radix.cpp, which restates the
big-topic planning and the launch order of lagassign_execute().
"""
import ctypes
import os
import subprocess
import zlib

import numpy as np
import pytest

import oracle
from helpers import args_of, random_flat, uniform_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "kafka-lag-based-assignor_b200", "csrc")


def _cuda_include():
    for d in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if d and os.path.exists(os.path.join(d, "include", "vector_types.h")):
            return os.path.join(d, "include")
    return None


@pytest.fixture(scope="module")
def emu():
    inc = _cuda_include()
    if inc is None:
        pytest.skip("CUDA headers (vector_types.h) not found")
    out = os.path.join(EMU, "_build")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "libemu_onchip.so")
    srcs = [os.path.join(EMU, f) for f in ("emu_onchip.cpp", "cuda_emu.hpp")] + \
           [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(s) for s in srcs):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        r = subprocess.run([cxx, "-O1", "-std=c++17", "-pthread", "-fPIC", "-shared", "-w", "-I", EMU, "-I", CSRC,
                            "-I", inc, "-o", lib, os.path.join(EMU, "emu_onchip.cpp")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    L = ctypes.CDLL(lib)
    L.emu_assign_onchip.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 11
    return L


@pytest.fixture(scope="module")
def emu_radix():
    inc = _cuda_include()
    if inc is None:
        pytest.skip("CUDA headers (vector_types.h) not found")
    out = os.path.join(EMU, "_build")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "libemu_radix.so")
    srcs = [os.path.join(EMU, f) for f in ("emu_radix.cpp", "cuda_emu.hpp")] + \
           [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(s) for s in srcs):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        r = subprocess.run([cxx, "-O1", "-std=c++17", "-pthread", "-fPIC", "-shared", "-w", "-I", EMU, "-I", CSRC,
                            "-I", inc, "-o", lib, os.path.join(EMU, "emu_radix.cpp")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    L = ctypes.CDLL(lib)
    L.emu_assign_big.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 11 + [ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data


def run_emulated(L, b, totals):
    T = len(b["topic_off"]) - 1
    N = int(b["topic_off"][-1])
    C = (b["cset_off"][1:] - b["cset_off"][:-1])[b["topic_cset"]] if T else np.zeros(0, np.int64)
    tc_off = np.concatenate([[0], np.cumsum(C)]).astype(np.int64)
    op = np.full(N, -7, np.int32)
    oc = np.full(N, -7, np.int32)
    ot = np.zeros(int(tc_off[-1]), np.int64) if totals else None      # the library zero-fills these too
    on = np.zeros(int(tc_off[-1]), np.int32) if totals else None
    n = L.emu_assign_onchip(T, _ptr(b["topic_off"]), _ptr(b["lag"]), _ptr(b["partition"]), _ptr(b["topic_cset"]),
                            _ptr(b["cset_off"]), _ptr(b["cset_rank"]), _ptr(tc_off), _ptr(op), _ptr(oc), _ptr(ot), _ptr(on))
    assert n >= 0
    return dict(out_partition=op, out_consumer_rank=oc, out_total=ot, out_count=on)


def run_emulated_radix(L, b, totals, min_big_p, sms=2):
    T = len(b["topic_off"]) - 1
    N = int(b["topic_off"][-1])
    C = (b["cset_off"][1:] - b["cset_off"][:-1])[b["topic_cset"]] if T else np.zeros(0, np.int64)
    tc_off = np.concatenate([[0], np.cumsum(C)]).astype(np.int64)
    op = np.full(N, -7, np.int32)
    oc = np.full(N, -7, np.int32)
    ot = np.zeros(int(tc_off[-1]), np.int64) if totals else None
    on = np.zeros(int(tc_off[-1]), np.int32) if totals else None
    passes = ctypes.c_int(0)
    n = L.emu_assign_big(T, _ptr(b["topic_off"]), _ptr(b["lag"]), _ptr(b["partition"]), _ptr(b["topic_cset"]),
                         _ptr(b["cset_off"]), _ptr(b["cset_rank"]), _ptr(tc_off), _ptr(op), _ptr(oc), _ptr(ot), _ptr(on),
                         min_big_p, sms, ctypes.byref(passes))
    assert n >= 0
    return dict(out_partition=op, out_consumer_rank=oc, out_total=ot, out_count=on), passes.value


def check(L, b, totals=True, radix_min_p=None):
    if radix_min_p is None:
        got = run_emulated(L, b, totals)
    else:
        got, _ = run_emulated_radix(L, b, totals, radix_min_p)
    exp = oracle.assign_batch(*args_of(b), want_totals=True, threads=4)
    off, cs = b["topic_off"], b["topic_cset"]
    for t in range(len(off) - 1):
        lo, hi = int(off[t]), int(off[t + 1])
        C = int(b["cset_off"][cs[t] + 1] - b["cset_off"][cs[t]])
        if C == 0:                                    # left to k_unassigned_topics in the product
            assert (got["out_consumer_rank"][lo:hi] == -7).all()
            continue
        assert np.array_equal(got["out_partition"][lo:hi], exp["out_partition"][lo:hi]), (t, hi - lo, C, "partition")
        assert np.array_equal(got["out_consumer_rank"][lo:hi], exp["out_consumer_rank"][lo:hi]), (t, hi - lo, C, "consumer")
    if totals:
        assert np.array_equal(got["out_total"], exp["out_total"])
        assert np.array_equal(got["out_count"], exp["out_count"])


@pytest.mark.parametrize("lag_kind", ["small", "dups", "u40", "full"])
@pytest.mark.parametrize("ids", [None, "sorted", "shuffled", "dups"])
def test_warp_kernel_on_cpu(emu, lag_kind, ids):
    rng = np.random.default_rng(zlib.crc32(repr((lag_kind, ids)).encode()))
    for C, totals in ((1, True), (7, False), (40, True), (300, ids is not None)):
        b = random_flat(rng, T=4, pmin=0, pmax=256, C=C, lag_kind=lag_kind, explicit_ids=ids, n_sets=2)
        check(emu, b, totals)


def test_warp_kernel_every_boundary_size(emu):
    rng = np.random.default_rng(5)
    sizes = [1, 2, 3, 31, 32, 33, 63, 64, 65, 127, 128, 129, 200, 255, 256]
    for C in (3, 64):
        P = np.array(sizes)
        topic_off = np.concatenate([[0], np.cumsum(P)]).astype(np.int64)
        N = int(topic_off[-1])
        b = dict(topic_off=topic_off, lag=rng.integers(0, 1 << 20, size=N, dtype=np.int64), partition=None,
                 topic_cset=np.zeros(len(P), np.int32), cset_off=np.array([0, C], np.int64),
                 cset_rank=np.arange(C, dtype=np.int32))
        check(emu, b, totals=False)      # PLAIN variant
        check(emu, b, totals=True)


def test_warp_kernel_aligned_and_unaligned_stores(emu):
    """uniform batches of P % 4 == 0 take the 16-byte stores from registers; P = 255 the staged path."""
    rng = np.random.default_rng(6)
    for P, C in ((256, 4096), (252, 300), (255, 300), (64, 64), (100, 16)):
        b = uniform_batch(5, P, C, rng.integers(0, 1 << 40, size=5 * P, dtype=np.int64))
        check(emu, b, totals=False)


def test_warp_kernel_special_lags(emu):
    rng = np.random.default_rng(7)
    P, C = 200, 9
    for lag in (np.zeros(P, np.int64),                                         # all equal: sort skipped
                np.full(P, -5, np.int64),
                rng.integers(-(1 << 63), (1 << 63) - 1, size=P, dtype=np.int64),   # general three-word sort, totals wrap
                (np.int64(1) << 62) + rng.integers(0, 3, size=P, dtype=np.int64),   # wrap-around of the totals
                np.where(rng.random(P) < 0.5, 0, rng.integers(0, 1 << 33, size=P, dtype=np.int64)),
                (rng.integers(0, 1 << 40, size=P, dtype=np.int64) >> 17) << 17):    # truncated sort sees duplicates: fallback
        check(emu, uniform_batch(1, P, C, lag.astype(np.int64)), totals=True)


@pytest.mark.parametrize("case", [
    # (pmin, pmax, C, lag kind, ids)          size class / fill variant
    (257, 512, 600, "u40", None),             # EPT 2, P <= C
    (257, 512, 40, "small", "shuffled"),      # EPT 2, warp fill, ids put in order first
    (513, 1024, 300, "dups", None),           # EPT 4, CTA fill with 2 consumers per thread
    (513, 1024, 700, "full", "sorted"),       # EPT 4, CTA fill EPT 4, three-word sort
    (1025, 2048, 1500, "u40", None),          # EPT 8, CTA fill EPT 8
    (1025, 2048, 5, "small", "dups"),         # EPT 8, warp fill, duplicate ids
    (2049, 4096, 3000, "small", None),        # EPT 16, CTA fill EPT 16
    (2049, 4096, 4096, "u40", None),          # EPT 16, P <= C
])
def test_cta_kernels_on_cpu(emu, case):
    pmin, pmax, C, kind, ids = case
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()))
    b = random_flat(rng, T=2, pmin=pmin, pmax=pmax, C=C, lag_kind=kind, explicit_ids=ids, n_sets=1)
    check(emu, b, totals=True)


def test_cta_kernel_idle_topic_and_equal_rounds(emu):
    """All-zero lags: the sort is skipped and the rounds keep the consumer order (sortedness test)."""
    b = uniform_batch(2, 1000, 300, np.zeros(2000, np.int64))
    check(emu, b, totals=True)
    lag = np.concatenate([np.arange(500, 0, -1), np.zeros(500)]).astype(np.int64)    # busy head, idle tail
    check(emu, uniform_batch(1, 1000, 300, lag), totals=True)


@pytest.mark.parametrize("case", [
    # (T, pmin, pmax, C, lag kind, ids, totals, threshold)
    (1, 4097, 6000, 1200, "small", None, True, 4096),        # few digit passes, multi-round fill in registers
    (3, 1, 3000, 50, "u40", None, False, 0),                 # one-tile topics, short last tiles, warp-sized fill
    (1, 5100, 9000, 5000, "full", "shuffled", True, 4096),   # all 8 lag passes + id tie passes, P > C with C > 4096 (shared-memory fill)
    (1, 9000, 9000, 300, "dups", "dups", True, 4096),        # duplicate ids and lags
    (2, 100, 9000, 20000, "small", "sorted", False, 0),      # P <= C: no fill kernel; ids given but already ascending
])
def test_radix_path_on_cpu(emu_radix, case):
    T, pmin, pmax, C, kind, ids, totals, thr = case
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()))
    b = random_flat(rng, T=T, pmin=pmin, pmax=pmax, C=C, lag_kind=kind, explicit_ids=ids, n_sets=1)
    check(emu_radix, b, totals, radix_min_p=thr)


def test_radix_path_skips_constant_digits(emu_radix):
    """All-equal lags run zero passes (order = input order, closed-form fill); lags below 2^16 run two."""
    b = uniform_batch(1, 6000, 64, np.full(6000, 42, np.int64))
    got, passes = run_emulated_radix(emu_radix, b, True, 4096)
    assert passes == 0
    check(emu_radix, b, True, radix_min_p=4096)
    rng = np.random.default_rng(11)
    b = uniform_batch(1, 6000, 7000, rng.integers(0, 1 << 16, size=6000, dtype=np.int64))
    got, passes = run_emulated_radix(emu_radix, b, False, 4096)
    assert passes == 2
    check(emu_radix, b, False, radix_min_p=4096)


def test_radix_certificate_catches_unordered_atomics(emu_radix, monkeypatch):
    """The fast ranking of k_big_scatter (one shared-memory atomic per element) is stable only if the lanes of
    one instruction are served in lane order. The emulator's lanes are free-running host threads, so with
    EMU_FAST_RANK=1 that order is arbitrary: k_big_verify (the certificate lagassign_execute() reads before it
    trusts the sort) must then either report the failure (-2) or the result must be exact anyway."""
    monkeypatch.setenv("EMU_FAST_RANK", "1")
    rng = np.random.default_rng(77)
    caught = 0
    for _ in range(4):
        b = random_flat(rng, T=3, pmin=3000, pmax=9000, C=40, lag_kind="dups", n_sets=1, full_sets=True)
        T = len(b["topic_off"]) - 1
        N = int(b["topic_off"][-1])
        C = (b["cset_off"][1:] - b["cset_off"][:-1])[b["topic_cset"]]
        tc_off = np.concatenate([[0], np.cumsum(C)]).astype(np.int64)
        # two executed passes are needed for instability to matter: lags 0..3 in the low byte, 0..3 in the next
        b["lag"] = (b["lag"] + 256 * rng.integers(0, 4, size=N)).astype(np.int64)
        op = np.full(N, -7, np.int32); oc = np.full(N, -7, np.int32)
        n = emu_radix.emu_assign_big(T, _ptr(b["topic_off"]), _ptr(b["lag"]), None, _ptr(b["topic_cset"]),
                                     _ptr(b["cset_off"]), _ptr(b["cset_rank"]), _ptr(tc_off), _ptr(op), _ptr(oc), None, None,
                                     2000, 2, None)
        if n == -2:
            caught += 1
            continue
        exp = oracle.assign_batch(*args_of(b), threads=4)
        assert np.array_equal(op, exp["out_partition"]) and np.array_equal(oc, exp["out_consumer_rank"])
    assert caught > 0, "host threads never reordered same-address atomics?"


def test_radix_path_more_than_8192_consumers_and_more_partitions(emu_radix):
    """P > C > FILL_CTA_MAX_C (8192): the round fill runs in a global-memory workspace (cta_round_fill_any).
    Round 1 rejected this shape with LAGASSIGN_ERR_UNSUPPORTED."""
    rng = np.random.default_rng(91)
    b = uniform_batch(1, 9000 * 2 + 77, 9000, rng.integers(0, 1 << 30, size=9000 * 2 + 77, dtype=np.int64))
    check(emu_radix, b, True, radix_min_p=0)


# ---- fused computePartitionLag (Main:376-404): the kernels build their keys from raw offsets -------------------

def _offsets_for(rng, n):
    end = rng.integers(0, 1 << 40, size=n, dtype=np.int64)
    begin = (end * rng.random(n)).astype(np.int64)
    committed = rng.integers(0, 1 << 40, size=n, dtype=np.int64)      # sometimes beyond end -> clamped to 0
    has = (rng.random(n) < 0.7).astype(np.uint8)
    # a few wrapping cases: end - committed overflows int64 and wraps like Java's long subtraction
    k = min(n, 5)
    end[:k] = np.iinfo(np.int64).max - np.arange(k)
    committed[:k] = -(1 << 62) - np.arange(k)
    has[:k] = 1
    return begin, end, committed, has


def _lags_of(begin, end, committed, has, latest):
    return np.array([oracle.compute_partition_lag(int(committed[i]) if has[i] else None, int(begin[i]), int(end[i]),
                                                  "latest" if latest else "earliest") for i in range(len(end))], dtype=np.int64)


@pytest.mark.parametrize("latest", [0, 1])
@pytest.mark.parametrize("ids,totals", [(None, False), ("shuffled", True)])
def test_onchip_kernels_from_raw_offsets(emu, latest, ids, totals):
    """k_warp_topics<PLAIN, LAG_OFFS>, k_warp_topics<false, LAG_EITHER> and k_cta_topics evaluate the lag formula
    themselves: the lag array handed to them is poisoned."""
    rng = np.random.default_rng(5 + latest)
    b = random_flat(rng, T=14, pmin=1, pmax=700, C=40, lag_kind="u40", explicit_ids=ids, n_sets=2)
    n = len(b["lag"])
    begin, end, committed, has = _offsets_for(rng, n)
    exp_lag = _lags_of(begin, end, committed, has, latest)
    poisoned = dict(b, lag=np.full(n, 0x5555555555555555, np.int64))
    emu.emu_onchip_set_offsets.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int]
    emu.emu_onchip_set_offsets(_ptr(begin), _ptr(end), _ptr(committed), _ptr(has), latest)
    got = run_emulated(emu, poisoned, totals)
    exp = oracle.assign_batch(*args_of(dict(b, lag=exp_lag)), want_totals=True, threads=4)
    assert np.array_equal(got["out_partition"], exp["out_partition"])
    assert np.array_equal(got["out_consumer_rank"], exp["out_consumer_rank"])
    if totals:
        assert np.array_equal(got["out_total"], exp["out_total"]) and np.array_equal(got["out_count"], exp["out_count"])


def test_radix_path_from_raw_offsets(emu_radix):
    """k_big_range computes the lags from the offsets while it builds the keys and stores them for the later passes."""
    rng = np.random.default_rng(17)
    b = random_flat(rng, T=2, pmin=4200, pmax=7000, C=900, lag_kind="u40", n_sets=1)
    n = len(b["lag"])
    begin, end, committed, has = _offsets_for(rng, n)
    exp_lag = _lags_of(begin, end, committed, has, 0)
    emu_radix.emu_big_set_offsets.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int]
    emu_radix.emu_big_set_offsets(_ptr(begin), _ptr(end), _ptr(committed), _ptr(has), 0)
    got, _ = run_emulated_radix(emu_radix, dict(b, lag=np.full(n, -1, np.int64)), True, 4096)
    exp = oracle.assign_batch(*args_of(dict(b, lag=exp_lag)), want_totals=True, threads=4)
    for k in ("out_partition", "out_consumer_rank", "out_total", "out_count"):
        assert np.array_equal(got[k], exp[k]), k
