"""The compiled-out experimental kernel variants, kept honest without a GPU: each macro
is switched on in the host-emulated build of the kernels (tests/emu) and a few shapes that reach the changed
code are checked against the oracle. The default build of the product never defines these macros."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from helpers import random_flat, uniform_batch
import test_kernels_emulated_cpu as tk


def _build(src, name, flags):
    inc = tk._cuda_include()
    if inc is None:
        pytest.skip("CUDA headers (vector_types.h) not found")
    out = os.path.join(tk.EMU, "_build")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, name)
    deps = [os.path.join(tk.EMU, f) for f in (src, "cuda_emu.hpp")] + \
           [os.path.join(tk.CSRC, f) for f in os.listdir(tk.CSRC) if f.endswith(".cuh")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        r = subprocess.run([cxx, "-O1", "-std=c++17", "-pthread", "-fPIC", "-shared", "-w"] + flags +
                           ["-I", tk.EMU, "-I", tk.CSRC, "-I", inc, "-o", lib, os.path.join(tk.EMU, src)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    return ctypes.CDLL(lib)


@pytest.fixture(scope="module")
def onchip_variants():
    L = _build("emu_onchip.cpp", "libemu_onchip_variants.so",
               ["-DLAGASSIGN_WARP_GROUPS", "-DLAGASSIGN_WARP_GENERAL_BY_COPY", "-DLAGASSIGN_PACKED_WARP_FILL",
                "-DLAGASSIGN_FILL_W1", "-DLAGASSIGN_FILL_OPTIMISTIC_W2"])
    L.emu_assign_onchip.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 11
    return L


@pytest.fixture(scope="module")
def radix_variants():
    L = _build("emu_radix.cpp", "libemu_radix_variants.so",
               ["-DLAGASSIGN_HIST_TILE_SKIP", "-DLAGASSIGN_FILL_W1", "-DLAGASSIGN_FILL_OPTIMISTIC_W2",
                "-DLAGASSIGN_BIGFILL_THREADS=1024", "-DLAGASSIGN_PACKED_WARP_FILL"])
    L.emu_assign_big.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 11 + [ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
    return L


def test_four_topics_per_warp(onchip_variants):
    L = onchip_variants
    rng = np.random.default_rng(21)
    before = L.emu_grouped_fast_warps_count()
    tk.check(L, uniform_batch(9, 256, 4096, rng.integers(0, 1 << 40, size=9 * 256, dtype=np.int64)), totals=False)
    tk.check(L, uniform_batch(8, 132, 200, rng.integers(0, 1 << 12, size=8 * 132, dtype=np.int64)), totals=False)
    assert L.emu_grouped_fast_warps_count() - before >= 4          # the fast path ran
    # one topic of a warp whose truncated keys collide: that warp takes the per-topic path, results unchanged
    lag = rng.integers(0, 1 << 40, size=8 * 256, dtype=np.int64)
    lag[3 * 256: 4 * 256] = rng.integers(0, 16, size=256)
    lag[3 * 256 + 7] = 1 << 39
    before = L.emu_grouped_fast_warps_count()
    tk.check(L, uniform_batch(8, 256, 4096, lag), totals=False)
    assert L.emu_grouped_fast_warps_count() - before == 1
    # shapes that never qualify (ragged, rounds, small, explicit ids / totals -> other kernel)
    for ids, totals in ((None, False), ("shuffled", False), (None, True)):
        tk.check(L, random_flat(rng, T=9, pmin=0, pmax=256, C=40, lag_kind="u40", explicit_ids=ids, n_sets=2), totals)


@pytest.mark.parametrize("lag_kind", ["small", "u40", "full"])      # one-word / two-word / three-word consumer sorts
def test_packed_warp_fill(onchip_variants, lag_kind):
    rng = np.random.default_rng(22)
    for C in (3, 31, 64, 200):
        tk.check(onchip_variants, random_flat(rng, T=3, pmin=C + 1, pmax=256, C=C, lag_kind=lag_kind, explicit_ids=None, n_sets=1), True)


@pytest.mark.parametrize("lag_kind", ["small", "u40", "full"])      # exact one-word / exact two-word / optimistic two-word
def test_cta_fill_variants(onchip_variants, lag_kind):
    rng = np.random.default_rng(23)
    for C, P in ((300, 1000), (1100, 4000)):
        tk.check(onchip_variants, random_flat(rng, T=1, pmin=P, pmax=P, C=C, lag_kind=lag_kind, explicit_ids=None, n_sets=1), True)
    # optimistic sort whose truncated totals collide: exact three-word fallback
    lag = np.concatenate([[np.int64(1) << 62], rng.integers(0, 4, size=999)]).astype(np.int64)
    tk.check(onchip_variants, uniform_batch(1, 1000, 300, lag), True)


def test_radix_variants(radix_variants):
    rng = np.random.default_rng(24)
    for C, kind, ids in ((700, "small", None), (1100, "full", "shuffled"), (40, "u40", None)):
        b = random_flat(rng, T=1, pmin=3000, pmax=6000, C=C, lag_kind=kind, explicit_ids=ids, n_sets=1)
        tk.check(radix_variants, b, True, radix_min_p=0)
    for lag in (np.zeros(5000, np.int64), np.concatenate([rng.integers(0, 200, 4999), [1 << 50]]).astype(np.int64)):
        tk.check(radix_variants, uniform_batch(1, 5000, 1200, lag), True, radix_min_p=0)
