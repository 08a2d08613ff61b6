"""Fuzz of the on-chip kernels under the host emulation (tests/emu) against the oracle: random shapes, lag
kinds, id layouts and consumer counts. Needs no GPU.  usage: python scripts/emu_fuzz.py <first seed> <cases> [radix]   ("radix": the device-wide path, every topic routed through it)
(run python -m pytest tests/test_kernels_emulated_cpu.py once first: it builds tests/emu/_build/libemu_onchip.so)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, ctypes
import oracle
from helpers import random_flat, args_of
import test_kernels_emulated_cpu as tk
seed0 = int(sys.argv[1]); n = int(sys.argv[2])
t0 = time.time(); fails = 0
if len(sys.argv) > 3 and sys.argv[3] == "radix":
    R = ctypes.CDLL(os.environ.get("EMU_LIB") or os.path.join(ROOT, "tests", "emu", "_build", "libemu_radix.so"))
    R.emu_assign_big.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 11 + [ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
    for s in range(seed0, seed0 + n):
        rng = np.random.default_rng(s)
        kind = ["small", "dups", "u40", "full"][int(rng.integers(0, 4))]
        ids = [None, None, "sorted", "shuffled", "dups"][int(rng.integers(0, 5))]
        C = int(rng.choice([1, 3, 40, 256, 300, 700, 1200, 3000, 5000, 9000]))
        b = random_flat(rng, T=int(rng.integers(1, 4)), pmin=1, pmax=int(rng.choice([300, 4096, 4097, 9000, 13000])), C=C,
                        lag_kind=kind, explicit_ids=ids, n_sets=1)
        try:
            tk.check(R, b, bool(rng.random() < 0.5), radix_min_p=0)
        except AssertionError as e:
            fails += 1
            print("FAIL seed", s, kind, ids, C, str(e)[:200], flush=True)
    print(f"radix seeds {seed0}..{seed0+n-1}: {fails} failures in {time.time()-t0:.0f} s", flush=True)
    sys.exit(1 if fails else 0)
L = ctypes.CDLL(os.environ.get("EMU_LIB") or os.path.join(ROOT, "tests", "emu", "_build", "libemu_onchip.so"))   # EMU_LIB: a variant build
L.emu_assign_onchip.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 11
for s in range(seed0, seed0 + n):
    rng = np.random.default_rng(s)
    kind = ["small", "dups", "u40", "full"][int(rng.integers(0, 4))]
    ids = [None, None, "sorted", "shuffled", "dups"][int(rng.integers(0, 5))]
    big = rng.random() < 0.25
    pmax = int(rng.choice([300, 600, 1100, 2100, 4096])) if big else int(rng.choice([8, 33, 70, 130, 256]))
    C = int(rng.choice([1, 2, 5, 31, 32, 33, 64, 100, 255, 256, 257, 500, 513, 1000, 1025, 2047, 2049, 4000, 5000]))
    T = 2 if big else int(rng.integers(1, 12))
    b = random_flat(rng, T=T, pmin=0, pmax=pmax, C=C, lag_kind=kind, explicit_ids=ids, n_sets=int(rng.integers(1, 4)))
    totals = bool(rng.random() < 0.5)
    try:
        tk.check(L, b, totals)
    except AssertionError as e:
        fails += 1
        print("FAIL seed", s, kind, ids, pmax, C, T, totals, str(e)[:200], flush=True)
print(f"seeds {seed0}..{seed0+n-1}: {fails} failures in {time.time()-t0:.0f} s", flush=True)
