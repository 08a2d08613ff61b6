"""Element forms of the round fill (fill.cuh) under the host emulation (tests/emu): the one-word / two-word /
three-word consumer sorts of the warp fill and of the CTA / device-wide fill, chosen at run time from the spread
of a round's totals. Each shape below reaches one of them; results are checked against the oracle. (These were
compile-time variants in round 1; the ones that won their A/B on a B200 are now the default code.)"""
import numpy as np
import pytest

from helpers import random_flat, uniform_batch
import test_kernels_emulated_cpu as tk

emu = tk.emu
emu_radix = tk.emu_radix


@pytest.mark.parametrize("lag_kind", ["small", "u40", "full"])      # one-word / two-word / three-word consumer sorts
def test_packed_warp_fill(emu, lag_kind):
    rng = np.random.default_rng(22)
    for C in (3, 31, 64, 200):
        tk.check(emu, random_flat(rng, T=3, pmin=C + 1, pmax=256, C=C, lag_kind=lag_kind, explicit_ids=None, n_sets=1), True)


@pytest.mark.parametrize("lag_kind", ["small", "u40", "full"])      # exact one-word / exact two-word / three-word
def test_cta_fill_forms(emu, lag_kind):
    rng = np.random.default_rng(23)
    for C, P in ((300, 1000), (1100, 4000)):
        tk.check(emu, random_flat(rng, T=1, pmin=P, pmax=P, C=C, lag_kind=lag_kind, explicit_ids=None, n_sets=1), True)
    # one huge lag among tiny ones: the spread of the totals changes form between rounds
    lag = np.concatenate([[np.int64(1) << 62], rng.integers(0, 4, size=999)]).astype(np.int64)
    tk.check(emu, uniform_batch(1, 1000, 300, lag), True)


def test_device_fill_forms(emu_radix):
    rng = np.random.default_rng(24)
    for C, kind, ids in ((700, "small", None), (1100, "full", "shuffled"), (40, "u40", None)):
        b = random_flat(rng, T=1, pmin=3000, pmax=6000, C=C, lag_kind=kind, explicit_ids=ids, n_sets=1)
        tk.check(emu_radix, b, True, radix_min_p=0)
    for lag in (np.zeros(5000, np.int64), np.concatenate([rng.integers(0, 200, 4999), [1 << 50]]).astype(np.int64)):
        tk.check(emu_radix, uniform_batch(1, 5000, 1200, lag), True, radix_min_p=0)
