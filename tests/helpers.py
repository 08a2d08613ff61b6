"""Input generators shared by the CPU and GPU tests."""
from __future__ import annotations

import numpy as np

MASK64 = (1 << 64) - 1


def splitmix64(idx: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser of (uint64) idx, vectorised."""
    with np.errstate(over="ignore"):
        z = (idx.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def random_flat(rng, T, pmin, pmax, C, lag_kind="u40", explicit_ids=None, n_sets=1, full_sets=False):
    """A random batch in the flat layout of include/lagassign.h.

    explicit_ids: None -> partition=NULL; "sorted" -> ascending explicit ids;
    "shuffled" -> a permutation per topic; "dups" -> random ids with repeats.
    """
    P = rng.integers(pmin, pmax + 1, size=T)
    topic_off = np.concatenate([[0], np.cumsum(P)]).astype(np.int64)
    N = int(topic_off[-1])
    if lag_kind == "full":
        lag = rng.integers(-(1 << 63), (1 << 63) - 1, size=N, dtype=np.int64, endpoint=True)
    elif lag_kind == "dups":
        lag = rng.integers(0, 4, size=N, dtype=np.int64)
    elif lag_kind == "zero":
        lag = np.zeros(N, dtype=np.int64)
    elif lag_kind == "small":
        lag = rng.integers(0, 1 << 20, size=N, dtype=np.int64)
    elif lag_kind == "neg":
        lag = rng.integers(-1000, 1000, size=N, dtype=np.int64)
    elif lag_kind == "bigpos":
        lag = rng.integers((1 << 62), (1 << 63) - 1, size=N, dtype=np.int64)
    else:
        lag = rng.integers(0, 1 << 40, size=N, dtype=np.int64)
    part = None
    if explicit_ids == "sorted":
        part = np.concatenate([np.sort(rng.choice(max(int(p) * 3, 1), size=int(p), replace=False)) for p in P]
                              + [np.zeros(0, np.int64)]).astype(np.int32)
    elif explicit_ids == "shuffled":
        part = np.concatenate([rng.permutation(int(p)) for p in P] + [np.zeros(0, np.int64)]).astype(np.int32)
    elif explicit_ids == "dups":
        part = np.concatenate([rng.integers(0, max(int(p) // 2, 1), size=int(p)) for p in P]
                              + [np.zeros(0, np.int64)]).astype(np.int32)
    sets = []
    for _ in range(n_sets):
        c = C if full_sets else int(rng.integers(0, C + 1))
        sets.append(np.sort(rng.choice(C, size=c, replace=False)).astype(np.int32))
    cset_off = np.concatenate([[0], np.cumsum([len(s) for s in sets])]).astype(np.int64)
    cset_rank = np.concatenate(sets + [np.zeros(0, np.int32)]).astype(np.int32)
    topic_cset = rng.integers(0, n_sets, size=T).astype(np.int32)
    return dict(topic_off=topic_off, lag=lag, partition=part, topic_cset=topic_cset,
                cset_off=cset_off, cset_rank=cset_rank)


def uniform_batch(T, P, C, lag):
    """T topics of P partitions each, one consumer set of C members (ranks 0..C-1)."""
    return dict(topic_off=(np.arange(T + 1, dtype=np.int64) * P), lag=np.ascontiguousarray(lag, dtype=np.int64),
                partition=None, topic_cset=np.zeros(T, np.int32),
                cset_off=np.array([0, C], np.int64), cset_rank=np.arange(C, dtype=np.int32))


def baseline_config(n, scale=1.0):
    """The BASELINE.json configs, optionally scaled down in
    topic count / partition count for tests. Returns (batch dict, description)."""
    if n == 1:   # README example
        return uniform_batch(1, 3, 2, [100000, 50000, 60000]), "README: 1 topic x 3 partitions, 2 consumers"
    if n == 2:
        P = max(1, int((1 << 20) * scale))
        C = 1024
        lag = (splitmix64(np.uint64(0xB200) + np.arange(P, dtype=np.uint64)) >> np.uint64(1)).astype(np.int64)
        return uniform_batch(1, P, C, lag), f"1 topic x {P} partitions, {C} consumers, uniform int64 lag"
    if n == 3:
        T = max(1, int(10000 * scale))
        P, C = 64, 256
        z = np.random.default_rng(3).zipf(1.1, T * P)
        lag = np.minimum(z, 1 << 62).astype(np.int64)
        return uniform_batch(T, P, C, lag), f"{T} topics x {P} partitions, {C} consumers, Zipf(1.1) lag"
    if n == 4:
        T = max(1, int(100000 * scale))
        P, C = 256, 4096
        lag = (splitmix64(np.uint64(4 << 32) + np.arange(T * P, dtype=np.uint64)) >> np.uint64(24)).astype(np.int64)
        return uniform_batch(T, P, C, lag), f"{T} topics x {P} partitions, {C} consumers, uniform [0,2^40) lag"
    if n == 5:
        T = max(2, int(1000 * scale))
        u = np.random.default_rng(5).random(T)
        P = np.floor(2.0 ** (20.0 * u)).astype(np.int64)
        P[0] = 1
        P[1] = 1 << 20
        if scale < 1.0:
            P = np.maximum(1, (P * scale).astype(np.int64))
            P[1] = max(1, int((1 << 20) * scale))
        C = 1024
        topic_off = np.concatenate([[0], np.cumsum(P)]).astype(np.int64)
        N = int(topic_off[-1])
        return dict(topic_off=topic_off, lag=np.zeros(N, np.int64), partition=None,
                    topic_cset=np.zeros(T, np.int32), cset_off=np.array([0, C], np.int64),
                    cset_rank=np.arange(C, dtype=np.int32)), \
            f"{T} topics, power-law partition counts (N={N}), {C} consumers, all-zero lag"
    if n == 6:
        # SURVEY.md 8(d) "scaled sort-only variant": the device-wide radix sort with a
        # working set far beyond the 126 MB L2; P <= C so the fill is the trivial round.
        T = max(1, int(64 * scale))
        P = 1 << 20
        lag = (splitmix64(np.uint64(0xB200) + np.arange(T * P, dtype=np.uint64)) >> np.uint64(1)).astype(np.int64)
        return uniform_batch(T, P, P, lag), f"{T} topics x {P} partitions, {P} consumers, uniform int64 lag (sort-only variant)"
    if n == 7:
        # not a BASELINE config: a "typical deployment" shape for the multi-round fill — many
        # topics with more partitions than consumers, small lags, half of the partitions idle
        T = max(1, int(20000 * scale))
        P, C = 200, 16
        rng = np.random.default_rng(7)
        lag = rng.integers(0, 1 << 20, size=T * P, dtype=np.int64)
        lag[rng.random(T * P) < 0.5] = 0
        return uniform_batch(T, P, C, lag), f"{T} topics x {P} partitions, {C} consumers, small lags, 50% zero (multi-round fill)"
    raise ValueError(n)


def args_of(b):
    return (b["topic_off"], b["lag"], b["partition"], b["topic_cset"], b["cset_off"], b["cset_rank"])
