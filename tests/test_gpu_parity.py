"""Parity of the CUDA path (through the C ABI) with the oracle: bit-exact.

Layout of this file follows the reference's own test class
(LagBasedPartitionAssignorTest.java): first its vectors through the plugin-shaped
host mirror, then generated cases per kernel bin, then BASELINE-shaped configs.
"""
import json
import os
import zlib

import numpy as np
import pytest

import oracle
from helpers import args_of, baseline_config, random_flat, uniform_batch

pytestmark = pytest.mark.gpu

KEYS = ("out_partition", "out_consumer_rank", "out_total", "out_count")


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__
    __graft_entry__.build()
    import kafka_lag_based_assignor_b200 as p
    return p


@pytest.fixture(scope="module")
def flat(pkg):
    with pkg.FlatAssignor(0, want_totals=True) as f:
        yield f


def check(flat, batch, mode=oracle.ROUNDS, threads=8):
    got = flat.assign_batch(*args_of(batch))
    exp = oracle.assign_batch(*args_of(batch), mode=mode, want_totals=True, threads=threads)
    for k in KEYS:
        if not np.array_equal(got[k], exp[k]):
            bad = np.nonzero(got[k] != exp[k])[0]
            raise AssertionError(f"{k}: {len(bad)} mismatches, first at {bad[:5]}: "
                                 f"got {got[k][bad[:5]]} expected {exp[k][bad[:5]]}; stats={flat.stats()}")
    return got


# ---- the reference's own vectors, written like its tests -----------------------------------

def _tpl(pkg, topic, lags, parts=None):
    parts = range(len(lags)) if parts is None else parts
    return [pkg.TopicPartitionLag(topic, p, l) for p, l in zip(parts, lags)]


def test_assign_reference_vector(pkg):
    """LagBasedPartitionAssignorTest.testAssign (Test:82-132), list order included."""
    partition_lag_per_topic = {
        "topic1": _tpl(pkg, "topic1", [100000, 100000, 500, 1]),
        "topic2": _tpl(pkg, "topic2", [900000, 100000]),
    }
    subscriptions = {"consumer-1": ["topic1", "topic2"], "consumer-2": ["topic1"]}
    expected = {
        "consumer-1": [("topic1", 0), ("topic1", 2), ("topic2", 0), ("topic2", 1)],
        "consumer-2": [("topic1", 1), ("topic1", 3)],
    }
    actual = pkg.LagBasedPartitionAssignor().assign(partition_lag_per_topic, subscriptions)
    assert actual == expected


def test_assign_readme_example(pkg):
    """README.md:42-57 / Main:57-69."""
    a = pkg.LagBasedPartitionAssignor()
    actual = a.assign({"t0": _tpl(pkg, "t0", [100000, 50000, 60000])}, {"C0": ["t0"], "C1": ["t0"]})
    assert {m: sorted(v) for m, v in actual.items()} == {"C0": [("t0", 0)], "C1": [("t0", 1), ("t0", 2)]}
    assert actual["C1"] == [("t0", 2), ("t0", 1)]     # assignment order is lag-descending


def test_assign_with_zero_lags(pkg):
    """testAssignWithZeroLags (Test:134-175): max <= min + 1."""
    actual = pkg.LagBasedPartitionAssignor().assign(
        {"topic1": _tpl(pkg, "topic1", [0] * 7)},
        {"consumer-1": ["topic1"], "consumer-2": ["topic1"]})
    sizes = [len(v) for v in actual.values()]
    assert max(sizes) <= min(sizes) + 1 and sum(sizes) == 7


def test_assign_with_heavily_skewed_lags(pkg):
    """testAssignWithHeavilySkewedLags (Test:177-228)."""
    actual = pkg.LagBasedPartitionAssignor().assign(
        {"topic1": _tpl(pkg, "topic1", [360, 359, 230, 118, 444, 122, 65, 111, 455000, 424000])},
        {"consumer-1": ["topic1"], "consumer-2": ["topic1"], "consumer-3": ["topic1"]})
    sizes = [len(v) for v in actual.values()]
    assert max(sizes) <= min(sizes) + 1 and sum(sizes) == 10


def test_golden_files_through_host_mirror(pkg, golden_dir):
    a = pkg.LagBasedPartitionAssignor()
    kats = json.load(open(os.path.join(golden_dir, "reference_kats.json")))["assign"]
    reg = json.load(open(os.path.join(golden_dir, "oracle_regression.json")))["cases"]
    for case in kats + reg:
        lags = {t: [pkg.TopicPartitionLag(*e) for e in v] for t, v in case["partition_lag_per_topic"].items()}
        subs = {m: ts for m, ts in case["subscriptions"]}
        got = a.assign(lags, subs)
        exp = {m: [tuple(tp) for tp in v] for m, v in case["expected"].items()}
        if case.get("order_sensitive", True):
            assert got == exp, case["name"]
        else:
            assert {m: sorted(v) for m, v in got.items()} == {m: sorted(v) for m, v in exp.items()}, case["name"]


# ---- generated cases per kernel bin ----------------------------------------------------------

LAG_KINDS = ["u40", "full", "dups", "zero", "neg", "bigpos"]


@pytest.mark.parametrize("lag_kind", LAG_KINDS)
@pytest.mark.parametrize("ids", [None, "sorted", "shuffled", "dups"])
def test_warp_bin(flat, lag_kind, ids):
    """P <= 256: one warp per topic; consumer sets from empty to larger than P."""
    rng = np.random.default_rng(zlib.crc32(repr((lag_kind, ids)).encode()))
    for C in (1, 2, 5, 31, 32, 33, 100, 255, 256, 300):
        b = random_flat(rng, T=60, pmin=0, pmax=256, C=C, lag_kind=lag_kind, explicit_ids=ids, n_sets=4)
        check(flat, b, mode=oracle.LITERAL)


def test_warp_bin_every_size(flat):
    rng = np.random.default_rng(11)
    sizes = np.arange(0, 258)
    sizes = sizes[sizes <= 256]
    topic_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    for C in (3, 64, 256):
        b = dict(topic_off=topic_off, lag=rng.integers(0, 50, size=int(topic_off[-1]), dtype=np.int64),
                 partition=None, topic_cset=np.zeros(len(sizes), np.int32),
                 cset_off=np.array([0, C], np.int64), cset_rank=np.arange(C, dtype=np.int32) * 2 + 1)
        check(flat, b, mode=oracle.LITERAL)


@pytest.mark.parametrize("lag_kind", ["u40", "full", "dups", "zero"])
@pytest.mark.parametrize("ids", [None, "shuffled", "dups"])
def test_cta_bin(flat, lag_kind, ids):
    """256 < P <= 4096: one CTA per topic; warp-level and CTA-level round fill."""
    rng = np.random.default_rng(zlib.crc32(repr(("cta", lag_kind, ids)).encode()))
    for C in (1, 7, 200, 256, 257, 1000, 4000, 5000):
        b = random_flat(rng, T=12, pmin=257, pmax=4096, C=C, lag_kind=lag_kind, explicit_ids=ids,
                        n_sets=3, full_sets=(C >= 257))
        check(flat, b)


@pytest.mark.parametrize("lag_kind", ["u40", "full", "dups", "zero", "small"])
@pytest.mark.parametrize("ids", [None, "sorted", "shuffled", "dups"])
def test_device_bin(flat, lag_kind, ids):
    """P > 4096: segmented radix sort + finalize / round fill."""
    rng = np.random.default_rng(zlib.crc32(repr(("dev", lag_kind, ids)).encode()))
    for C, pmax in ((3, 9000), (300, 30000), (1024, 70000), (8192, 20000), (50000, 40000)):
        b = random_flat(rng, T=5, pmin=4097, pmax=pmax, C=C, lag_kind=lag_kind, explicit_ids=ids,
                        n_sets=2, full_sets=True)
        got = check(flat, b)
        st = flat.stats()
        assert st["topics_device"] == 5
        if lag_kind == "zero" and ids in (None, "sorted"):
            assert st["radix_passes_run"] == 0      # every digit constant: nothing moves
        if lag_kind == "small" and ids is None:
            assert st["radix_passes_run"] <= 3      # lags < 2^20: upper five digits constant
        assert got is not None


def test_mixed_bins_and_edge_topics(flat):
    """All bins in one batch, empty topics, empty consumer sets, repeated execute()."""
    rng = np.random.default_rng(5)
    sizes = np.array([0, 1, 2, 31, 32, 33, 255, 256, 257, 511, 4095, 4096, 4097, 8192, 8193, 12289, 0, 100, 5000])
    topic_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    N = int(topic_off[-1])
    sets = [np.zeros(0, np.int32), np.array([4], np.int32), np.arange(0, 40, 2, dtype=np.int32),
            np.arange(600, dtype=np.int32), np.arange(6000, dtype=np.int32)]
    cset_off = np.concatenate([[0], np.cumsum([len(s) for s in sets])]).astype(np.int64)
    b = dict(topic_off=topic_off, lag=rng.integers(-5, 1 << 33, size=N, dtype=np.int64),
             partition=rng.permutation(N).astype(np.int32),
             topic_cset=(np.arange(len(sizes)) % len(sets)).astype(np.int32),
             cset_off=cset_off, cset_rank=np.concatenate(sets).astype(np.int32))
    first = check(flat, b)
    # execute() twice on the same upload gives the same answer
    flat.upload(*args_of(b))
    flat.execute()
    flat.execute()
    again = flat.download()
    for k in KEYS:
        assert np.array_equal(first[k], again[k]), k
    st = flat.stats()
    assert st["topics_unassigned"] > 0 and st["topics_warp"] > 0 and st["topics_cta"] > 0 and st["topics_device"] > 0
    assert st["kernel_launches"] > 0


def test_chunked_pipeline_path(pkg):
    """Large batches of on-chip topics go through the chunked H2D / kernel / D2H pipeline
    of lagassign_assign_batch(); same answer as the oracle, pinned or pageable buffers."""
    rng = np.random.default_rng(77)
    b = random_flat(rng, T=30000, pmin=0, pmax=600, C=40, lag_kind="small", explicit_ids="shuffled", n_sets=4)
    assert len(b["lag"]) > (1 << 22)
    exp = oracle.assign_batch(*args_of(b), threads=8)
    with pkg.FlatAssignor(0) as f:
        got = f.assign_batch(*args_of(b))
        st = f.stats()
        assert st["topics_device"] == 0 and st["kernel_launches"] >= 4     # several chunks
        assert np.array_equal(got["out_partition"], exp["out_partition"])
        assert np.array_equal(got["out_consumer_rank"], exp["out_consumer_rank"])
        # pinned buffers, implicit ids
        b2 = dict(b, partition=None)
        exp2 = oracle.assign_batch(*args_of(b2), threads=8)
        lag = pkg.pinned_empty(len(b2["lag"]), np.int64)
        lag[:] = b2["lag"]
        op = pkg.pinned_empty(len(lag), np.int32)
        oc = pkg.pinned_empty(len(lag), np.int32)
        f.assign_batch(b2["topic_off"], lag, None, b2["topic_cset"], b2["cset_off"], b2["cset_rank"], op, oc)
        assert np.array_equal(op, exp2["out_partition"]) and np.array_equal(oc, exp2["out_consumer_rank"])
        # execute() again on what the pipeline left resident gives the same result
        f.execute()
        again = f.download()
        assert np.array_equal(again["out_partition"], exp2["out_partition"])


def test_plan_cache_same_shape_new_lags(pkg):
    """A batch with the shape of the previous one (same topics / consumer sets, fresh lags)
    reuses the plan and the metadata on the device; results follow the NEW data, and any change
    of shape or of an option that sizes outputs makes a new plan."""
    rng = np.random.default_rng(1234)
    b1 = random_flat(rng, T=300, pmin=0, pmax=4000, C=37, lag_kind="small", explicit_ids="shuffled", n_sets=3)
    b1["topic_off"] = np.concatenate([b1["topic_off"], [b1["topic_off"][-1] + 20000]]).astype(np.int64)   # + a radix-path topic
    b1["topic_cset"] = np.concatenate([b1["topic_cset"], [0]]).astype(np.int32)
    n = int(b1["topic_off"][-1])
    b1["lag"] = rng.integers(0, 50, size=n, dtype=np.int64)
    b1["partition"] = None
    with pkg.FlatAssignor(0, want_totals=True) as f:
        for step in range(3):
            b = dict(b1, lag=rng.integers(-5, 1 << (10 * step + 3), size=n, dtype=np.int64))
            got = f.assign_batch(*args_of(b))
            exp = oracle.assign_batch(*args_of(b), want_totals=True, threads=8)
            for k in KEYS:
                assert np.array_equal(got[k], exp[k]), (step, k)
            assert f.stats()["n_partitions"] == n and f.stats()["topics_device"] == 1
            if step > 0:
                assert f.stats()["h2d_bytes"] == n * 8        # nothing but the lags went up again
        # upload()/execute()/download() on the cached plan, with explicit ids this time (a new plan: ids are part of the shape flags)
        ids = np.concatenate([rng.permutation(int(hi - lo)) for lo, hi in zip(b1["topic_off"][:-1], b1["topic_off"][1:])]).astype(np.int32)
        b = dict(b1, lag=rng.integers(0, 3, size=n, dtype=np.int64), partition=ids)
        got = f.assign_batch(*args_of(b))
        exp = oracle.assign_batch(*args_of(b), want_totals=True, threads=8)
        for k in KEYS:
            assert np.array_equal(got[k], exp[k]), k
        # same shape, one consumer set changed in place: must not hit the cache
        b = dict(b, cset_rank=b["cset_rank"].copy())
        b["cset_rank"][-1] += 1
        got = f.assign_batch(*args_of(b))
        exp = oracle.assign_batch(*args_of(b), want_totals=True, threads=8)
        for k in KEYS:
            assert np.array_equal(got[k], exp[k]), k
        # fewer topics
        T2 = 100
        b = dict(b, topic_off=b["topic_off"][:T2 + 1], topic_cset=b["topic_cset"][:T2])
        m = int(b["topic_off"][-1])
        b["lag"], b["partition"] = b["lag"][:m], b["partition"][:m]
        got = f.assign_batch(*args_of(b))
        exp = oracle.assign_batch(*args_of(b), want_totals=True, threads=8)
        for k in KEYS:
            assert np.array_equal(got[k], exp[k]), k


def test_pipeline_chunk_option(pkg):
    """LAGASSIGN_OPT_PIPE_CHUNK only changes how assign_batch() overlaps its copies."""
    rng = np.random.default_rng(78)
    b = random_flat(rng, T=20000, pmin=1, pmax=500, C=64, lag_kind="small", explicit_ids=None, n_sets=2)
    assert len(b["lag"]) > (1 << 21)
    exp = oracle.assign_batch(*args_of(b), threads=8)
    with pkg.FlatAssignor(0) as f:
        for chunk in (0, 1 << 18, 1 << 20, 1 << 30):
            f.set_option(pkg.OPT_PIPE_CHUNK, chunk)
            got = f.assign_batch(*args_of(b))
            assert np.array_equal(got["out_partition"], exp["out_partition"]), chunk
            assert np.array_equal(got["out_consumer_rank"], exp["out_consumer_rank"]), chunk
        with pytest.raises(pkg.LagAssignError):
            f.set_option(pkg.OPT_PIPE_CHUNK, -1)


def test_single_consumer_and_single_partition(flat):
    rng = np.random.default_rng(9)
    for P, C in ((1, 1), (1, 5000), (257, 1), (5000, 1), (70000, 1), (70000, 2)):
        lag = rng.integers(0, 1000, size=P, dtype=np.int64)
        check(flat, uniform_batch(1, P, C, lag))


def test_wraparound_totals(flat):
    """Totals wrap like Java long (Main:265) and the wrapped value drives later picks."""
    big = (1 << 63) - 1
    for P, C in ((64, 3), (1000, 7), (6000, 300), (20000, 1024)):
        rng = np.random.default_rng(P)
        lag = rng.integers(big // 2, big, size=P, dtype=np.int64)
        check(flat, uniform_batch(1, P, C, lag))


def test_argument_errors(pkg):
    with pkg.FlatAssignor(0) as f:
        with pytest.raises(pkg.LagAssignError) as e:   # ranks of a set must ascend
            f.assign_batch([0, 2], [1, 2], None, [0], [0, 2], [3, 3])
        assert e.value.code == -1
        with pytest.raises(pkg.LagAssignError) as e:   # topic_cset out of range
            f.assign_batch([0, 2], [1, 2], None, [4], [0, 2], [0, 1])
        assert e.value.code == -1
        with pytest.raises(pkg.LagAssignError) as e:   # offsets not monotone
            f.assign_batch([0, 2, 1], [1], None, [0, 0], [0, 2], [0, 1])
        assert e.value.code == -1
        with pytest.raises(pkg.LagAssignError) as e:   # execute before upload
            f.execute()
        assert e.value.code == -4
        # (P > C > 8192 was LAGASSIGN_ERR_UNSUPPORTED in round 1; it is computed now: test_more_than_8192_consumers_*)


def test_fused_lag_computation(pkg):
    """upload_offsets(): computePartitionLag (Main:376-404) on the device, then assign."""
    rng = np.random.default_rng(21)
    b = random_flat(rng, T=40, pmin=0, pmax=6000, C=50, lag_kind="u40", n_sets=2)
    N = len(b["lag"])
    end = rng.integers(0, 1 << 40, size=N, dtype=np.int64)
    begin = (end * rng.random(N)).astype(np.int64)
    committed = rng.integers(0, 1 << 40, size=N, dtype=np.int64)   # sometimes beyond end -> clamp
    has = (rng.random(N) < 0.7).astype(np.uint8)
    for latest in (0, 1):
        exp_lag = np.array([oracle.compute_partition_lag(int(committed[i]) if has[i] else None, int(begin[i]),
                                                         int(end[i]), "latest" if latest else "earliest")
                            for i in range(N)], dtype=np.int64)
        with pkg.FlatAssignor(0) as f:
            f.upload_offsets(b["topic_off"], begin, end, committed, has, latest, None,
                             b["topic_cset"], b["cset_off"], b["cset_rank"])
            f.execute()
            got = f.download()
            assert np.array_equal(f.download_lags(), exp_lag)
        exp = oracle.assign_batch(b["topic_off"], exp_lag, None, b["topic_cset"], b["cset_off"], b["cset_rank"])
        assert np.array_equal(got["out_partition"], exp["out_partition"])
        assert np.array_equal(got["out_consumer_rank"], exp["out_consumer_rank"])


# ---- BASELINE.json configs --------------------------------------------------------------------

def test_config1_readme(flat):
    b, _ = baseline_config(1)
    got = check(flat, b, mode=oracle.LITERAL)
    assert got["out_partition"].tolist() == [0, 2, 1] and got["out_consumer_rank"].tolist() == [0, 1, 1]
    assert got["out_total"].tolist() == [100000, 110000]      # README.md:54-57


def test_config2_full(flat):
    """1 topic x 2^20 partitions, 1024 consumers: 1024 dependent rounds; totals wrap."""
    b, _ = baseline_config(2)
    got = check(flat, b)
    assert flat.stats()["radix_passes_run"] == 8
    assert (got["out_count"] == 1024).all()


def test_config3_full(flat):
    b, _ = baseline_config(3)
    check(flat, b)


def test_config4_scaled(flat):
    b, _ = baseline_config(4, scale=0.05)
    check(flat, b)


def test_config5_scaled_vs_oracle(flat):
    b, _ = baseline_config(5, scale=0.02)
    check(flat, b)


def test_config4_full_properties(pkg):
    """Full size: 100 000 topics x 256. Size-independent properties (P < C: one
    partial round): every topic's out_partition is a permutation of 0..255 whose
    lags are non-increasing with ties by partition id; sorted partition k goes to
    member rank k; plus a sampled comparison with the oracle."""
    b, _ = baseline_config(4)
    with pkg.FlatAssignor(0) as f:
        got = f.assign_batch(*args_of(b))
    T, P = 100000, 256
    part = got["out_partition"].reshape(T, P)
    cons = got["out_consumer_rank"].reshape(T, P)
    assert (np.sort(part, axis=1) == np.arange(P, dtype=np.int32)).all()
    assert (cons == np.arange(P, dtype=np.int32)).all()
    lag = b["lag"].reshape(T, P)
    sl = np.take_along_axis(lag, part.astype(np.int64), axis=1)
    d = np.diff(sl, axis=1)
    assert (d <= 0).all()
    assert (np.diff(part, axis=1)[d == 0] > 0).all()
    lo, hi = 5000, 5600
    exp = oracle.assign_batch(*args_of(b), topic_lo=lo, topic_hi=hi)
    assert np.array_equal(exp["out_partition"][lo * P:hi * P], got["out_partition"][lo * P:hi * P])
    assert np.array_equal(exp["out_consumer_rank"][lo * P:hi * P], got["out_consumer_rank"][lo * P:hi * P])


def test_config5_full_closed_form(pkg):
    """All-zero lag: the sort is the identity and partition i goes to member i mod C
     — a self-checking config at full size."""
    b, _ = baseline_config(5)
    with pkg.FlatAssignor(0) as f:
        got = f.assign_batch(*args_of(b))
        st = f.stats()
    off = b["topic_off"]
    N = int(off[-1])
    local = np.arange(N, dtype=np.int64) - np.repeat(off[:-1], np.diff(off))
    assert np.array_equal(got["out_partition"], local.astype(np.int32))
    assert np.array_equal(got["out_consumer_rank"], (local % 1024).astype(np.int32))
    assert st["radix_passes_run"] == 0 and st["topics_device"] > 0


def _all_threads():
    return max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))


@pytest.mark.parametrize("n,totals", [(4, False), (5, True), (6, False)])
def test_config_full_size_vs_oracle_every_topic(pkg, n, totals):
    """Full-size BASELINE configs 4 and 5 and the sort-only variant (config 6, 64 x 2^20 keys through all
    eight radix passes): EVERY topic compared bit-exactly with the oracle (round formulation on all host
    threads), not a sample. Totals are left out where they would be 100 000 x 4096 entries."""
    b, _ = baseline_config(n)
    with pkg.FlatAssignor(0, want_totals=totals) as f:
        got = f.assign_batch(*args_of(b))
    exp = oracle.assign_batch(*args_of(b), mode=oracle.ROUNDS, want_totals=totals, threads=_all_threads())
    for k in KEYS if totals else KEYS[:2]:
        assert np.array_equal(got[k], exp[k]), k


def test_more_than_8192_consumers_and_more_partitions(flat):
    """P > C > 8192: multi-round fill over more consumers than fit on chip (a global-memory workspace); the
    reference handles any shape, round 1 returned LAGASSIGN_ERR_UNSUPPORTED here. Mixed with ordinary topics."""
    rng = np.random.default_rng(92)
    sizes = np.array([40000, 300, 9000 * 3 + 5, 70000, 5000])
    topic_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    lag = rng.integers(0, 1 << 44, size=int(topic_off[-1]), dtype=np.int64)
    lag[int(topic_off[3]):] //= (1 << 30)                          # duplicates in the last two topics
    sets = [np.arange(9000, dtype=np.int32), np.sort(rng.choice(40000, size=20000, replace=False)).astype(np.int32),
            np.arange(100, dtype=np.int32)]
    cset_off = np.concatenate([[0], np.cumsum([len(x) for x in sets])]).astype(np.int64)
    b = dict(topic_off=topic_off, lag=lag, partition=None, topic_cset=np.array([0, 2, 0, 1, 1], np.int32),
             cset_off=cset_off, cset_rank=np.concatenate(sets).astype(np.int32))
    check(flat, b)


# ---- multi-device context (lagassign_create_multi) ---------------------------------------------------

def _multi_devices(pkg, want):
    n = pkg.device_count()
    return [i % n for i in range(want)]       # with fewer GPUs than ranges, devices repeat (the split is the same)


@pytest.mark.parametrize("n_dev", [2, 3, 8])
def test_multi_device_context_matches_oracle(pkg, n_dev):
    """One process, several devices: contiguous topic ranges, each device copying its results into the caller's
    arrays. Mixed bins, ragged sizes, empty topics, totals; compared bit-exactly with the oracle."""
    rng = np.random.default_rng(40 + n_dev)
    b = random_flat(rng, T=500, pmin=0, pmax=700, C=300, lag_kind="u40", explicit_ids="shuffled", n_sets=4)
    big = random_flat(rng, T=6, pmin=4000, pmax=30000, C=50, lag_kind="small", n_sets=1, full_sets=True)
    # append the large topics (their consumer set becomes set number 4)
    b["topic_off"] = np.concatenate([b["topic_off"], b["topic_off"][-1] + big["topic_off"][1:]])
    b["lag"] = np.concatenate([b["lag"], big["lag"]])
    b["partition"] = np.concatenate([b["partition"], np.concatenate([np.arange(int(p), dtype=np.int32) for p in np.diff(big["topic_off"])])])
    b["topic_cset"] = np.concatenate([b["topic_cset"], np.full(6, 4, np.int32)])
    b["cset_off"] = np.concatenate([b["cset_off"], [b["cset_off"][-1] + len(big["cset_rank"])]]).astype(np.int64)
    b["cset_rank"] = np.concatenate([b["cset_rank"], big["cset_rank"]]).astype(np.int32)
    with pkg.FlatAssignor(_multi_devices(pkg, n_dev), want_totals=True) as f:
        got = f.assign_batch(*args_of(b))
        st = f.stats()
        got2 = None
        f.upload(*args_of(b)); f.execute(); got2 = f.download()       # split form
    exp = oracle.assign_batch(*args_of(b), want_totals=True, threads=8)
    for k in KEYS:
        assert np.array_equal(got[k], exp[k]), k
        assert np.array_equal(got2[k], exp[k]), k
    assert st["n_partitions"] == len(b["lag"]) and st["topics_device"] == 6


def test_multi_device_context_config4_slice(pkg):
    """BASELINE config 4's shape (P < C, one consumer set) over 4 ranges, pipelined copies inside each."""
    b, _ = baseline_config(4, scale=0.2)
    with pkg.FlatAssignor(_multi_devices(pkg, 4)) as f:
        got = f.assign_batch(*args_of(b))
    exp = oracle.assign_batch(*args_of(b), threads=_all_threads())
    assert np.array_equal(got["out_partition"], exp["out_partition"])
    assert np.array_equal(got["out_consumer_rank"], exp["out_consumer_rank"])


def _check_narrow(got, b, exp):
    """The narrow form carries the same information as the int32 arrays (include/lagassign.h)."""
    off = np.asarray(b["topic_off"])
    assert np.array_equal(got["out_partition"].astype(np.int64), exp["out_partition"].astype(np.int64))
    if got["cons_width"] == 0:
        # every topic has P <= C or no subscriber: sorted partition k -> k-th member of the topic's set
        cs = np.asarray(b["topic_cset"]); coff = np.asarray(b["cset_off"]); cr = np.asarray(b["cset_rank"])
        for t in range(len(off) - 1):
            lo, hi = int(off[t]), int(off[t + 1])
            C = int(coff[cs[t] + 1] - coff[cs[t]])
            want = cr[coff[cs[t]]:coff[cs[t]] + hi - lo] if C else np.full(hi - lo, -1, np.int32)
            assert np.array_equal(exp["out_consumer_rank"][lo:hi], want), t
    else:
        c = got["out_consumer_rank"].astype(np.int64)
        c[c == (1 << (8 * got["cons_width"])) - 1] = -1
        assert np.array_equal(c, exp["out_consumer_rank"].astype(np.int64))


@pytest.mark.parametrize("case", ["config4", "multi_round_small", "wide_ids", "explicit_ids", "big_topics", "empty_sets"])
@pytest.mark.parametrize("n_dev", [1, 3])
def test_assign_batch_narrow(pkg, case, n_dev):
    """lagassign_assign_batch_narrow: 1 / 2 / 4-byte partition ids, 0 / 2-byte consumers, pipelined and plain paths,
    one device and a multi-device context."""
    rng = np.random.default_rng(99)
    if case == "config4":
        b, _ = baseline_config(4, scale=0.12); want = (1, 0)                    # pipelined path (3 M partitions)
    elif case == "multi_round_small":
        b = random_flat(rng, T=400, pmin=0, pmax=256, C=7, lag_kind="u40", n_sets=3); want = (1, 2)
    elif case == "wide_ids":
        b = random_flat(rng, T=60, pmin=0, pmax=3000, C=5000, lag_kind="small", n_sets=2, full_sets=True); want = (2, 0)
    elif case == "explicit_ids":
        b = random_flat(rng, T=50, pmin=0, pmax=700, C=20, lag_kind="dups", explicit_ids="shuffled", n_sets=2); want = (4, 2)
    elif case == "big_topics":
        b = random_flat(rng, T=3, pmin=66000, pmax=80000, C=300, lag_kind="full", n_sets=1, full_sets=True); want = (4, 2)
    else:
        b = random_flat(rng, T=30, pmin=1, pmax=200, C=300, lag_kind="small", n_sets=3, full_sets=True)
        b["cset_off"] = np.concatenate([b["cset_off"], b["cset_off"][-1:]])      # one more, empty, set ...
        b["topic_cset"] = b["topic_cset"].copy(); b["topic_cset"][::5] = len(b["cset_off"]) - 2   # ... for every fifth topic
        want = (1, 0)
    dev = 0 if n_dev == 1 else _multi_devices(pkg, n_dev)
    with pkg.FlatAssignor(dev) as f:
        got = f.assign_batch_narrow(*args_of(b))
        got2 = f.assign_batch_narrow(*args_of(b))          # plan-cache hit: same widths, same result
    assert (got["part_width"], got["cons_width"]) == want == (got2["part_width"], got2["cons_width"])
    exp = oracle.assign_batch(*args_of(b), threads=_all_threads())
    _check_narrow(got, b, exp)
    _check_narrow(got2, b, exp)


def test_multi_device_context_rejects_device_only_calls(pkg):
    with pkg.FlatAssignor(_multi_devices(pkg, 2)) as f:
        with pytest.raises(pkg.LagAssignError) as e:
            f.set_option(pkg.OPT_GROUP, 1)
        assert e.value.code == -6
        with pytest.raises(pkg.LagAssignError) as e:
            f.bind_device_outputs(1 << 20, 1 << 21)
        assert e.value.code == -6


# ---- result exchange of the one-process-per-GPU arrangement ------------------------------------------

def _exchange_roundtrip(pkg, b, world):
    """pack / unpack on ONE device, standing in for `world` ranks: every shard's results (the oracle's) are narrowed
    into its slot, the slots are widened; must reproduce the oracle's arrays for the whole batch."""
    import ctypes
    import torch
    from kafka_lag_based_assignor_b200 import distributed as D
    L = pkg.load_library()
    exp = oracle.assign_batch(*args_of(b), threads=8)
    bounds = D.shard_bounds(b["topic_off"], world)
    counts = D.partition_counts(b["topic_off"], bounds)
    pw, cw = D.exchange_format(b)
    slot = int(L.lagassign_exchange_slot_bytes(max(counts), pw, cw))
    dev = torch.device("cuda", 0)
    vp = ctypes.c_void_p
    st = vp(torch.cuda.current_stream().cuda_stream)
    slots = torch.full((world * max(slot, 16),), 0xAB, dtype=torch.uint8, device=dev)
    start = 0
    for r, n in enumerate(counts):
        dp = torch.from_numpy(exp["out_partition"][start:start + n].copy()).to(dev) if n else torch.zeros(1, dtype=torch.int32, device=dev)
        dc = torch.from_numpy(exp["out_consumer_rank"][start:start + n].copy()).to(dev) if n else torch.zeros(1, dtype=torch.int32, device=dev)
        assert L.lagassign_exchange_pack(st, n, vp(dp.data_ptr()), vp(dc.data_ptr()), pw, cw, vp(slots.data_ptr() + r * slot)) == 0
        start += n
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
    T = len(b["topic_off"]) - 1
    d_off, d_tc, d_co = t(b["topic_off"], np.int64), t(b["topic_cset"], np.int32), t(b["cset_off"], np.int64)
    d_cr = t(b["cset_rank"], np.int32) if len(b["cset_rank"]) else torch.zeros(1, dtype=torch.int32, device=dev)
    d_first = t([lo for lo, _ in bounds] + [T], np.int64)
    N = int(b["topic_off"][-1])
    op = torch.full((max(N, 1),), -9, dtype=torch.int32, device=dev)
    oc = torch.full((max(N, 1),), -9, dtype=torch.int32, device=dev)
    assert L.lagassign_exchange_unpack(st, T, vp(d_off.data_ptr()), vp(d_tc.data_ptr()), vp(d_co.data_ptr()), vp(d_cr.data_ptr()),
                                       vp(d_first.data_ptr()), world, vp(slots.data_ptr()), slot, pw, cw,
                                       vp(op.data_ptr()), vp(oc.data_ptr())) == 0
    torch.cuda.synchronize()
    assert np.array_equal(op[:N].cpu().numpy(), exp["out_partition"]), (pw, cw)
    assert np.array_equal(oc[:N].cpu().numpy(), exp["out_consumer_rank"]), (pw, cw)
    return pw, cw


def test_result_exchange_pack_unpack(pkg):
    rng = np.random.default_rng(61)
    seen = set()
    b4, _ = baseline_config(4, scale=0.01)
    seen.add(_exchange_roundtrip(pkg, b4, 8))                                                      # (1, 0): consumers derived
    seen.add(_exchange_roundtrip(pkg, random_flat(rng, T=200, pmin=0, pmax=200, C=9, lag_kind="dups", n_sets=3), 3))   # (1, 2), some unsubscribed topics
    seen.add(_exchange_roundtrip(pkg, random_flat(rng, T=40, pmin=0, pmax=3000, C=5000, lag_kind="u40", n_sets=2, full_sets=True), 2))  # (2, 0)
    seen.add(_exchange_roundtrip(pkg, random_flat(rng, T=30, pmin=0, pmax=2000, C=50, lag_kind="small", explicit_ids="shuffled", n_sets=2), 5))  # (2, 2)
    wide = random_flat(rng, T=3, pmin=66000, pmax=70000, C=70000, lag_kind="zero", n_sets=1, full_sets=True)
    wide["cset_rank"] = (wide["cset_rank"].astype(np.int64) + 100000).astype(np.int32)             # ranks beyond 16 bits
    seen.add(_exchange_roundtrip(pkg, wide, 2))                                                    # (4, 0)
    mixed = random_flat(rng, T=4, pmin=66000, pmax=70000, C=300, lag_kind="zero", n_sets=1, full_sets=True)
    mixed["cset_rank"] = (mixed["cset_rank"].astype(np.int64) * 300).astype(np.int32)
    seen.add(_exchange_roundtrip(pkg, mixed, 3))                                                   # (4, 4)
    assert {(1, 0), (1, 2), (2, 0), (2, 2), (4, 0), (4, 4)} <= seen, seen


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_nccl_gather_matches_oracle(pkg, world):
    """torchrun, one process per GPU: kernels -> narrow -> NCCL all-gather -> widen; every rank compares BOTH
    gathered arrays of the whole batch with the oracle (tests/dist_gpu_worker.py). Needs `world` GPUs."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(os.path.dirname(os.path.abspath(__file__)), "dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("-> ok") == 4 * world, r.stdout


# ---- remaining C-ABI entry points -------------------------------------------------------------------

def test_profile_option_and_stats(pkg):
    """LAGASSIGN_OPT_PROFILE: per-kernel CUDA-event times of the last execute()."""
    b, _ = baseline_config(5, scale=0.02)
    with pkg.FlatAssignor(0) as f:
        f.set_option(pkg.OPT_PROFILE, 1)
        f.upload(*args_of(b))
        f.execute()
        prof = f.profile()
        st = f.stats()
        assert "k_warp_topics" in prof and all(ms >= 0 and n >= 1 for ms, n in prof.values())
        assert 0 < sum(n for _, n in prof.values()) <= st["kernel_launches"]   # the two scan kernels share one scope
        f.set_option(pkg.OPT_PROFILE, 0)
        f.execute()
        assert f.profile() == {}
        got = f.download()
    exp = oracle.assign_batch(*args_of(b))
    assert np.array_equal(got["out_partition"], exp["out_partition"])
    assert np.array_equal(got["out_consumer_rank"], exp["out_consumer_rank"])


def test_bind_device_outputs(pkg):
    """execute() can write straight into caller-owned device memory (what the all-gather uses)."""
    import torch
    rng = np.random.default_rng(3)
    b = random_flat(rng, T=50, pmin=0, pmax=9000, C=30, lag_kind="u40", n_sets=2)
    N = len(b["lag"])
    dp = torch.full((N + 8,), -7, dtype=torch.int32, device="cuda")
    dc = torch.full((N + 8,), -7, dtype=torch.int32, device="cuda")
    with pkg.FlatAssignor(0) as f:
        f.upload(*args_of(b))
        f.bind_device_outputs(dp.data_ptr() + 16, dc.data_ptr() + 16)
        f.execute()
        f.sync()
        got = f.download()          # download follows the binding
        f.bind_device_outputs(None, None)
    exp = oracle.assign_batch(*args_of(b))
    assert np.array_equal(dp[4:4 + N].cpu().numpy(), exp["out_partition"])
    assert np.array_equal(dc[4:4 + N].cpu().numpy(), exp["out_consumer_rank"])
    assert (dp[:4] == -7).all() and (dp[4 + N:] == -7).all()
    assert np.array_equal(got["out_partition"], exp["out_partition"])


def test_realistic_many_rounds(flat):
    """The common real-world shape: more partitions than consumers (several rounds per topic),
    small lags with zero tails and duplicates, a few dozen consumers."""
    rng = np.random.default_rng(99)
    for C, pmax in ((3, 40), (16, 200), (40, 256), (200, 256), (300, 3000), (700, 9000)):
        b = random_flat(rng, T=300 if pmax <= 256 else 12, pmin=1, pmax=pmax, C=C, lag_kind="small",
                        n_sets=3, full_sets=True)
        z = rng.random(len(b["lag"])) < 0.5          # half of the partitions are caught up
        b["lag"] = np.where(z, 0, b["lag"]).astype(np.int64)
        check(flat, b)


def test_random_shape_sweep(flat):
    """Many random (P, C) shapes in ONE batch per seed: every bin, every fill path, random
    consumer-set sharing, mixed lag distributions. Bit-exact against the oracle."""
    for seed in range(6):
        rng = np.random.default_rng(1000 + seed)
        T = 400
        # log-uniform partition counts from 1 to ~20000, a few empty topics
        P = np.floor(2.0 ** (rng.random(T) * 14.3)).astype(np.int64)
        P[rng.random(T) < 0.03] = 0
        topic_off = np.concatenate([[0], np.cumsum(P)]).astype(np.int64)
        N = int(topic_off[-1])
        kinds = rng.integers(0, 4, size=T)
        lag = np.empty(N, np.int64)
        for t in range(T):
            lo, hi = int(topic_off[t]), int(topic_off[t + 1])
            n = hi - lo
            if kinds[t] == 0:
                lag[lo:hi] = rng.integers(0, 1 << 20, size=n)
            elif kinds[t] == 1:
                lag[lo:hi] = rng.integers(0, 3, size=n)
            elif kinds[t] == 2:
                lag[lo:hi] = rng.integers(-(1 << 63), (1 << 63) - 1, size=n, dtype=np.int64, endpoint=True)
            else:
                v = rng.integers(0, 1 << 45, size=n)
                v[rng.random(n) < 0.6] = 0
                lag[lo:hi] = v
        sizes = [0, 1, 2, 3, 17, 32, 33, 100, 256, 257, 600, 1024, 2000, 5000]
        M = 6000
        sets = [np.sort(rng.choice(M, size=c, replace=False)).astype(np.int32) for c in sizes]
        cset_off = np.concatenate([[0], np.cumsum([len(x) for x in sets])]).astype(np.int64)
        b = dict(topic_off=topic_off, lag=lag, partition=None if seed % 2 == 0 else
                 np.concatenate([rng.permutation(int(p)) for p in P] + [np.zeros(0, np.int64)]).astype(np.int32),
                 topic_cset=rng.integers(0, len(sets), size=T).astype(np.int32), cset_off=cset_off,
                 cset_rank=np.concatenate(sets).astype(np.int32))
        check(flat, b)


def _expected_grouping(b, res):
    """Reference list order (Main:264): per member, topics in batch order, then assignment order
    = a stable partition of the flat result by member rank."""
    off = b["topic_off"]
    topic_of = np.repeat(np.arange(len(off) - 1, dtype=np.int32), np.diff(off))
    cons = res["out_consumer_rank"]
    keep = np.nonzero(cons >= 0)[0]
    order = keep[np.argsort(cons[keep], kind="stable")]
    M = int(b["cset_rank"].max()) + 1 if len(b["cset_rank"]) else 0
    counts = np.bincount(cons[keep], minlength=M)
    member_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return member_off, topic_of[order], res["out_partition"][order]


@pytest.mark.parametrize("shape", ["small", "mixed", "config4"])
def test_grouping_by_member(pkg, shape):
    """LAGASSIGN_OPT_GROUP: the reference's Map<member, List<TopicPartition>> in CSR form,
    built on the device."""
    rng = np.random.default_rng(5)
    if shape == "small":
        b = random_flat(rng, T=7, pmin=0, pmax=30, C=5, lag_kind="dups", n_sets=3)
    elif shape == "mixed":
        b = random_flat(rng, T=300, pmin=0, pmax=6000, C=700, lag_kind="u40", explicit_ids="shuffled", n_sets=5)
    else:
        b, _ = baseline_config(4, scale=0.1)
    with pkg.FlatAssignor(0) as f:
        f.set_option(pkg.OPT_GROUP, 1)
        f.upload(*args_of(b))
        f.execute()
        res = f.download()
        g = f.download_grouping()
        f.execute()                      # repeatable
        g2 = f.download_grouping()
    exp = oracle.assign_batch(*args_of(b), threads=8)
    assert np.array_equal(res["out_partition"], exp["out_partition"])
    assert np.array_equal(res["out_consumer_rank"], exp["out_consumer_rank"])
    e_off, e_topic, e_part = _expected_grouping(b, exp)
    assert np.array_equal(g["member_off"], e_off)
    assert np.array_equal(g["topic"], e_topic)
    assert np.array_equal(g["partition"], e_part)
    for k in g:
        assert np.array_equal(g[k], g2[k])


def test_grouping_reference_vector(pkg):
    """Test:112-125 of the reference, read from the grouped arrays: consumer-1 =
    [topic1-0, topic1-2, topic2-0, topic2-1], consumer-2 = [topic1-1, topic1-3]."""
    topic_off = np.array([0, 4, 6], np.int64)           # topic1, topic2 in HashMap order
    lag = np.array([100000, 100000, 500, 1, 900000, 100000], np.int64)
    with pkg.FlatAssignor(0) as f:
        f.set_option(pkg.OPT_GROUP, 1)
        f.upload(topic_off, lag, None, [0, 1], [0, 2, 3], [0, 1, 0])
        f.execute()
        g = f.download_grouping()
    assert g["member_off"].tolist() == [0, 4, 6]
    assert list(zip(g["topic"].tolist(), g["partition"].tolist())) == [(0, 0), (0, 2), (1, 0), (1, 1), (0, 1), (0, 3)]


# ---- consumersPerTopic + consumer-set de-duplication on the device --------------------------------

def _invert_reference(n_topics, sub_off, sub_topic):
    """Plain Python restatement of Main:410-426 on integers, then the de-duplication rule of include/lagassign.h:
    per topic the ascending set of members, identical sets stored once, numbered by first appearance."""
    per_topic = [set() for _ in range(n_topics)]
    for m in range(len(sub_off) - 1):
        for e in range(int(sub_off[m]), int(sub_off[m + 1])):
            per_topic[int(sub_topic[e])].add(m)
    ids, sets, topic_cset = {}, [], []
    for t in range(n_topics):
        key = tuple(sorted(per_topic[t]))
        if key not in ids:
            ids[key] = len(sets)
            sets.append(key)
        topic_cset.append(ids[key])
    cset_off = np.concatenate([[0], np.cumsum([len(s) for s in sets])]).astype(np.int64)
    cset_rank = np.array([m for s in sets for m in s], dtype=np.int32)
    return np.array(topic_cset, np.int32), cset_off, cset_rank


@pytest.mark.parametrize("case", ["few_sets", "all_distinct", "one_set", "empty_topics", "repeats", "wide"])
def test_invert_subscriptions_on_device(pkg, case):
    import zlib
    rng = np.random.default_rng(zlib.crc32(case.encode()))
    if case == "few_sets":        # a handful of subscription patterns shared by many members' topics
        M, T = 300, 5000
        patterns = [np.sort(rng.choice(T, size=int(rng.integers(1, T)), replace=False)) for _ in range(4)]
        subs = [patterns[int(rng.integers(0, 4))] for _ in range(M)]
    elif case == "all_distinct":
        M, T = 70, 900
        subs = [np.sort(rng.choice(T, size=int(rng.integers(0, T)), replace=False)) for _ in range(M)]
    elif case == "one_set":       # BASELINE config 4's shape: everybody subscribes to everything
        M, T = 4096, 3000
        subs = [np.arange(T)] * M
    elif case == "empty_topics":
        M, T = 40, 1000
        subs = [np.sort(rng.choice(T // 2, size=int(rng.integers(0, 50)), replace=False)) for _ in range(M)]
    elif case == "repeats":       # a topic listed twice by one member counts once; order inside a list is free
        M, T = 33, 64
        subs = [rng.integers(0, T, size=int(rng.integers(0, 200))) for _ in range(M)]
    else:                         # more than one 32-member word per row, not a multiple of 32
        M, T = 1000 + 7, 257
        subs = [np.sort(rng.choice(T, size=int(rng.integers(0, T)), replace=False)) for _ in range(M)]
    sub_off = np.concatenate([[0], np.cumsum([len(s) for s in subs])]).astype(np.int64)
    sub_topic = np.concatenate(list(subs) + [np.zeros(0, np.int64)]).astype(np.int32)
    got = pkg.invert_subscriptions(T, sub_off, sub_topic)
    tc, co, cr = _invert_reference(T, sub_off, sub_topic)
    assert np.array_equal(got["topic_cset"], tc)
    assert np.array_equal(got["cset_off"], co)
    assert np.array_equal(got["cset_rank"], cr)


def test_invert_subscriptions_matches_host_flatten(pkg):
    """The same consumer sets as LagBasedPartitionAssignor::flatten (host, strings) for a generated group."""
    rng = np.random.default_rng(5)
    members = [f"m{int(i)}" for i in rng.permutation(60)]
    topics = [f"topic-{i}" for i in range(200)]
    subs = {m: [topics[int(i)] for i in rng.permutation(200)[:int(rng.integers(1, 200))]] for m in members}
    lags = {t: [(t, p, int(rng.integers(0, 1000))) for p in range(3)] for t in topics}
    a = pkg.LagBasedPartitionAssignor()
    flat, topic_order, by_rank = a.flatten(lags, subs)
    rank = {m: r for r, m in enumerate(by_rank)}
    tindex = {t: i for i, t in enumerate(topic_order)}
    lists = [[] for _ in by_rank]
    for m, ts in subs.items():
        lists[rank[m]] = [tindex[t] for t in ts if t in tindex]
    sub_off = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int64)
    sub_topic = np.array([t for x in lists for t in x], np.int32)
    got = pkg.invert_subscriptions(len(topic_order), sub_off, sub_topic)
    assert np.array_equal(got["topic_cset"], flat["topic_cset"])
    assert np.array_equal(got["cset_off"], flat["cset_off"])
    assert np.array_equal(got["cset_rank"], flat["cset_rank"])


def test_invert_subscriptions_argument_errors(pkg):
    with pytest.raises(pkg.LagAssignError) as e:
        pkg.invert_subscriptions(4, [0, 2], [1, 9])            # topic index out of range
    assert e.value.code == -1


def test_host_mirror_assign_from_offsets(pkg):
    """lagassign_group_add_topic_offsets: the plugin-shaped mirror fed with raw offsets (what readTopicPartitionLags
    gathers, Main:338-356) gives the assignment of the same group fed with computePartitionLag's lags (Main:376-404)."""
    rng = np.random.default_rng(31)
    members = [f"consumer-{int(i)}" for i in rng.permutation(9)]
    topics = [f"orders.{i}" for i in range(25)]
    subs = {m: [topics[int(i)] for i in rng.permutation(25)[:int(rng.integers(1, 25))]] for m in members}
    for mode in ("latest", "earliest", "LATEST", "none"):
        offsets, lags = {}, {}
        for t in topics:
            P = int(rng.integers(0, 40))
            ids = [int(x) for x in (range(P) if rng.random() < 0.5 else rng.permutation(P + 5)[:P])]
            rows, lrows = [], []
            for pid in ids:
                end = int(rng.integers(0, 1 << 40))
                begin = int(end * rng.random())
                com = None if rng.random() < 0.3 else int(rng.integers(0, 1 << 40))
                rows.append((pid, begin, end, com))
                lrows.append((t, pid, oracle.compute_partition_lag(com, begin, end, mode)))
            offsets[t], lags[t] = rows, lrows
        a = pkg.LagBasedPartitionAssignor(0)
        got = a.assign_from_offsets(offsets, subs, auto_offset_reset=mode)
        exp = a.assign(lags, subs)
        assert got == exp
        a.close()
