"""The host side of assign() without a GPU: the C++ mirror's flatten / unflatten
(lagassign_group_flatten, lagassign_group_unflatten; reference steps Main:171-176, :258-259,
:410-426 before the device call and Main:264 after it), driven with the ORACLE standing in
for the device step. The product path never does that: lagassign_group_assign calls the CUDA
library between the same two functions (checked on the GPU by tests/test_gpu_parity.py).
"""
import json
import os

import numpy as np
import pytest

import oracle
from oracle import flat as oflat
from oracle import oracle_py
from helpers import args_of

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__
    __graft_entry__.build()
    import kafka_lag_based_assignor_b200 as p
    return p


def _through_host(pkg, lags, subs, want_flat=False):
    """flatten (C++) -> oracle on the flat arrays -> unflatten (C++)."""
    a = pkg.LagBasedPartitionAssignor()
    flat, topics, by_rank = a.flatten({t: [tuple(e) for e in v] for t, v in lags.items()}, dict(subs))
    out = oracle.assign_batch(*args_of(flat), mode=oracle.LITERAL)
    res = a.unflatten(out["out_partition"], out["out_consumer_rank"])
    res = {m: [tuple(tp) for tp in v] for m, v in res.items()}
    return (res, flat, topics, by_rank) if want_flat else res


def _names(rng, n, alphabet):
    out = set()
    while len(out) < n:
        k = int(rng.integers(1, 7))
        out.add("".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), size=k)))
    return sorted(out, key=lambda s: rng.random())


def _random_case(seed):
    rng = np.random.default_rng(seed)
    # ids chosen so that String.compareTo differs from numeric / UTF-8 byte / code-point order:
    # "C10" < "C2", BMP private-use U+E000 > any surrogate pair's first unit U+D83D
    member_alpha = list("C0129abZ_-") + ["é", "", "\U0001F600"]
    members = _names(rng, int(rng.integers(1, 14)), member_alpha)
    topics = _names(rng, int(rng.integers(1, 40)), list("tT.-_019") + ["ü"])
    subs = []
    for m in members:
        k = int(rng.integers(0, len(topics) + 1))
        mine = [topics[int(i)] for i in rng.permutation(len(topics))[:k]]
        if mine and rng.random() < 0.3:
            mine.append(mine[0])                      # a topic listed twice by one member
        subs.append((m, mine))
    lags = {}
    for t in topics + ["never-subscribed"]:
        if rng.random() < 0.15:
            continue                                  # subscribed topic without metadata (Main:182)
        P = int(rng.integers(0, 12))
        ids = list(range(P)) if rng.random() < 0.6 else [int(x) for x in rng.permutation(P + 3)[:P]]
        lg = rng.integers(0, 5, size=P) if rng.random() < 0.5 else rng.integers(-3, 1 << 40, size=P)
        lags[t] = [(t, ids[i], int(lg[i])) for i in range(P)]
    return lags, subs


def test_reference_vectors_through_flatten_and_unflatten(pkg):
    kats = json.load(open(GOLDEN))["assign"]
    assert kats
    for k in kats:
        lags = {t: [tuple(e) for e in v] for t, v in k["partition_lag_per_topic"].items()}
        subs = [(m, list(ts)) for m, ts in k["subscriptions"]]
        got = _through_host(pkg, lags, subs)
        exp = {m: [tuple(tp) for tp in v] for m, v in k["expected"].items()}
        if k["order_sensitive"]:
            assert got == exp, k["name"]
        else:
            assert {m: sorted(v) for m, v in got.items()} == {m: sorted(v) for m, v in exp.items()}, k["name"]


@pytest.mark.parametrize("seed", range(40))
def test_flatten_matches_the_independent_python_flatten(pkg, seed):
    lags, subs = _random_case(seed)
    _, flat, topics, by_rank = _through_host(pkg, lags, subs, want_flat=True)
    oflat_, otopics, omembers = oflat.flatten({t: list(v) for t, v in lags.items()}, subs)
    assert topics == otopics                          # HashMap iteration order (Main:177)
    assert by_rank == omembers                        # String.compareTo ranks (Main:258-259)
    for key in ("topic_off", "lag", "topic_cset", "cset_off", "cset_rank"):
        assert np.array_equal(flat[key], oflat_[key]), key
    if flat["partition"] is None:                     # ids 0..P-1 in order everywhere: NULL in the ABI
        off = oflat_["topic_off"]
        exp = np.concatenate([np.arange(off[i + 1] - off[i]) for i in range(len(off) - 1)] + [np.zeros(0, np.int64)])
        assert np.array_equal(oflat_["partition"], exp.astype(np.int32))
    else:
        assert np.array_equal(flat["partition"], oflat_["partition"])


@pytest.mark.parametrize("seed", range(40, 70))
def test_map_level_result_equals_the_literal_python_restatement(pkg, seed):
    """flatten -> (oracle on flat arrays) -> unflatten == Main:166-277 restated on maps."""
    lags, subs = _random_case(seed)
    got = _through_host(pkg, lags, subs)
    exp = oracle_py.assign({t: list(v) for t, v in lags.items()}, subs)
    assert list(got.keys()) == [m for m, _ in subs]   # every member has a list (Main:171-174)
    assert got == {m: [tuple(tp) for tp in v] for m, v in exp.items()}


def test_consumer_sets_are_deduplicated_and_ranked(pkg):
    """100 topics, 3 distinct subscriber sets -> 3 consumer sets in the flat batch; members in
    String.compareTo order ("C10" < "C2")."""
    members = ["C2", "C10", "C1", "c1"]
    topics = [f"t{i:03d}" for i in range(100)]
    subs = [("C2", topics), ("C10", topics[:50]), ("C1", topics[25:]), ("c1", [])]
    lags = {t: [(t, p, 10 * p + i) for p in range(3)] for i, t in enumerate(topics)}
    a = pkg.LagBasedPartitionAssignor()
    flat, order, by_rank = a.flatten(lags, dict(subs))
    assert by_rank == ["C1", "C10", "C2", "c1"]
    assert len(flat["cset_off"]) - 1 == 3
    assert flat["partition"] is None
    sets = {tuple(flat["cset_rank"][flat["cset_off"][s]:flat["cset_off"][s + 1]]) for s in range(3)}
    assert sets == {(1, 2), (0, 1, 2), (0, 2)}
    assert sorted(order) == topics


def test_unflatten_argument_errors(pkg):
    a = pkg.LagBasedPartitionAssignor()
    with pytest.raises(pkg.LagAssignError):           # before any flatten
        a._flat_members = []
        a.unflatten(np.zeros(1, np.int32), np.zeros(1, np.int32))
    a.flatten({"t": [("t", 0, 5), ("t", 1, 6)]}, {"m": ["t"]})
    with pytest.raises(pkg.LagAssignError):           # a rank that names no member
        a.unflatten(np.array([0, 1], np.int32), np.array([0, 7], np.int32))
    assert a.unflatten(np.array([1, 0], np.int32), np.array([0, -1], np.int32)) == {"m": [("t", 1)]}


# ---- topic order = iteration order of the java.util.HashMap built by consumersPerTopic, tree bins included ---------

def _topic_order_cpp(pkg, topics):
    """Topics in the order the C++ mirror visits them (JavaHashMapOrder) when one member subscribes to `topics`."""
    a = pkg.LagBasedPartitionAssignor()
    lags = {t: [(t, 0, 1)] for t in topics}
    _, order, _ = a.flatten(lags, {"m": list(topics)})
    return order


def _colliding_names(rng, n, low_bits):
    """n distinct names whose spread String.hashCode agrees in its `low_bits` lowest bits (they share a bucket until the
    table has more than 2^low_bits slots, then split both ways)."""
    from java_hashmap_model import string_hash
    out, target = [], None
    i = 0
    while len(out) < n:
        s = "topic-%d-%d" % (int(rng.integers(0, 1 << 30)), i)
        i += 1
        h = string_hash(s)
        h = (h ^ (h >> 16)) & ((1 << low_bits) - 1)
        if target is None:
            target = h
        if h == target:
            out.append(s)
    return out


@pytest.mark.parametrize("case", ["plain", "equal_hashes", "shared_low_bits", "two_tree_bins", "big"])
def test_topic_order_matches_the_python_hashmap_model(pkg, case):
    import itertools
    from java_hashmap_model import JavaHashMapModel
    rng = np.random.default_rng({"plain": 1, "equal_hashes": 2, "shared_low_bits": 3, "two_tree_bins": 4, "big": 5}[case])
    filler = ["f%d" % i for i in range(200)]
    if case == "plain":
        topics = ["t%05d" % int(i) for i in rng.permutation(3000)]      # (such regular names do fill a bin now and then)
        expect_tree = None
    elif case == "equal_hashes":
        # "Aa" and "BB" have the same hashCode, so do all concatenations: one bucket for ever; 9th key -> resize (table
        # below 64 slots), ... then a tree; later keys are linked after their tree parents
        fam = ["".join(p) for p in itertools.product(["Aa", "BB"], repeat=5)]
        order = [int(i) for i in rng.permutation(len(fam))]
        topics = [fam[i] for i in order[:12]] + filler[:70] + [fam[i] for i in order[12:]] + filler[70:]
        expect_tree = True
    elif case == "shared_low_bits":
        col = _colliding_names(rng, 40, 9)         # share a bucket up to 512 slots; the 1024-slot table splits the tree
        topics = filler[:50] + col[:25] + ["g%d" % i for i in range(600)] + col[25:] + ["h%d" % i for i in range(400)]
        expect_tree = True
    elif case == "two_tree_bins":
        a, b = _colliding_names(rng, 14, 8), _colliding_names(rng, 11, 8)
        topics = filler[:60]
        for i in range(14):
            topics.append(a[i])
            if i < 11:
                topics.append(b[i])
        topics += ["z%d" % i for i in range(300)]   # growth past 256 slots: both trees split, short halves untreeify
        expect_tree = True
    else:
        topics = ["events.v1.tenant%d.stream" % int(i) for i in rng.permutation(60000)]
        expect_tree = None
    model = JavaHashMapModel()
    for t in topics:
        model.put(t)
    if expect_tree is not None:
        assert model.treeified == expect_tree
    got = _topic_order_cpp(pkg, topics)
    assert got == model.order()
    assert sorted(got) == sorted(set(topics))


def test_topic_order_without_tree_bins_is_the_plain_chain_order(pkg):
    """Where no bin is treeified the order is: buckets by index, keys of a bucket in insertion order (round 1's model)."""
    from java_hashmap_model import string_hash
    rng = np.random.default_rng(9)
    topics = ["n%d" % int(i) for i in rng.permutation(500)]
    cap = 16
    while len(topics) > cap * 3 // 4:
        cap *= 2
    buckets = {}
    for t in topics:
        h = string_hash(t)
        buckets.setdefault((h ^ (h >> 16)) & (cap - 1), []).append(t)
    expect = [t for i in sorted(buckets) for t in buckets[i]]
    assert _topic_order_cpp(pkg, topics) == expect
