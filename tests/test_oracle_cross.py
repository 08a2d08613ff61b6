"""Cross-checks between the three restatements (oracle_py literal, C literal, C
rounds) on generated inputs, including int64 wrap-around, negative lags, P < C,
unordered / duplicated partition ids and member ids whose String.compareTo order
differs from arrival order. CPU only."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle
from oracle import flat as oflat
from oracle import oracle_py

I64 = st.integers(min_value=-(1 << 63), max_value=(1 << 63) - 1)
SMALL = st.integers(min_value=0, max_value=6)
LAG = st.one_of(SMALL, st.integers(min_value=0, max_value=10**6), I64,
                st.sampled_from([(1 << 63) - 1, -(1 << 63), (1 << 62), -(1 << 62)]))


@st.composite
def groups(draw):
    n_members = draw(st.integers(1, 7))
    n_topics = draw(st.integers(1, 4))
    members = draw(st.lists(st.text(alphabet="abC0129-é\U0001F600", min_size=1, max_size=4),
                            min_size=n_members, max_size=n_members, unique=True))
    topics = [f"topic{i}" for i in range(n_topics)]
    subs = []
    for m in members:
        ts = draw(st.lists(st.sampled_from(topics), max_size=n_topics, unique=True))
        subs.append((m, ts))
    lags = {}
    for t in topics:
        if draw(st.booleans()) or True:
            p = draw(st.integers(0, 12))
            ids = draw(st.one_of(st.just(list(range(p))),
                                 st.lists(st.integers(0, 15), min_size=p, max_size=p)))
            lags[t] = [(t, ids[i], draw(LAG)) for i in range(p)]
    return lags, subs


@settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(groups())
def test_three_oracles_agree(g):
    lags, subs = g
    exp = oracle_py.assign({t: list(v) for t, v in lags.items()}, subs)
    fl, topic_order, members = oflat.flatten(lags, subs)
    outs = []
    for mode in (oracle.LITERAL, oracle.ROUNDS):
        r = oracle.assign_batch(fl["topic_off"], fl["lag"], fl["partition"], fl["topic_cset"],
                                fl["cset_off"], fl["cset_rank"], mode=mode, want_totals=True)
        got = oflat.unflatten(fl, topic_order, members, r["out_partition"], r["out_consumer_rank"],
                              [m for m, _ in subs])
        assert got == exp
        outs.append(r)
    for k in ("out_partition", "out_consumer_rank", "out_total", "out_count"):
        assert np.array_equal(outs[0][k], outs[1][k]), k


def _random_flat(rng, T, pmax, C, lag_kind, explicit_ids=False, n_sets=1):
    P = rng.integers(0, pmax + 1, size=T)
    topic_off = np.concatenate([[0], np.cumsum(P)]).astype(np.int64)
    N = int(topic_off[-1])
    if lag_kind == "full":
        lag = rng.integers(-(1 << 63), (1 << 63) - 1, size=N, dtype=np.int64, endpoint=True)
    elif lag_kind == "dups":
        lag = rng.integers(0, 4, size=N, dtype=np.int64)
    else:
        lag = rng.integers(0, 1 << 40, size=N, dtype=np.int64)
    part = None
    if explicit_ids:
        part = np.concatenate([rng.permutation(int(p)) for p in P] + [np.zeros(0, np.int64)]).astype(np.int32)
    sets = []
    for _ in range(n_sets):
        c = int(rng.integers(0, C + 1))
        sets.append(np.sort(rng.choice(C, size=c, replace=False)).astype(np.int32))
    cset_off = np.concatenate([[0], np.cumsum([len(s) for s in sets])]).astype(np.int64)
    cset_rank = np.concatenate(sets + [np.zeros(0, np.int32)]).astype(np.int32)
    topic_cset = rng.integers(0, n_sets, size=T).astype(np.int32)
    return topic_off, lag, part, topic_cset, cset_off, cset_rank


@pytest.mark.parametrize("lag_kind", ["full", "dups", "u40"])
@pytest.mark.parametrize("explicit_ids", [False, True])
def test_literal_equals_rounds_flat(lag_kind, explicit_ids):
    """The round formulation equals the literal scan on
    larger random batches, wrap-around included."""
    rng = np.random.default_rng(hash((lag_kind, explicit_ids)) & 0xFFFF)
    for C, pmax in [(1, 40), (3, 100), (17, 300), (64, 64), (100, 30)]:
        args = _random_flat(rng, T=40, pmax=pmax, C=C, lag_kind=lag_kind,
                            explicit_ids=explicit_ids, n_sets=3)
        a = oracle.assign_batch(*args, mode=oracle.LITERAL, want_totals=True)
        b = oracle.assign_batch(*args, mode=oracle.ROUNDS, want_totals=True, threads=4)
        for k in ("out_partition", "out_consumer_rank", "out_total", "out_count"):
            assert np.array_equal(a[k], b[k]), (k, C, pmax)
        assert a["assigned"] == b["assigned"]


def test_topic_range_sampling_leaves_others_untouched():
    rng = np.random.default_rng(7)
    args = _random_flat(rng, T=30, pmax=20, C=5, lag_kind="u40")
    full = oracle.assign_batch(*args, mode=oracle.ROUNDS)
    part = oracle.assign_batch(*args, mode=oracle.ROUNDS, topic_lo=10, topic_hi=20)
    off = args[0]
    lo, hi = int(off[10]), int(off[20])
    assert np.array_equal(full["out_partition"][lo:hi], part["out_partition"][lo:hi])
    assert (part["out_partition"][:lo] == -2).all() and (part["out_partition"][hi:] == -2).all()
