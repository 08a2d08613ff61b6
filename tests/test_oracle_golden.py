"""The oracle against the reference's own vectors (tests/golden/reference_kats.json,
transcribed from LagBasedPartitionAssignorTest.java and README.md) and against the
committed regression vectors. CPU only."""
import json
import os

import numpy as np
import pytest

import oracle
from oracle import flat as oflat
from oracle import oracle_py


def _load(golden_dir, name):
    with open(os.path.join(golden_dir, name)) as f:
        return json.load(f)


def _as_inputs(case):
    lags = {t: [tuple(e) for e in v] for t, v in case["partition_lag_per_topic"].items()}
    subs = [(m, list(ts)) for m, ts in case["subscriptions"]]
    return lags, subs


def _norm(res, order_sensitive=True):
    if order_sensitive:
        return {m: [tuple(tp) for tp in v] for m, v in res.items()}
    return {m: sorted(tuple(tp) for tp in v) for m, v in res.items()}


def _run_c_oracle(lags, subs, mode):
    fl, topic_order, members = oflat.flatten(lags, subs)
    r = oracle.assign_batch(fl["topic_off"], fl["lag"], fl["partition"], fl["topic_cset"],
                            fl["cset_off"], fl["cset_rank"], mode=mode, want_totals=True)
    res = oflat.unflatten(fl, topic_order, members, r["out_partition"], r["out_consumer_rank"],
                          [m for m, _ in subs])
    return res, r, fl, topic_order, members


@pytest.mark.parametrize("mode", [oracle.LITERAL, oracle.ROUNDS])
def test_reference_assign_kats(golden_dir, mode):
    kats = _load(golden_dir, "reference_kats.json")
    for case in kats["assign"]:
        lags, subs = _as_inputs(case)
        exp = _norm(case["expected"], case["order_sensitive"])
        got_py = oracle_py.assign({t: list(v) for t, v in lags.items()}, subs)
        assert _norm(got_py, case["order_sensitive"]) == exp, case["name"]
        got_c, r, fl, topic_order, members = _run_c_oracle(lags, subs, mode)
        assert _norm(got_c, case["order_sensitive"]) == exp, case["name"]
        if "expected_totals" in case:
            # totals per member of the (single) topic, README:54-57
            (topic, tot), = case["expected_totals"].items()
            ranks = oracle_py.rank_members(m for m, _ in subs)
            for m, v in tot.items():
                assert int(r["out_total"][ranks[m]]) == v


def test_readme_list_order_is_lag_descending(golden_dir):
    """C1's list is [p2, p1] (p2 has the larger lag), README shows a set."""
    case = _load(golden_dir, "reference_kats.json")["assign"][1]
    lags, subs = _as_inputs(case)
    got = oracle_py.assign(lags, subs)
    assert got["C1"] == [("t0", 2), ("t0", 1)]


@pytest.mark.parametrize("mode", [oracle.LITERAL, oracle.ROUNDS])
def test_reference_balance_properties(golden_dir, mode):
    """Test:170-173 and Test:221-224: max assigned <= min assigned + 1."""
    for case in _load(golden_dir, "reference_kats.json")["balance_properties"]:
        lags, subs = _as_inputs(case)
        got, *_ = _run_c_oracle(lags, subs, mode)
        sizes = [len(v) for v in got.values()]
        assert max(sizes) <= min(sizes) + 1, case["name"]
        assert sum(sizes) == sum(len(v) for v in lags.values())


def test_reference_compute_partition_lag(golden_dir):
    for k in _load(golden_dir, "reference_kats.json")["compute_partition_lag"]:
        assert oracle.compute_partition_lag(k["committed"], k["begin"], k["end"], k["reset"]) == k["expected"], k
        assert oracle_py.compute_partition_lag(k["committed"], k["begin"], k["end"], k["reset"]) == k["expected"], k


def test_compute_partition_lag_edges():
    assert oracle.compute_partition_lag(None, 5, 10, "LATEST") == 0          # equalsIgnoreCase
    assert oracle.compute_partition_lag(None, 5, 10, "none") == 5            # anything else = earliest
    assert oracle.compute_partition_lag(20, 5, 10, "latest") == 0            # clamp
    big = (1 << 63) - 1
    # end - next wraps negative -> clamped to 0, like Java
    assert oracle.compute_partition_lag(-big, 0, big, "latest") == oracle_py.compute_partition_lag(-big, 0, big, "latest") == 0


@pytest.mark.parametrize("mode", [oracle.LITERAL, oracle.ROUNDS])
def test_regression_vectors(golden_dir, mode):
    for case in _load(golden_dir, "oracle_regression.json")["cases"]:
        lags, subs = _as_inputs(case)
        exp = _norm(case["expected"])
        got, r, fl, topic_order, members = _run_c_oracle(lags, subs, mode)
        assert _norm(got) == exp, case["name"]
        # totals / counts blocks
        ranks = {m: i for i, m in enumerate(members)}
        tc = 0
        for ti, t in enumerate(topic_order):
            s = int(fl["topic_cset"][ti])
            cs = fl["cset_rank"][int(fl["cset_off"][s]):int(fl["cset_off"][s + 1])]
            for j, rk in enumerate(cs):
                m = members[int(rk)]
                e_tot, e_cnt = case["expected_totals"][t][m]
                assert int(r["out_total"][tc + j]) == e_tot, (case["name"], t, m)
                assert int(r["out_count"][tc + j]) == e_cnt, (case["name"], t, m)
            tc += len(cs)
        assert ranks  # members listed


def test_member_id_order_is_utf16():
    assert oracle.compare_member_ids("C10", "C2") < 0
    assert oracle.compare_member_ids("a", "ab") < 0
    assert oracle.compare_member_ids("b", "ab") > 0
    assert oracle.compare_member_ids("x", "x") == 0
    # surrogate pair (U+1F600) sorts before U+FFFD in UTF-16 code units
    assert oracle.compare_member_ids("\U0001F600", "\uFFFD") < 0
    assert oracle_py.java_compare_to("\U0001F600", "\uFFFD") < 0
    ids = ["C2", "C10", "C1", "\uFFFD", "\U0001F600", "", "c"]
    r = oracle.rank_members(ids)
    rp = oracle_py.rank_members(ids)
    assert [int(x) for x in r] == [rp[m] for m in ids]


def test_java_hashmap_topic_order():
    """'topic1' before 'topic2' (buckets 0 and 1 of a 16-slot table)."""
    assert oracle_py.java_string_hash("topic1") == 0xCC42D8E2
    o = oracle_py.JavaHashMapOrder()
    for k in ["topic2", "topic1"]:
        o.put_if_absent(k)
    assert o.order() == ["topic1", "topic2"]
    # growth past 12 entries keeps every key exactly once
    o = oracle_py.JavaHashMapOrder()
    keys = [f"t{i}" for i in range(100)]
    for k in keys:
        o.put_if_absent(k)
    assert sorted(o.order()) == sorted(keys) and o.cap == 256
