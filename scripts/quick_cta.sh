#!/bin/bash
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout=300 -x -k "cta_bin or mixed or random_shape or realistic or config5 or grouping or chunked" 2>&1 | tail -4
python - <<'PY'
import sys, time
sys.path.insert(0, "tests")
import numpy as np, torch
import kafka_lag_based_assignor_b200 as pkg
import oracle
from helpers import uniform_batch, args_of
rng = np.random.default_rng(1)
for (T, P, C) in ((10000, 1000, 64), (2000, 4000, 500), (5000, 400, 4096)):
    b = uniform_batch(T, P, C, rng.integers(0, 1 << 30, size=T * P, dtype=np.int64))
    with pkg.FlatAssignor(0) as f:
        f.set_option(pkg.OPT_PROFILE, 1)
        f.upload(*args_of(b))
        for _ in range(3): f.execute()
        f.sync()
        prof = f.profile()
        got = f.download()
    exp = oracle.assign_batch(*args_of(b), threads=16, topic_lo=0, topic_hi=50)
    n = 50 * P
    ok = np.array_equal(got["out_partition"][:n], exp["out_partition"][:n]) and np.array_equal(got["out_consumer_rank"][:n], exp["out_consumer_rank"][:n])
    ms = prof["k_cta_topics"][0] / prof["k_cta_topics"][1]
    print(f"T={T} P={P} C={C}: k_cta_topics {ms:.3f} ms -> {T*P/ms/1e6:.1f} G assignments/s, parity {ok}")
PY
