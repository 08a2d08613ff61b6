import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import kafka_lag_based_assignor_b200 as pkg
import oracle
from helpers import random_flat, args_of

rng = np.random.default_rng(1)
C = int(sys.argv[1]) if len(sys.argv) > 1 else 3
pmax = int(sys.argv[2]) if len(sys.argv) > 2 else 9000
b = random_flat(rng, T=3, pmin=4097, pmax=pmax, C=C, lag_kind="u40", explicit_ids=None, n_sets=2, full_sets=True)
with pkg.FlatAssignor(0, want_totals=True) as f:
    got = f.assign_batch(*args_of(b))
    print(f.stats())
exp = oracle.assign_batch(*args_of(b), want_totals=True)
for k in ("out_partition", "out_consumer_rank", "out_total", "out_count"):
    print(k, np.array_equal(got[k], exp[k]))
