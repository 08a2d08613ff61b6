"""SURVEY 8f row 3 at BASELINE config 4's scale: lagassign_invert_subscriptions over M members x T topics, everybody
subscribed to everything (M * T subscription entries), host arrays in and out.  usage: invert_bench.py [M] [T]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import kafka_lag_based_assignor_b200 as pkg
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
sub_off = (np.arange(M + 1, dtype=np.int64) * T)
sub_topic = pkg.pinned_empty(M * T, np.int32)            # pinned: the 4 B per entry cross PCIe at full speed
row = np.arange(T, dtype=np.int32)
for m in range(M):
    sub_topic[m * T:(m + 1) * T] = row
pkg.invert_subscriptions(T, sub_off[:2], sub_topic[:T])   # warm-up (context, module load)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter()
    out = pkg.invert_subscriptions(T, sub_off, sub_topic)
    best = min(best, time.perf_counter() - t0)
ok = len(out["cset_off"]) == 2 and int(out["cset_off"][1]) == M and bool((out["topic_cset"] == 0).all()) and \
    bool((out["cset_rank"] == np.arange(M)).all())
print("invert_subscriptions: %d members x %d topics = %.3g entries (%.2f GB in): %.1f ms per call = %.1f ns per entry, "
      "1 set of %d members out: %s" % (M, T, M * T, M * T * 4 / 1e9, best * 1e3, best / (M * T) * 1e9, M, "ok" if ok else "MISMATCH"))
