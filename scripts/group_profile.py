"""Per-kernel times of the grouping step (LAGASSIGN_OPT_GROUP) at a BASELINE config.  usage: group_profile.py [config]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import kafka_lag_based_assignor_b200 as pkg
from helpers import args_of, baseline_config
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
b, desc = baseline_config(cfg)
with pkg.FlatAssignor(0) as f:
    f.set_option(pkg.OPT_GROUP, 1)
    f.set_option(pkg.OPT_PROFILE, 1)
    f.upload(*args_of(b))
    for _ in range(3):
        f.execute(); f.sync()
    print(desc)
    for k, (ms, n) in sorted(f.profile().items(), key=lambda kv: -kv[1][0]):
        print("  %-40s %8.4f ms  %d launches" % (k, ms, n))
