#!/bin/bash
# A/B of variants/*.so (scripts/build_variants.sh) against the committed library on a B200: the GPU parity tests
# that touch the changed path, then the configs whose dominant kernel it is. Run through gpurun.
run() {  # label, pytest -k expression, configs...
  echo "== $1"; shift
  timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -q -x --timeout=380 -k "$1" 2>&1 | tail -1; shift
  for c in "$@"; do
    timeout -s KILL 200 python bench.py --config $c --steps 10 --no-cpu-baseline 2>/dev/null | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('config', d['config']['config'], '%.4e'%d['value'], {k: round(v, 4) for k, v in d['kernel_ms_per_step'].items()})"
  done
}
K1="warp_bin or realistic or random_shape or config3 or wraparound"
K2="cta_bin or device_bin or config2 or mixed or single_consumer or wraparound or realistic"
unset LAGASSIGN_LIBRARY; run "committed" "$K1 or $K2 or config5 or grouping" 4 7 2 3 5 6
export LAGASSIGN_LIBRARY=$PWD/variants/liblagassign_nostl.so; run "warp kernel without address-taken key/pay arrays" "$K1" 4 7 3
export LAGASSIGN_LIBRARY=$PWD/variants/liblagassign_groups.so; run "four topics per warp (+ by-copy general sort)" "$K1 or config4 or mixed or chunked" 4 7 3
export LAGASSIGN_LIBRARY=$PWD/variants/liblagassign_packedfill.so; run "packed warp fill" "$K1" 7 3
export LAGASSIGN_LIBRARY=$PWD/variants/liblagassign_optw2.so; run "optimistic two-word fill" "$K2" 2
export LAGASSIGN_LIBRARY=$PWD/variants/liblagassign_fill1024.so; run "optimistic two-word fill, 1024-thread k_big_fill" "device_bin or config2 or mixed or single_consumer or wraparound" 2
export LAGASSIGN_LIBRARY=$PWD/variants/liblagassign_histskip.so; run "k_big_hist with per-tile constant-digit skip" "device_bin or config2 or config5 or mixed or grouping" 5 2 6
# the one-word CTA fill needs small lags and 256 < C < P: not a bench config, timed here
for v in "" fillw1; do
  if [ -n "$v" ]; then export LAGASSIGN_LIBRARY=$PWD/variants/liblagassign_$v.so; else unset LAGASSIGN_LIBRARY; fi
  echo "== CTA fill, small lags: ${v:-committed}"
  [ -n "$v" ] && timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -q -x --timeout=380 -k "$K2" 2>&1 | tail -1
  python - <<'PY'
import sys
sys.path.insert(0, "tests")
import numpy as np
import kafka_lag_based_assignor_b200 as pkg
from helpers import uniform_batch, args_of
rng = np.random.default_rng(1)
for (T, P, C) in ((2000, 4000, 500), (4000, 2000, 300), (1, 1 << 20, 1024)):
    b = uniform_batch(T, P, C, rng.integers(0, 1000, size=T * P, dtype=np.int64))
    with pkg.FlatAssignor(0) as f:
        f.set_option(pkg.OPT_PROFILE, 1)
        f.upload(*args_of(b))
        for _ in range(3): f.execute()
        f.sync()
        print(T, P, C, {k: round(v[0] / v[1], 4) for k, v in f.profile().items() if k in ("k_cta_topics", "k_big_fill")})
PY
done
