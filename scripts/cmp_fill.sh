#!/bin/bash
for v in default newfill; do
  for c in 7 2; do
    echo "== $v config $c"
    if [ $v != default ]; then export LAGASSIGN_LIBRARY=$PWD/variants/liblagassign_$v.so; else unset LAGASSIGN_LIBRARY; fi
    timeout -s KILL 300 python bench.py --config $c --steps 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4e'%d['value'], {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, d['parity_spot_check'])"
  done
done
