#!/bin/bash
# A/B of experimental builds of the library (variants/*.so) against the committed one:
# parity subset + config-4 kernel time for each. Run through gpurun.
for v in "" imad4 imad2; do
  if [ -n "$v" ]; then export LAGASSIGN_LIBRARY=$PWD/variants/liblagassign_$v.so; else unset LAGASSIGN_LIBRARY; fi
  echo "== variant: ${v:-committed}"
  timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -x --timeout=280 -k "warp_bin or config4_scaled or config3" 2>&1 | tail -1
  for i in 1; do
    timeout -s KILL 200 python bench.py --config 4 --steps 20 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4e'%d['value'], d['kernel_ms_per_step'], d['clocks'].get('sm_mhz'))"
  done
done
