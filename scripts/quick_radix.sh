#!/bin/bash
# quick check of the device-wide path: parity tests + sort-only bench (+ phase counters)
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout=300 -x -k "device_bin or mixed or config2 or config5 or single_consumer or wraparound or fused" 2>&1 | tail -3
for c in 6 2; do
  echo "== config $c"
  timeout -s KILL 300 python bench.py --config $c --steps 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, d['roofline']['kernel'], d['roofline']['frac'], d['parity_spot_check'])"
done
if [ -f variants/liblagassign_timing.so ]; then
  LAGASSIGN_LIBRARY=$PWD/variants/liblagassign_timing.so timeout -s KILL 300 python bench.py --config 6 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep "^pass" | tail -2
fi
