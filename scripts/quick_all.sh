#!/bin/bash
# full GPU parity suite + a short bench of every config
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout=600 -x 2>&1 | tail -4
for c in 4 3 5 2 6; do
  echo "== config $c"
  timeout -s KILL 300 python bench.py --config $c --steps 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4e'%d['value'], {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, d['roofline']['kernel'], d['roofline']['frac'], d['parity_spot_check'], 'e2e %.3e'%d['e2e']['value'])"
done
