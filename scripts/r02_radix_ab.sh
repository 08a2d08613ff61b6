#!/bin/bash
# round 2: parity + timing of the rebuilt radix pass (TMA loads, atomic ranking + certificate, digit bytes)
K="device_bin or config2 or config5 or config6 or mixed or grouping or random_shape or realistic or full_size or bind_device or multi_device or fused or argument or jni or 8192"
b() { timeout -s KILL 300 python bench.py --config $1 --steps 10 --no-cpu-baseline 2>gpurun_out/err_$2_c$1.txt | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2 config', d['config']['config'], '%.4e'%d['value'], 'parity', d['parity_spot_check'], {k: round(v, 4) for k, v in d['kernel_ms_per_step'].items()})" || tail -5 gpurun_out/err_$2_c$1.txt; }
echo "== default"; unset LAGASSIGN_LIBRARY
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_jni_shim.py -q -x --timeout=600 -m gpu -k "$K" 2>&1 | tail -3
b 6 default; b 2 default; b 5 default
for v in "$@"; do
  echo "== variant $v"
  export LAGASSIGN_LIBRARY=$PWD/variants/liblagassign_$v.so
  case $v in
    *timing*) timeout -s KILL 300 python bench.py --config 6 --steps 1 --warmup 3 --no-cpu-baseline --hot-path-only 2>/dev/null | grep "cycles per tile" | tail -4 ;;
    *) timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout=500 -k "device_bin or config2" 2>&1 | tail -2; b 6 $v ;;
  esac
done
