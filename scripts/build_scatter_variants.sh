#!/bin/bash
# Variants of the device-wide radix path for A/B runs on a B200 (variants/ is git-ignored, shipped by gpurun):
#   liblagassign_sc256.so     k_big_scatter with 256 threads x 16 items (default: 512 x 8)
#   liblagassign_sctiming.so  default shape with per-phase cycle counters printed by CTA 0 / 200 (-DLAGASSIGN_RT_TIMING)
# Build the normal library first (python kafka-lag-based-assignor_b200/_build.py).
set -e
cd "$(dirname "$0")/.."
B=kafka-lag-based-assignor_b200/build
S=kafka-lag-based-assignor_b200/csrc
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
mkdir -p variants /tmp/lagassign_variants
(nvcc $F -DLAGASSIGN_SC_THREADS=256 -c -o /tmp/lagassign_variants/launch_big_sc256.o $S/launch_big.cu) &
(nvcc $F -DLAGASSIGN_RT_TIMING -c -o /tmp/lagassign_variants/launch_big_timing.o $S/launch_big.cu) &
for extra in "$@"; do   # name=flags, e.g. sc1024="-DLAGASSIGN_SC_THREADS=1024"
  n=${extra%%=*}; fl=${extra#*=}
  (nvcc $F $fl -c -o /tmp/lagassign_variants/launch_big_$n.o $S/launch_big.cu) &
done
wait
L="nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -Xcompiler -fPIC"
$L -o variants/liblagassign_sc256.so $(ls $B/*.o | grep -v launch_big.o) /tmp/lagassign_variants/launch_big_sc256.o
$L -o variants/liblagassign_sctiming.so $(ls $B/*.o | grep -v launch_big.o) /tmp/lagassign_variants/launch_big_timing.o
for extra in "$@"; do
  n=${extra%%=*}
  $L -o variants/liblagassign_$n.so $(ls $B/*.o | grep -v launch_big.o) /tmp/lagassign_variants/launch_big_$n.o
done
ls -la variants | grep -E "sc|timing"
