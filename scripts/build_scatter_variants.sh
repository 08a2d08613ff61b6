#!/bin/bash
# Variants of the device-wide radix path for A/B runs on a B200 (variants/ is git-ignored, shipped by gpurun).
# usage: build_scatter_variants.sh name="-Dflags ..." [name2="..."]   -> variants/liblagassign_<name>.so
#   e.g. s3t1024="-DLAGASSIGN_SC_THREADS=1024 -DLAGASSIGN_RT_STAGES=3 -DLAGASSIGN_RT_MINBLOCKS=1"
#        timing="-DLAGASSIGN_RT_TIMING"      (per-phase cycle counters printed by CTA 0 / 200)
# Build the normal library first (python kafka-lag-based-assignor_b200/_build.py).
set -e
cd "$(dirname "$0")/.."
B=kafka-lag-based-assignor_b200/build
S=kafka-lag-based-assignor_b200/csrc
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
mkdir -p variants /tmp/lagassign_variants
for extra in "$@"; do
  n=${extra%%=*}; fl=${extra#*=}
  (nvcc $F $fl -Xptxas -v -c -o /tmp/lagassign_variants/launch_big_$n.o $S/launch_big.cu 2>&1 | grep -A2 "k_big_scatterILb0ELb0" | grep -E "registers|spill" | sed "s/^/$n: /") &
done
wait
L="nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -Xcompiler -fPIC"
for extra in "$@"; do
  n=${extra%%=*}
  $L -o variants/liblagassign_$n.so $(ls $B/*.o | grep -v launch_big.o) /tmp/lagassign_variants/launch_big_$n.o
done
ls variants
