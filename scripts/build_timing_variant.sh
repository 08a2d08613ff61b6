#!/bin/bash
# variants/liblagassign_timing.so = the library with per-tile cycle counters in k_big_scatter
# (-DLAGASSIGN_RT_TIMING); used by scripts/measure_round.sh through LAGASSIGN_LIBRARY. Build the normal
# library first (python kafka-lag-based-assignor_b200/_build.py).
cd "$(dirname "$0")/.."
B=kafka-lag-based-assignor_b200/build
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DLAGASSIGN_RT_TIMING \
  -c -o /tmp/launch_big_timing.o kafka-lag-based-assignor_b200/csrc/launch_big.cu || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -Xcompiler -fPIC \
  -o variants/liblagassign_timing.so $(ls $B/*.o | grep -v launch_big.o) /tmp/launch_big_timing.o
