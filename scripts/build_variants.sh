#!/bin/bash
# Build the compiled-out experimental variants of the library into variants/ (git-ignored, shipped by gpurun):
#   liblagassign_packedfill.so   -DLAGASSIGN_PACKED_WARP_FILL      warp round fill: one-/two-word consumer sort
#   liblagassign_groups.so       -DLAGASSIGN_WARP_GROUPS           warp bin: four 129..256-partition topics per warp (8 lanes x 32 slots)
#   liblagassign_nostl.so        -DLAGASSIGN_WARP_GENERAL_BY_COPY  warp kernel: key[]/pay[] stay in registers on the common path
#   liblagassign_fillw1.so       -DLAGASSIGN_FILL_W1               CTA / big round fill: exact one-word (32-bit) consumer sort for small lags
#   liblagassign_optw2.so        -DLAGASSIGN_FILL_OPTIMISTIC_W2    CTA / big round fill: optimistic truncated two-word sort
#   liblagassign_histskip.so     -DLAGASSIGN_HIST_TILE_SKIP        k_big_hist: one atomic per tile for digits on which the tile agrees
#   liblagassign_fill1024.so     the same + -DLAGASSIGN_BIGFILL_THREADS=1024   k_big_fill with 1024 threads (32 warps for the
#                                sequential round chain instead of 8; consumers 1 / 2 / 4 per thread)
# Both are verified against the oracle under the host emulation (tests/emu; EMU_LIB=... scripts/emu_fuzz.py) and
# have never been timed on a B200. A/B them with:  gpurun -- 'bash scripts/ab_variants.sh'
# Build the normal library first (python kafka-lag-based-assignor_b200/_build.py).
set -e
cd "$(dirname "$0")/.."
B=kafka-lag-based-assignor_b200/build
S=kafka-lag-based-assignor_b200/csrc
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
mkdir -p variants /tmp/lagassign_variants
(nvcc $F -DLAGASSIGN_PACKED_WARP_FILL -c -o /tmp/lagassign_variants/launch_warp.o $S/launch_warp.cu) &
(nvcc $F -DLAGASSIGN_WARP_GROUPS -DLAGASSIGN_WARP_GENERAL_BY_COPY -c -o /tmp/lagassign_variants/launch_warp_groups.o $S/launch_warp.cu) &
(nvcc $F -DLAGASSIGN_WARP_GENERAL_BY_COPY -c -o /tmp/lagassign_variants/launch_warp_copy.o $S/launch_warp.cu) &
for c in 0 1 2 3; do (nvcc $F -DLAGASSIGN_FILL_OPTIMISTIC_W2 -c -o /tmp/lagassign_variants/launch_cta_c$c.o $S/launch_cta_c$c.cu) & done
for c in 0 1 2 3; do (nvcc $F -DLAGASSIGN_FILL_W1 -c -o /tmp/lagassign_variants/launch_cta_w1_c$c.o $S/launch_cta_c$c.cu) & done
(nvcc $F -DLAGASSIGN_FILL_W1 -c -o /tmp/lagassign_variants/launch_big_w1.o $S/launch_big.cu) &
(nvcc $F -DLAGASSIGN_FILL_OPTIMISTIC_W2 -c -o /tmp/lagassign_variants/launch_big.o $S/launch_big.cu) &
(nvcc $F -DLAGASSIGN_HIST_TILE_SKIP -c -o /tmp/lagassign_variants/launch_big_hskip.o $S/launch_big.cu) &
(nvcc $F -DLAGASSIGN_FILL_OPTIMISTIC_W2 -DLAGASSIGN_BIGFILL_THREADS=1024 -c -o /tmp/lagassign_variants/launch_big_1024.o $S/launch_big.cu) &
wait
L="nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -Xcompiler -fPIC"
$L -o variants/liblagassign_groups.so $(ls $B/*.o | grep -v launch_warp.o) /tmp/lagassign_variants/launch_warp_groups.o
$L -o variants/liblagassign_nostl.so $(ls $B/*.o | grep -v launch_warp.o) /tmp/lagassign_variants/launch_warp_copy.o
$L -o variants/liblagassign_packedfill.so $(ls $B/*.o | grep -v launch_warp.o) /tmp/lagassign_variants/launch_warp.o
$L -o variants/liblagassign_optw2.so $(ls $B/*.o | grep -v "launch_cta_c[0-3].o\|launch_big.o") /tmp/lagassign_variants/launch_cta_c[0-3].o /tmp/lagassign_variants/launch_big.o
$L -o variants/liblagassign_fillw1.so $(ls $B/*.o | grep -v "launch_cta_c[0-3].o\|launch_big.o") /tmp/lagassign_variants/launch_cta_w1_c[0-3].o /tmp/lagassign_variants/launch_big_w1.o
$L -o variants/liblagassign_histskip.so $(ls $B/*.o | grep -v launch_big.o) /tmp/lagassign_variants/launch_big_hskip.o
$L -o variants/liblagassign_fill1024.so $(ls $B/*.o | grep -v launch_big.o) /tmp/lagassign_variants/launch_big_1024.o
ls -la variants
