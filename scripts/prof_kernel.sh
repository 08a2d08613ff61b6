#!/bin/bash
# one ncu --set full capture of a kernel inside a bench.py step:  prof_kernel.sh <kernel regex> <config> <skip> <name>
mkdir -p gpurun_out
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:$1 -s $3 -c 1 \
  -o gpurun_out/$4 -f python bench.py --config $2 --steps 1 --warmup 3 --no-cpu-baseline --hot-path-only > /dev/null 2>&1
ls -la gpurun_out/$4.ncu-rep
