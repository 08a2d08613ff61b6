#!/bin/bash
# ncu launch lists of the hot path only (see scripts/measure_round.sh). Run through gpurun.
mkdir -p gpurun_out
for c in 4 6 2; do
  timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_c$c.csv python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --hot-path-only > /dev/null 2>&1
  wc -l gpurun_out/launches_c$c.csv
done
