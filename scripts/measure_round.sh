#!/bin/bash
timeout -s KILL 400 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -2
# Round-end measurement pass, run on a B200 through gpurun:
#   gpurun --timeout 2400 -- 'bash scripts/measure_round.sh'
# Writes everything under gpurun_out/ (copied into profiles/ afterwards).
mkdir -p gpurun_out
cat MEASURED_PEAKS.json 2>/dev/null
nproc
scripts/ubench/membw > gpurun_out/membw.txt 2>&1; cat gpurun_out/membw.txt
for c in 4 2; do
  timeout -s KILL 240 python bench.py --config $c > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err
  tail -c 300 gpurun_out/bench_c$c.err
done
for c in 3 5 6 7; do   # CPU baselines of these are in DESIGN.md (oracle unchanged); keep the box time for ncu
  timeout -s KILL 150 python bench.py --config $c --no-cpu-baseline > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err
  tail -c 300 gpurun_out/bench_c$c.err
done
timeout -s KILL 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_c4.json 2> gpurun_out/bench_reference_c4.err
(timeout 100 python scripts/group_profile.py 4; timeout 100 python scripts/group_profile.py 5) > gpurun_out/group_profile.txt 2>&1
for c in 4 6 2; do
  timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_c$c.csv python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --hot-path-only > /dev/null 2>&1
done
cap() {  # kernel regex, config, skip, output name; the report is reduced to its raw page on the box (gpurun_out is capped at 64 MiB)
  timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:$1 -s $3 -c 1 -f \
    -o gpurun_out/$4 python bench.py --config $2 --steps 1 --warmup 3 --no-cpu-baseline --hot-path-only > /dev/null 2>&1
  n=${4#prof_}
  ncu -i gpurun_out/$4.ncu-rep --page raw --csv > gpurun_out/raw_$n.csv 2>/dev/null
  rm -f gpurun_out/$4.ncu-rep
}
cap k_warp_topics 4 3 prof_warp_c4
cap k_big_scatter 6 10 prof_scatter_c6
cap k_big_tile_hist 6 10 prof_tilehist_c6
cap k_big_finalize 6 1 prof_finalize_c6
cap k_big_range 6 1 prof_range_c6
if [ -z "$FEWCAPS" ]; then
cap k_big_fill 2 1 prof_fill_c2
cap k_cta_topics 5 14 prof_cta8_c5   # 4 size classes per execute: launch 14 = EPT 8 of the timed step
fi
LAGASSIGN_LIBRARY=$PWD/variants/liblagassign_timing.so timeout -s KILL 300 python bench.py --config 6 --steps 1 --warmup 3 \
  --no-cpu-baseline 2>&1 | grep "^pass" > gpurun_out/scatter_phase_cycles.txt
# compute-sanitizer memcheck over the whole GPU suite (log kept; summary line printed); NOSAN=1 skips it
[ -n "$NOSAN" ] || timeout -s KILL 420 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck_full.log \
  python -m pytest tests/test_gpu_parity.py -q -x --timeout=400 2>&1 | tail -2
[ -n "$NOSAN" ] || tail -1 gpurun_out/memcheck_full.log
ls -la gpurun_out
