#!/bin/bash
# Turn what scripts/measure_round.sh left in gpurun_out/ into the committed summaries in profiles/.
# usage: bash scripts/make_profiles.sh r01
R=${1:-r01}
cd "$(dirname "$0")/.."
# (scripts/measure_round.sh leaves gpurun_out/raw_<name>.csv: the raw page of each ncu capture)
filt() { grep -v "pcsamp\|per_second\|peak_sustained_elapsed"; }
hdr() { echo "# $1"; echo "# ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 python bench.py --config $3 --steps 1 --warmup 3 --no-cpu-baseline   (gpurun, 1 x B200, scripts/measure_round.sh)"; }
have() { [ -s gpurun_out/raw_$1.csv ]; }   # a capture that was not repeated in this pass keeps its earlier summary
have warp_c4 && (hdr "k_warp_topics, config 4 (100000 topics x 256 partitions)" k_warp_topics 4; python scripts/ncu_summary.py gpurun_out/raw_warp_c4.csv | filt) > profiles/${R}_ncu_k_warp_topics_config4.txt
have finalize_c6 && (hdr "k_big_finalize<true> (finalize + certificate of the sort), config 6" k_big_finalize 6; python scripts/ncu_summary.py gpurun_out/raw_finalize_c6.csv | filt) > profiles/${R}_ncu_k_big_finalize_config6.txt
have range_c6 && (hdr "k_big_range (OR / AND of every topic's keys: which digit passes can be skipped), config 6" k_big_range 6; python scripts/ncu_summary.py gpurun_out/raw_range_c6.csv | filt) > profiles/${R}_ncu_k_big_range_config6.txt
have scatter_c6 && (hdr "k_big_scatter, config 6 (64 topics x 2^20 elements), one digit pass" k_big_scatter 6; python scripts/ncu_summary.py gpurun_out/raw_scatter_c6.csv | filt
 echo; echo "# per-tile cycle counters of the same kernel (build with -DLAGASSIGN_RT_TIMING; thread 0 of CTA 0 / 200; tile = 4096 elements, 2 CTAs/SM):"; tail -4 gpurun_out/scatter_phase_cycles.txt) > profiles/${R}_ncu_k_big_scatter_config6.txt
have tilehist_c6 && (hdr "k_big_tile_hist, config 6" k_big_tile_hist 6; python scripts/ncu_summary.py gpurun_out/raw_tilehist_c6.csv | filt) > profiles/${R}_ncu_k_big_tile_hist_config6.txt
have fill_c2 && (hdr "k_big_fill, config 2 (1 topic x 2^20 partitions, 1024 consumers: 1024 dependent rounds on one SM)" k_big_fill 2; python scripts/ncu_summary.py gpurun_out/raw_fill_c2.csv | filt) > profiles/${R}_ncu_k_big_fill_config2.txt
have cta8_c5 && (hdr "k_cta_topics<8> (1025..2048 partitions per topic), config 5" k_cta_topics 5; python scripts/ncu_summary.py gpurun_out/raw_cta8_c5.csv | filt) > profiles/${R}_ncu_k_cta_topics_config5.txt
[ -s gpurun_out/bench_reference_c4.json ] && cp gpurun_out/bench_reference_c4.json profiles/${R}_bench_reference_config4.json
if [ -s gpurun_out/memcheck_full.log ]; then
  (echo "# compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py (whole GPU suite), scripts/measure_round.sh"; grep -c . gpurun_out/memcheck_full.log | sed 's/^/# log lines: /'; tail -2 gpurun_out/memcheck_full.log) > profiles/${R}_sanitizer.txt
  [ -s gpurun_out/racecheck.log ] && (echo "# compute-sanitizer --tool racecheck, tests test_mixed_bins_and_edge_topics + test_wraparound_totals (scripts/sanitize.sh)"; tail -1 gpurun_out/racecheck.log) >> profiles/${R}_sanitizer.txt
fi
for c in 2 4 6; do python - <<PY
import csv,collections
rows=list(csv.reader(open("gpurun_out/launches_c$c.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
hdr=rows[hi]; ix={h:i for i,h in enumerate(hdr)}
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<len(hdr): continue
    k=r[ix["Kernel Name"]].split("(")[0]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[ix["Metric Value"]])
tot=sum(v[1] for v in agg.values())
with open("profiles/${R}_launches_config$c.txt","w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --hot-path-only\n")
    f.write("# --hot-path-only: the process launches nothing but the execute() steps (3 warm-up + 2 timed + 3 with per-kernel events);\n")
    f.write("# no e2e calls, no next-rows section. Cold-cache serialised times: compare SHARES with bench.py's share_of_step, not absolutes.\n")
    f.write("%-60s %8s %14s %7s\n"%("kernel","launches","total ns","share"))
    for k,(n,t) in agg.items(): f.write("%-60s %8d %14.0f %6.1f%%\n"%(k[:60],n,t,100*t/tot))
PY
done
for c in 2 3 4 5 6 7; do [ -s gpurun_out/bench_c$c.json ] && cp gpurun_out/bench_c$c.json profiles/${R}_bench_config$c.json; done
[ -s gpurun_out/membw.txt ] && (echo "# scripts/ubench/membw: streaming bandwidth probes, 1 x B200 (scripts/measure_round.sh)"; cat gpurun_out/membw.txt) > profiles/${R}_membw.txt
[ -s gpurun_out/group_profile.txt ] && (echo "# scripts/group_profile.py 4 / 5: per-kernel times of a step with LAGASSIGN_OPT_GROUP (the child context's radix kernels are not listed)"; cat gpurun_out/group_profile.txt) > profiles/${R}_group_profile.txt
python - <<PY
import csv, json
def dram(path):
    rows=list(csv.reader(open(path)))
    d=dict(zip(rows[0],rows[2])); u=dict(zip(rows[0],rows[1]))
    mult={"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}
    return int(float(d["dram__bytes_read.sum"])*mult[u["dram__bytes_read.sum"]]+float(d["dram__bytes_write.sum"])*mult[u["dram__bytes_write.sum"]])
out={"_comment":"dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures summarised in profiles/*_ncu_*.txt; read by bench.py for roofline.traffic",
     "config4":{"k_warp_topics":dram("gpurun_out/raw_warp_c4.csv")},
     "config6":{"k_big_scatter":dram("gpurun_out/raw_scatter_c6.csv"),"k_big_tile_hist":dram("gpurun_out/raw_tilehist_c6.csv"),
                "k_big_finalize":dram("gpurun_out/raw_finalize_c6.csv"),"k_big_range":dram("gpurun_out/raw_range_c6.csv")}}
json.dump(out,open("profiles/traffic.json","w"),indent=1)
PY
ls profiles
