#!/bin/bash
# one-GPU check of the current tree: the whole GPU suite, then the headline configs (short form)
#   gpurun --timeout 1200 -- 'bash scripts/quick_check.sh [configs...]'
timeout -s KILL 400 python -m pytest tests -m gpu -q -x --timeout=300 2>&1 | tail -4
mkdir -p gpurun_out
for c in ${@:-4 6 5 2}; do
  timeout -s KILL 150 python bench.py --config $c --steps 10 --no-cpu-baseline 2>gpurun_out/err_c$c.txt | tail -1 > gpurun_out/quick_c$c.json
  python - <<PY || tail -5 gpurun_out/err_c$c.txt
import json
d=json.load(open("gpurun_out/quick_c$c.json"))
r=d["roofline"]
print("config $c value %.4e ms %.4f (unprofiled %.4f) e2e %.3e (int32 %.3e) parity %s | %s frac %s |"%(d["value"],d["ms_per_step"],d["ms_per_step_unprofiled"],d["e2e"]["value"],d["e2e"].get("value_int32",0),d["parity_spot_check"],r["kernel"],r["frac"] and round(r["frac"],3)),
      {k: round(v,4) for k,v in d["kernel_ms_per_step"].items()})
nr=d.get("next_rows")
if nr: print("   next rows:", {k:{kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk!="note"} for k,v in nr.items()})
PY
done
