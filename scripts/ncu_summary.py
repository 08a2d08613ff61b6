#!/usr/bin/env python
"""Print the metrics that matter from `ncu -i X.ncu-rep --page raw --csv` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct", "sm__throughput.avg.pct", "launch__registers_per_thread", "launch__occupancy_limit",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct", "smsp__issue_active.avg.pct",
    "smsp__inst_executed.sum", "bank_conflicts", "wavefronts_mem_shared.sum", "inst_executed_op_shared",
    "issue_stalled", "t_sectors_pipe_lsu_mem_global", "t_requests_pipe_lsu_mem_global", "lts__t_sectors_op",
    "inst_executed_pipe_", "thread_inst_executed_per_inst", "local_", "l1tex__throughput", "lts__throughput",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg ", "achieved_occupancy", "sm__pipe"]
for vals in rows[2:]:
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    print("=== kernel:", d.get("Kernel Name"), "id", d.get("ID"))
    for h in hdr:
        if any(p in h for p in pats):
            v = d[h]
            if v in ("", "0", "n/a"): continue
            print(f"  {h} = {v} {u[h]}")
