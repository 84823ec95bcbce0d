#!/usr/bin/env python
"""Key counters per kernel from an `ncu --page raw --csv` export, one JSON line per launch.
usage: ncu_summary.py raw.csv [kernel-substring]"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = sys.argv[2] if len(sys.argv) > 2 else ""
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_global_ld.sum"]
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    if want not in name:
        continue
    o = {"kernel": name.split("(")[0]}
    for k in KEYS:
        if k in col:
            o[k] = r[col[k]] + " " + units[col[k]]
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                v = float(r[col[h]])
            except ValueError:
                continue
            if v >= 0.15:
                o["stall_" + h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(v, 2)
    if len(sys.argv) > 3 and sys.argv[3] == "pipes":
        for h in hdr:
            if any(t in h for t in ("pipe_", "wavefronts", "l1tex__t_requests", "l1tex__t_sectors_pipe", "lsu_mem_global_op", "throughput.avg.pct", "l1tex__lsu", "l1tex__data_pipe")) and "pct" in h or "wavefronts.sum" in h:
                o[h] = r[col[h]]
    print(json.dumps(o))
