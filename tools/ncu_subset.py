"""Reduces `ncu -i rep --page raw --csv` (stdin) to a JSON of the metrics the roofline discussion uses, per kernel
launch:  ncu -i rep --page raw --csv | python tools/ncu_subset.py out.json"""
import csv
import json
import re
import sys

KEEP = re.compile(r"^(dram__bytes_(read|write)\.sum|gpu__time_duration\.sum|launch__(registers_per_thread|grid_size|block_size|occupancy_limit_registers|waves_per_multiprocessor)|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active|sm__inst_executed_pipe_(fmaheavy|fma|alu|lsu|xu)\.|sm__pipe_(fmaheavy|fma|alu)_cycles_active|"
                  r"smsp__inst_executed\.sum|smsp__inst_executed\.avg\.per_cycle_active|sm__inst_executed\.avg\.per_cycle_elapsed|l1tex__t_sector_hit_rate\.pct|lts__t_sector_hit_rate\.pct|"
                  r"l1tex__t_sector_pipe_lsu_mem_local_op_(ld|st)_hit_rate\.pct|smsp__average_warps?_issue_stalled_.*_per_issue_active\.ratio|smsp__thread_inst_executed_per_inst_executed\.ratio|"
                  r"sm__cycles_elapsed\.max|smsp__cycles_active\.avg|launch__kernel_name)")
rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, units = rows[hi], rows[hi + 1]
out = []
for r in rows[hi + 2:]:
    if len(r) != len(hdr):
        continue
    d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
    for h, u, v in zip(hdr, units, r):
        if KEEP.match(h):
            try:
                d[h + (f" [{u}]" if u else "")] = float(v.replace(",", ""))
            except ValueError:
                d[h] = v
    out.append(d)
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(len(out), "launches")
