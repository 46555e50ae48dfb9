"""Aggregates `ncu --page source --csv` (per-SASS-instruction rows) into address buckets so that stall samples can be
attributed to the device FUNCTIONS of a kernel (map the bucket offsets onto `cuobjdump -sass` of the same build).
Usage on the box:  ncu -i rep --page source --csv -k regex:<kernel> | python profiles/agg_by_addr.py out.json [bucket_bytes]"""
import csv
import json
import sys
from collections import defaultdict

bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[hi]
col = {c: i for i, c in enumerate(hdr)}
want = ["# Samples", "Instructions Executed", "stall_no_inst", "stall_wait", "stall_selected", "stall_dispatch", "stall_short_sb", "stall_long_sb", "stall_lg", "stall_math"]
agg = defaultdict(lambda: defaultdict(float))
addrs = []
for r in rows[hi + 1:]:
    try:
        a = int(r[col["Address"]], 16)
    except (ValueError, IndexError):
        continue
    addrs.append(a)
base = min(addrs)
for r in rows[hi + 1:]:
    try:
        a = int(r[col["Address"]], 16) - base
    except (ValueError, IndexError):
        continue
    b = a // bucket
    agg[b]["n_instr_static"] += 1
    if "IMAD.WIDE" in r[col["Source"]]:
        agg[b]["n_imad_wide_static"] += 1
    for w in want:
        if w in col:
            try:
                agg[b][w] += float(r[col[w]].replace(",", ""))
            except ValueError:
                pass
out = {"bucket_bytes": bucket, "base": hex(base), "buckets": {str(k * bucket): dict(v) for k, v in sorted(agg.items())}}
json.dump(out, open(sys.argv[1], "w"))
print(len(agg), "buckets")
