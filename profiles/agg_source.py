"""Aggregates `ncu --page source --csv` (per-SASS-instruction) into per-opcode totals so the result
fits the gpurun return limit.  Usage on the box:  ncu -i rep --page source --csv | python profiles/agg_source.py out.json"""
import csv
import json
import re
import sys
from collections import defaultdict

rows = list(csv.reader(sys.stdin))
hdr_i = next(i for i, r in enumerate(rows) if any("Source" == c or "SASS" in c for c in r))
hdr = rows[hdr_i]
out = {"header": hdr, "n_rows": len(rows) - hdr_i - 1}
src_col = next((i for i, c in enumerate(hdr) if c.strip() in ("Source", "SASS")), 1)
num_cols = [i for i, c in enumerate(hdr) if i != src_col]
agg = defaultdict(lambda: defaultdict(float))
cnt = defaultdict(int)
for r in rows[hdr_i + 1:]:
    if len(r) <= src_col:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[src_col])
    if not m:
        continue
    op = m.group(2)
    full = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[src_col]).group(2)
    key = op if op not in ("IMAD",) else ("IMAD.WIDE" if "WIDE" in full else "IMAD")
    cnt[key] += 1
    for i in num_cols:
        if i < len(r):
            try:
                agg[key][hdr[i]] += float(r[i].replace(",", ""))
            except ValueError:
                pass
out["static_instr_count"] = dict(cnt)
out["by_opcode"] = {k: dict(v) for k, v in agg.items()}
json.dump(out, open(sys.argv[1], "w"), indent=1)
tot = defaultdict(float)
for k, v in agg.items():
    for c, x in v.items():
        tot[c] += x
print(json.dumps({"totals": dict(tot)}, indent=1)[:3000])
