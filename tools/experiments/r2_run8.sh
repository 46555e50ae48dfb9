#!/bin/bash
# round 2, run 8: final exponentiation with products / compressed squarings on cells
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "miller or both_pairing or selftest or golden or edges or all_entry" 2>&1 | tail -4
timeout 600 python tools/kbench.py r2h_fesm verify 2>&1 | tail -1 | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches_verify.csv python tools/prof_verify.py 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2h_launches_verify.csv')))
hdr=None
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if any(k in d['Kernel Name'] for k in ('k_miller','k_final','k_verify')): print(d['Kernel Name'][:24], round(float(d['Metric Value'])/1e6,2),'ms')
PY
