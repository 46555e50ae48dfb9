#!/bin/bash
# round 2, run 10: final exponentiation with the saved powers / prefix products in shared memory: time and DRAM traffic
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "miller or both_pairing or selftest or golden or edges" 2>&1 | tail -2
timeout 600 python tools/kbench.py r2i_fecells verify 2>&1 | tail -1 | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread --clock-control none -k regex:'k_final_exp|k_miller' --csv --log-file gpurun_out/r2i_fe_traffic.csv python tools/prof_verify.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2i_fe_traffic.csv')))
hdr=None
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        print(d['Kernel Name'][:22], d['Metric Name'], d['Metric Value'], d['Metric Unit'])
PY
