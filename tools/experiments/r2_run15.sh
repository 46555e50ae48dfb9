#!/bin/bash
# round 2, run 15: G2 multi-scalar accumulation on shared-memory cells (k_g2_msm_acc_sm, 4 blocks/SM)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "msm or all_entry or edges or golden or config3 or config1" 2>&1 | tail -3
timeout 600 python tools/kbench.py r2m_g2cells combine 2>&1 | tail -1 | cut -c1-500
TCB200_MSM_ALGO=4 timeout 600 python tools/kbench.py r2m_g2reg combine 2>&1 | tail -1 | cut -c1-500
N3=16384 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_launches_combine.csv python tools/prof_small.py combine > /dev/null 2>&1
N3=2048 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_launches_combine_2048.csv python tools/prof_small.py combine > /dev/null 2>&1
for f in r2m_launches_combine r2m_launches_combine_2048; do python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/$f.csv')))
hdr=None
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        nm=d['Kernel Name'].split('(')[0]
        if any(k in nm for k in ('lagrange','msm','g2_sum')): print('$f', nm, d['Grid Size'], round(float(d['Metric Value'])/1e6,3),'ms')
PY
done
