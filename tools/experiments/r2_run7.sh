#!/bin/bash
# round 2, run 7: parity suite, launch lists of the other configs at full and 1/8 size (what one of 8 GPUs gets), memcheck of the new kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
DEG=1023 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches_other_configs.csv python tools/prof_small.py all > /dev/null 2>&1
DEG=1023 N3=2048 N4=512 N5=8192 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches_other_configs_eighth.csv python tools/prof_small.py all > /dev/null 2>&1
for f in r2g_launches_other_configs r2g_launches_other_configs_eighth; do echo == $f; python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/$f.csv')))
hdr=None
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        nm=d['Kernel Name'].split('(')[0]
        if any(k in nm for k in ('lagrange','msm','sum','commit','decode','finish')): print('  ',nm, d['Grid Size'], round(float(d['Metric Value'])/1e6,3),'ms')
PY
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests -m gpu -x -q -k "miller or both_pairing or edges or golden or poly or codecs" > gpurun_out/r2g_sanitizer_memcheck.txt 2>&1; tail -4 gpurun_out/r2g_sanitizer_memcheck.txt
