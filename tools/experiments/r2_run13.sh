#!/bin/bash
# round 2, run 13: final-exponentiation kernel with the sliced Fp2 ops (a), the Fp multiply (b) or both (c) as by-value calls instead of inlined
mkdir -p gpurun_out
L=threshold_crypto_b200/csrc
for v in fe_a fe_b fe_c; do
  TCB200_LIB=$L/libtcb200_$v.so timeout 600 python -m pytest tests -m gpu -x -q -k "both_pairing or selftest" 2>&1 | tail -1
  TCB200_LIB=$L/libtcb200_$v.so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_final_exp|k_miller|k_verify_g2' --csv --log-file gpurun_out/r2k_$v.csv python tools/prof_verify.py > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r2k_$v.csv')))
hdr=None
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        print('$v', d['Kernel Name'][:22], round(float(d['Metric Value'])/1e6,2), 'ms')
PY
done
