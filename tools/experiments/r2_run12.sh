#!/bin/bash
# round 2, run 12: k_final_exp_sm with 12 slots at 2 and 3 blocks/SM (engine 3)
mkdir -p gpurun_out
L=threshold_crypto_b200/csrc
for v in fesm2 fesm3; do
  TCB200_LIB=$L/libtcb200_$v.so timeout 600 python -m pytest tests -m gpu -x -q -k "miller or both_pairing" 2>&1 | tail -1
  TCB200_LIB=$L/libtcb200_$v.so timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_write.sum --clock-control none -k regex:'k_final_exp|k_miller' --csv --log-file gpurun_out/r2j_$v.csv python tools/prof_verify.py 3 > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r2j_$v.csv')))
hdr=None
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        print('$v', d['Kernel Name'][:22], d['Metric Name'], d['Metric Value'], d['Metric Unit'])
PY
done
