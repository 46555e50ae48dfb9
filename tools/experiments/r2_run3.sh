#!/bin/bash
# round 2, run 3: full GPU parity suite on the new build (vector loads, two-pass staging, split evaluation, smem Miller), all-config kernel bench,
# ncu --set full of the Miller / final-exponentiation kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python tools/kbench.py r2c_all 2>&1 | tail -1
N=16384 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_miller_quad|k_final_exp_quad' -c 2 -o gpurun_out/r2c_prof python tools/prof_verify.py 2 > gpurun_out/r2c_ncu.log 2>&1
ncu -i gpurun_out/r2c_prof.ncu-rep --page details --csv > gpurun_out/r2c_pairing_kernels_details.csv 2>/dev/null
ncu -i gpurun_out/r2c_prof.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_subset.py gpurun_out/r2c_pairing_kernels_raw_subset.json
for k in k_miller_quad k_final_exp_quad; do
  ncu -i gpurun_out/r2c_prof.ncu-rep --page source --csv -k regex:$k 2>/dev/null | python profiles/agg_source.py gpurun_out/r2c_${k}_by_opcode.json > /dev/null 2>&1
done
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -8
