#!/bin/bash
# round 2, run 14: full parity suite + bench on the build whose default pairing engine is k_miller_quad + k_final_exp_sm; ncu of k_final_exp_sm
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err; tail -2 gpurun_out/r2l_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2l_bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'])
for k,v in d['roofline']['kernels'].items(): print(' ', k, round(v['ms_per_launch'],2), round(v.get('frac',0),3))
print('dominant', d['roofline']['dominant_kernel'])
PY
N=65536 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_final_exp_sm' -c 1 -o gpurun_out/r2l_prof python tools/prof_verify.py 2 > gpurun_out/r2l_ncu.log 2>&1
ncu -i gpurun_out/r2l_prof.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_subset.py gpurun_out/r2l_k_final_exp_sm_raw_subset.json
ncu -i gpurun_out/r2l_prof.ncu-rep --page details --csv > gpurun_out/r2l_k_final_exp_sm_details.csv 2>/dev/null
ncu -i gpurun_out/r2l_prof.ncu-rep --page source --csv 2>/dev/null | python profiles/agg_source.py gpurun_out/r2l_k_final_exp_sm_by_opcode.json > /dev/null 2>&1
rm -f gpurun_out/*.ncu-rep
