#!/bin/bash
# round 2, run 16: k_g2_msm_acc_sm at 2 / 3 / 4 blocks per SM, forced G = 1 and 2; ncu stall picture of the default
mkdir -p gpurun_out
L=threshold_crypto_b200/csrc
for v in g2sm2 g2sm3; do TCB200_LIB=$L/libtcb200_$v.so timeout 600 python tools/kbench.py r2n_$v combine 2>&1 | tail -1 | cut -c1-300; done
N3=16384 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_g2_msm_acc_sm' -c 1 -o gpurun_out/r2n_prof python tools/prof_small.py combine > gpurun_out/r2n_ncu.log 2>&1
ncu -i gpurun_out/r2n_prof.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_subset.py gpurun_out/r2n_k_g2_msm_acc_sm_raw_subset.json
ncu -i gpurun_out/r2n_prof.ncu-rep --page source --csv 2>/dev/null | python profiles/agg_source.py gpurun_out/r2n_k_g2_msm_acc_sm_by_opcode.json > /dev/null 2>&1
rm -f gpurun_out/*.ncu-rep
python - <<'PY'
import json
r=json.load(open('gpurun_out/r2n_k_g2_msm_acc_sm_raw_subset.json'))
for k in r:
    for q,v in k.items():
        if any(s in q for s in ('time_duration','registers_per_thread [','fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed','warps_active','smsp__inst_executed.sum','local_op','lsu.avg.pct_of_peak_sustained_active')): print('  ',q,v)
    st=sorted(((s.split('stalled_')[1].split('_per')[0],v) for s,v in k.items() if 'issue_stalled' in s), key=lambda x:-x[1])[:7]
    print('     ', [(a,round(b,2)) for a,b in st])
PY
