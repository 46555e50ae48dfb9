#!/bin/bash
# round-1 session-2 GPU run: parity tests, MSM vs per-share A/B on configs 3/4/5, ncu of commit_eval and the MSM kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.txt
tail -3 gpurun_out/r2_pytest_gpu.txt
python tools/bench_configs.py > gpurun_out/r2_cfg_msm.json 2> gpurun_out/r2_cfg_msm.err; tail -2 gpurun_out/r2_cfg_msm.err
TCB200_PER_SHARE_TERMS=1 python tools/bench_configs.py > gpurun_out/r2_cfg_pershare.json 2> gpurun_out/r2_cfg_pershare.err
cat gpurun_out/r2_cfg_msm.json gpurun_out/r2_cfg_pershare.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_small.csv python tools/prof_small.py all > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_commit_eval|k_g2_msm_acc|k_g1_msm_acc' -c 3 -o gpurun_out/r2_prof python tools/prof_small.py all > gpurun_out/r2_ncu.log 2>&1
ncu -i gpurun_out/r2_prof.ncu-rep --page details --csv > gpurun_out/r2_prof_details.csv 2>/dev/null
for k in k_commit_eval k_g2_msm_acc k_g1_msm_acc; do
  ncu -i gpurun_out/r2_prof.ncu-rep --page source --csv -k regex:$k 2>/dev/null | python profiles/agg_source.py gpurun_out/r2_${k}_by_opcode.json > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep; rm -f gpurun_out/*.ncu-rep
