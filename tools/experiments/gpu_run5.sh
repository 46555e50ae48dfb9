#!/bin/bash
# session-2 GPU run 5: batch-affine MSM: parity, bench, per-kernel times for the three MSM algorithms
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r5_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r5_pytest_gpu.txt
tail -3 gpurun_out/r5_pytest_gpu.txt
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r5_bench_main.json 2> gpurun_out/r5_bench_main.err; cat gpurun_out/r5_bench_main.json
for a in 0 1; do
TCB200_MSM_ALGO=$a ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r5_launches_algo$a.csv python tools/prof_small.py combine > /dev/null 2>&1
TCB200_MSM_ALGO=$a ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r5_launches_g1_algo$a.csv python tools/prof_small.py decrypt > /dev/null 2>&1
echo "== algo $a"; grep -h -E "k_" gpurun_out/r5_launches_algo$a.csv gpurun_out/r5_launches_g1_algo$a.csv | awk -F'","' '{print $5, $9, $NF}' | sed 's/([a-z][^)]*)//' | cut -c1-120
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'msm_acc_ba' -c 2 -o gpurun_out/r5_prof python tools/prof_small.py all > gpurun_out/r5_ncu.log 2>&1
ncu -i gpurun_out/r5_prof.ncu-rep --page details --csv > gpurun_out/r5_prof_details.csv 2>/dev/null
for k in k_g2_msm_acc_ba k_g1_msm_acc_ba; do
  ncu -i gpurun_out/r5_prof.ncu-rep --page source --csv -k regex:$k 2>/dev/null | python profiles/agg_source.py gpurun_out/r5_${k}_by_opcode.json > /dev/null 2>&1
done
rm -f gpurun_out/*.ncu-rep
