#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "msm or all_entry or edges" > gpurun_out/r6_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r6_pytest_gpu.txt
tail -3 gpurun_out/r6_pytest_gpu.txt
for a in 0; do
TCB200_MSM_ALGO=$a ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r6_launches_algo$a.csv python tools/prof_small.py combine > /dev/null 2>&1
TCB200_MSM_ALGO=$a ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r6_launches_g1_algo$a.csv python tools/prof_small.py decrypt > /dev/null 2>&1
echo "== algo $a"; grep -h -E "k_" gpurun_out/r6_launches_algo$a.csv gpurun_out/r6_launches_g1_algo$a.csv | awk -F'","' '{print $5, $9, $NF}' | sed 's/([a-z][^)]*)//' | cut -c1-120
done
