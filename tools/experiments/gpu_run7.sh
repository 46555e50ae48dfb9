#!/bin/bash
# session-2 GPU run 7: fast inversion (Pornin optimised binary GCD): parity, bench, per-kernel times
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r7_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r7_pytest_gpu.txt
tail -3 gpurun_out/r7_pytest_gpu.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/r7_bench_main.json 2> gpurun_out/r7_bench_main.err; cat gpurun_out/r7_bench_main.json
DEG=255 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r7_launches_small.csv python tools/prof_small.py all > /dev/null 2>&1
grep -h -E "k_" gpurun_out/r7_launches_small.csv | awk -F'","' '{print $5, $9, $NF}' | sed 's/([a-z][^)]*)//' | cut -c1-120
