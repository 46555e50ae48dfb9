#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r12_bench.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r12_bench.json')); print(d['value'], d['ms_per_step'], d['combine']['value'], d['combine']['ms_per_step'])"
DEG=63 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r12_launches_small.csv python tools/prof_small.py all > /dev/null 2>&1
grep -h -E "k_" gpurun_out/r12_launches_small.csv | awk -F'","' '{print $5, $9, $NF}' | sed 's/([a-z][^)]*)//' | cut -c1-120
