#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/bench_configs.py > gpurun_out/r13_other_configs.json 2>/dev/null; cat gpurun_out/r13_other_configs.json | tr -d '\n ' ; echo
DEG=15 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r13_launches_small.csv python tools/prof_small.py all > /dev/null 2>&1
grep -h -E "k_g1|k_lagr|k_commit" gpurun_out/r13_launches_small.csv | awk -F'","' '{print $5, $9, $NF}' | sed 's/([a-z][^)]*)//' | cut -c1-100
