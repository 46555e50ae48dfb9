#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r10_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r10_pytest_gpu.txt
tail -3 gpurun_out/r10_pytest_gpu.txt
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r10_bench.json 2> gpurun_out/r10_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r10_bench.json')); print(d['value'], d['ms_per_step'], {k:v['ms_per_launch'] for k,v in d['roofline']['kernels'].items()}, d['combine']['ms_per_step'])"
