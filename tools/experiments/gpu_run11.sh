#!/bin/bash
mkdir -p gpurun_out
for lib in libtcb200.so libtcb200_csq.so; do
TCB200_LIB=$PWD/threshold_crypto_b200/csrc/$lib python bench.py --steps 5 --warmup 3 --no-cpu --no-combine > gpurun_out/r11_bench_$lib.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r11_bench_$lib.json')); print('$lib', d['value'], d['ms_per_step'], {k:v['ms_per_launch'] for k,v in d['roofline']['kernels'].items()})"
done
TCB200_LIB=$PWD/threshold_crypto_b200/csrc/libtcb200_csq.so python -m pytest tests -m gpu -x -q -k "selftest or all_entry or edges" 2>&1 | tail -2
