#!/bin/bash
# session-2 GPU run 3: parity, G1 kernels after the noinline multiply, pairing-kernel variant with the Fp multiply as a function
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r3_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r3_pytest_gpu.txt
tail -3 gpurun_out/r3_pytest_gpu.txt
DEG=1023 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_launches_small.csv python tools/prof_small.py all > /dev/null 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r3_bench_main.json 2> gpurun_out/r3_bench_main.err; cat gpurun_out/r3_bench_main.json
TCB200_LIB=$PWD/threshold_crypto_b200/csrc/libtcb200_pfn.so python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r3_bench_pfn.json 2> gpurun_out/r3_bench_pfn.err; cat gpurun_out/r3_bench_pfn.json
