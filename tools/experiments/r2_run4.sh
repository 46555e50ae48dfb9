#!/bin/bash
# round 2, run 4: bench.py at N=1 (both arms)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; tail -3 gpurun_out/r2d_bench_n1.err; cat gpurun_out/r2d_bench_n1.json | cut -c1-3000
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2d_bench_ref.json 2> gpurun_out/r2d_bench_ref.err; cat gpurun_out/r2d_bench_ref.json | cut -c1-800
