#!/bin/bash
# One measurement step on a B200: parity suite, A/B of the hash_g2 / verify knobs with tools/kbench.py, one bench line.
# Usage: bash tools/gpu_step.sh <tag>
T=${1:-step}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.txt
tail -4 gpurun_out/${T}_pytest_gpu.txt
for ha in 1 0; do for vh in 1 0; do
  HASH_ALGO=$ha VERIFY_HASH=$vh python tools/kbench.py ${T}_h${ha}v${vh} verify 2>&1 | tail -1 | cut -c1-900
done; done
python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; cut -c1-700 gpurun_out/${T}_bench_n1.json; tail -2 gpurun_out/${T}_bench_n1.err
