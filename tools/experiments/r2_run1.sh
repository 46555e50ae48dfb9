#!/bin/bash
# round 2, run 1: by-value call ABI for the non-inlined Fp / Fp2S multiplies; pairing kernel with Fp2S ops as calls at 255/168/128 registers
mkdir -p gpurun_out
L=threshold_crypto_b200/csrc
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
TCB200_LIB=$L/libtcb200_r1.so python tools/kbench.py r2a_r1 2>&1 | tail -1
TCB200_LIB=$L/libtcb200_base.so python tools/kbench.py r2a_byval 2>&1 | tail -1
for v in v1 v2 v3; do TCB200_LIB=$L/libtcb200_$v.so python tools/kbench.py r2a_$v verify 2>&1 | tail -1; done
TCB200_LIB=$L/libtcb200_base.so python -m pytest tests -m gpu -x -q 2>&1 | tail -3
