#!/bin/bash
# round 2, run 6: final-exponentiation kernel at 3 and 4 blocks/SM (168 / 128 registers)
mkdir -p gpurun_out
L=threshold_crypto_b200/csrc
for v in fe3 fe4; do TCB200_LIB=$L/libtcb200_$v.so timeout 600 python tools/kbench.py r2f_$v verify 2>&1 | tail -1 | cut -c1-420; done
