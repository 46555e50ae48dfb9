#!/bin/bash
# round 2, run 2: shared-memory Miller loop (quadsm.cuh) — bit-exactness against the register engine, timing, per-kernel launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "miller or both_pairing or selftest or golden or edges" 2>&1 | tail -5
timeout 600 python tools/kbench.py r2b_smem verify 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_verify.csv python tools/prof_verify.py > /dev/null 2>&1
grep -E "k_miller|k_final|k_verify" gpurun_out/r2b_launches_verify.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120
