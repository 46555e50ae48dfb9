#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --section SourceCounters --section WarpStateStats --clock-control none --import-source on -k regex:'k_verify_g2_quad|k_hash_g2' -s 4 -c 2 -o gpurun_out/r9_prof python bench.py --steps 1 --warmup 1 --no-cpu --no-combine > gpurun_out/r9_ncu.log 2>&1
for k in k_verify_g2_quad k_hash_g2; do
  ncu -i gpurun_out/r9_prof.ncu-rep --page source --csv -k regex:$k 2>/dev/null | python profiles/agg_by_addr.py gpurun_out/r9_${k}_by_addr.json 4096
done
rm -f gpurun_out/*.ncu-rep; tail -3 gpurun_out/r9_ncu.log
