#!/bin/bash
# Evidence run on one B200: parity suite, bench (both arms), launch list of the bench command, ncu --set full of the hot kernels,
# per-opcode stall attribution.  Usage: bash tools/gpu_final.sh <tag>
T=${1:-r2_final}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.txt
tail -3 gpurun_out/${T}_pytest_gpu.txt
python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; cut -c1-600 gpurun_out/${T}_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err; cut -c1-300 gpurun_out/${T}_bench_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-sharded > /dev/null 2>&1
# verify step = k_hash_g2_point, k_g2_clear, (k_hash_g2 over the flagged items: empty), k_miller_quad, k_final_exp_sm; skip the two warm-up steps
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_miller_quad|k_final_exp_sm|k_hash_g2_point|k_g2_clear' -s 8 -c 4 -o gpurun_out/${T}_prof python bench.py --steps 1 --warmup 1 --no-cpu --no-combine --no-others --no-sharded > gpurun_out/${T}_ncu.log 2>&1
ncu -i gpurun_out/${T}_prof.ncu-rep --page details --csv > gpurun_out/${T}_verify_kernels_details.csv 2>/dev/null
ncu -i gpurun_out/${T}_prof.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_subset.py gpurun_out/${T}_verify_kernels_raw_subset.json
for k in k_miller_quad k_final_exp_sm k_hash_g2_point k_g2_clear; do
  ncu -i gpurun_out/${T}_prof.ncu-rep --page source --csv -k regex:$k 2>/dev/null | python profiles/agg_source.py gpurun_out/${T}_${k}_by_opcode.json > /dev/null 2>&1
done
N3=16384 N4=4096 DEG=1023 N5=65536 timeout 900 ncu --set full --clock-control none -k regex:'k_g2_msm_acc|k_g1_msm_acc|k_commit_eval' -c 3 -o gpurun_out/${T}_prof2 python tools/prof_small.py all > gpurun_out/${T}_ncu2.log 2>&1
ncu -i gpurun_out/${T}_prof2.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_subset.py gpurun_out/${T}_other_kernels_raw_subset.json
rm -f gpurun_out/*.ncu-rep
ls gpurun_out | grep ${T}
