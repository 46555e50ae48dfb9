#!/bin/bash
# final evidence run of the round (one B200): parity suite, bench (both arms), other configs, launch list of the bench
# command, ncu --set full of the two verify kernels, memcheck of the new kernels
T=${1:-r1s2_final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.txt
tail -3 gpurun_out/${T}_pytest_gpu.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/${T}_clocks.csv &
SMI=$!
python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; cat gpurun_out/${T}_bench_n1.json
kill $SMI
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err; cat gpurun_out/${T}_bench_reference_arm.json
python tools/bench_configs.py > gpurun_out/${T}_other_configs.json 2> gpurun_out/${T}_other_configs.err; cat gpurun_out/${T}_other_configs.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
DEG=1023 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_other_configs.csv python tools/prof_small.py all > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_verify_g2_quad|k_hash_g2' -s 4 -c 2 -o gpurun_out/${T}_prof python bench.py --steps 1 --warmup 1 --no-cpu --no-combine > gpurun_out/${T}_ncu.log 2>&1
ncu -i gpurun_out/${T}_prof.ncu-rep --page details --csv > gpurun_out/${T}_verify_kernels_details.csv 2>/dev/null
ncu -i gpurun_out/${T}_prof.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_subset.py gpurun_out/${T}_verify_kernels_raw_subset.json
for k in k_verify_g2_quad k_hash_g2; do
  ncu -i gpurun_out/${T}_prof.ncu-rep --page source --csv -k regex:$k 2>/dev/null | python profiles/agg_source.py gpurun_out/${T}_${k}_by_opcode.json > /dev/null 2>&1
  ncu -i gpurun_out/${T}_prof.ncu-rep --page source --csv -k regex:$k 2>/dev/null | python profiles/agg_by_addr.py gpurun_out/${T}_${k}_by_addr.json 4096 > /dev/null 2>&1
done
rm -f gpurun_out/*.ncu-rep
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests -m gpu -x -q -k "msm or edges or golden" > gpurun_out/${T}_sanitizer_memcheck.txt 2>&1; tail -4 gpurun_out/${T}_sanitizer_memcheck.txt
