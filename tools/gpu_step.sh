#!/bin/bash
# One measurement step on a B200: parity suite, A/B of the hash_g2 knob with tools/kbench.py, launch list of the verify step, one bench line.
# Usage: bash tools/gpu_step.sh <tag> [hash algos, default "0 2"]
T=${1:-step}
ALGOS=${2:-"0 2"}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.txt
tail -4 gpurun_out/${T}_pytest_gpu.txt
for ha in $ALGOS; do
  HASH_ALGO=$ha python tools/kbench.py ${T}_h${ha} verify 2>&1 | tail -1 | cut -c1-900
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_verify.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-combine --no-others --no-sharded > /dev/null 2>&1
python - <<'P' ${T}
import csv, sys, collections
t = sys.argv[1]
rows = list(csv.reader(l for l in open(f"gpurun_out/{t}_launches_verify.csv") if l.startswith('"')))
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
    except Exception: pass
for k, v in agg.items(): print(f"{k:40s} n={len(v):3d} last={v[-1]/1e6:9.3f} ms")
P
python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; cut -c1-400 gpurun_out/${T}_bench_n1.json; tail -2 gpurun_out/${T}_bench_n1.err
