#!/bin/bash
# round 2, run 11: bench.py at N GPUs under torchrun (weak-scaling headline + sharded strong-scaling records)
N=${1:-8}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_final_bench_n$N.json 2> gpurun_out/r2_final_bench_n$N.err
echo rc=$?
tail -3 gpurun_out/r2_final_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_final_bench_n$N.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'])
    for k in ('combine_sharded','commit_eval_sharded','multi_device_ctx'):
        print(k, json.dumps(d[k])[:600])
    print('combine', d['combine']['value'], d['combine']['e2e']['value'], 'decrypt', d['decrypt']['value'], d['decrypt']['e2e']['value'], 'eval', d['commit_eval']['value'])
except Exception as e:
    print('parse failed', e)
PY
