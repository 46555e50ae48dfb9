#!/bin/bash
# round 2, run 5: bench.py at N=2 under torchrun (sharded strong-scaling records, two-device ctx check)
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
tail -5 gpurun_out/r2e_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2e_bench_n2.json').read().strip().splitlines()[-1])
    for k in ('value','e2e','combine_sharded','commit_eval_sharded','multi_device_ctx'):
        print(k, json.dumps(d[k])[:700])
    print('combine', d['combine']['value'], d['combine']['e2e']['value'], 'decrypt', d['decrypt']['value'], d['decrypt']['e2e']['value'])
except Exception as e:
    print('parse failed', e)
PY
