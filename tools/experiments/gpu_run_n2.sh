#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r1s2_final_bench_n2.json 2> gpurun_out/r1s2_final_bench_n2.err; tail -1 gpurun_out/r1s2_final_bench_n2.json | cut -c1-400
python -m pytest tests -m gpu -x -q -k "multi_device" 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_nccl_check.py > gpurun_out/r1s2_final_dist_nccl_check_2gpu.txt 2>&1; tail -3 gpurun_out/r1s2_final_dist_nccl_check_2gpu.txt
