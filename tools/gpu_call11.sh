#!/bin/bash
# round-2 GPU call 11 (2 GPUs): NCCL parity of the node-partitioned run, then the default bench at N = 2
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/c11_dist_check.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 4 > gpurun_out/c11_bench_2gpu.json 2> gpurun_out/c11_bench_2gpu.err
tail -2 gpurun_out/c11_dist_check.log
grep "^{" gpurun_out/c11_bench_2gpu.json | cut -c1-250
tail -3 gpurun_out/c11_bench_2gpu.err
