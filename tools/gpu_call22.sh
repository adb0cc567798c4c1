#!/bin/bash
# round-2 GPU call 22 (2 GPUs): the final library on the multi-rank path -- NCCL bench line at N = 2 (short) + the new periodic test
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_periodic.py -q -x -k "truncthresh0" > gpurun_out/c22_periodic.log 2>&1; tail -2 gpurun_out/c22_periodic.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 3 --no-profile > gpurun_out/c22_bench_2gpu.json 2> gpurun_out/c22_bench_2gpu.err
grep "^{" gpurun_out/c22_bench_2gpu.json | cut -c1-400; tail -3 gpurun_out/c22_bench_2gpu.err
