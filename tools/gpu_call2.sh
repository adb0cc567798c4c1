#!/bin/bash
# round-2 GPU call 2: un-squared block SVD + DMMA Kronecker carry: parity suite, SVD micro-benchmark, profiled bench
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -s --durations=10) > gpurun_out/c2_pytest.log 2>&1
MPBP_SVD_PHASES=1 timeout 300 python tools/svd_bench.py > gpurun_out/c2_svd_bench.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu --set outlier_split=1.7 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu --no-profile --set outlier_split=1.7 --set kron_mma=0 > gpurun_out/c2_bench_kc0.json 2> gpurun_out/c2_bench_kc0.err
grep -E "passed|failed" gpurun_out/c2_pytest.log | tail -3
cat gpurun_out/c2_svd_bench.log
cut -c1-300 gpurun_out/c2_bench.json
