#!/bin/bash
# round-2 GPU call 6: parity after the generic-path fix, bulk TSQR split, Pyy staging, SVD modes; A/B of grouping modes
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q --durations=5) > gpurun_out/c6_pytest.log 2>&1
MPBP_SVD_PHASES=1 MPBP_SVD_MODE=2 timeout 300 python tools/svd_bench.py > gpurun_out/c6_svd_bench_m2.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu --set outlier_split=1.7 > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu --no-profile --set outlier_split=1.7 --set group_mode=1 --set nstreams=8 > gpurun_out/c6_bench_gm1.json 2> gpurun_out/c6_bench_gm1.err
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu --no-profile --set outlier_split=1.7 --set svd_mode=2 > gpurun_out/c6_bench_svd2.json 2> gpurun_out/c6_bench_svd2.err
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu --no-profile --set outlier_split=1.7 --set bulk_split=1 > gpurun_out/c6_bench_nosplit.json 2> gpurun_out/c6_bench_nosplit.err
grep -E "passed|failed" gpurun_out/c6_pytest.log | tail -2
cat gpurun_out/c6_svd_bench_m2.log
for f in c6_bench c6_bench_gm1 c6_bench_svd2 c6_bench_nosplit; do cut -c1-200 gpurun_out/$f.json; done
