#!/bin/bash
# round-2 GPU call 7: parity with svd_mode 2 as default + operand staging in the Kronecker carry; profiled bench
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q --durations=5) > gpurun_out/c7_pytest.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu --set outlier_split=1.7 > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu --no-profile --set outlier_split=1.7 --set nstreams=8 > gpurun_out/c7_bench_s8.json 2> gpurun_out/c7_bench_s8.err
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu --no-profile --set outlier_split=1.7 --set group_mode=1 --set nstreams=6 > gpurun_out/c7_bench_gm1.json 2> gpurun_out/c7_bench_gm1.err
grep -E "passed|failed" gpurun_out/c7_pytest.log | tail -2
for f in c7_bench c7_bench_s8 c7_bench_gm1; do cut -c1-200 gpurun_out/$f.json; done
