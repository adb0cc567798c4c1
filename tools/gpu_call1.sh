#!/bin/bash
# round-2 GPU call 1: prototypes, parity suite, short benches (baseline + outlier split variants)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
for shape in "1600 400 148" "4000 400 148" "8800 400 148" "700 100 148" "257 17 8"; do
  timeout 120 ./tools/qr_chain2 $shape
done > gpurun_out/c1_chain2.log 2>&1
timeout 60 ./tools/kron_carry2 32 > gpurun_out/c1_kc2.log 2>&1
(time timeout 1500 python -m pytest tests -m gpu -x -q -s --durations=15) > gpurun_out/c1_pytest.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu > gpurun_out/c1_bench_base.json 2> gpurun_out/c1_bench_base.err
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu --no-profile --set outlier_split=2.5 > gpurun_out/c1_bench_os25.json 2> gpurun_out/c1_bench_os25.err
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu --no-profile --set outlier_split=1.7 > gpurun_out/c1_bench_os17.json 2> gpurun_out/c1_bench_os17.err
tail -5 gpurun_out/c1_chain2.log gpurun_out/c1_pytest.log
cat gpurun_out/c1_bench_base.json | cut -c1-600
