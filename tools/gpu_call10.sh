#!/bin/bash
# round-2 GPU call 10: parity + profiled bench after the DMMA carry projection in the SVD kernel
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q --durations=3) > gpurun_out/c10_pytest.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err
grep -E "passed|failed" gpurun_out/c10_pytest.log | tail -2
cut -c1-200 gpurun_out/c10_bench.json
