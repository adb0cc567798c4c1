#!/bin/bash
# round-2 GPU call 15: lean3 panel adopted -- full gpu test suite, ncu --set full of the isolated H=64 QR launch, bench
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q -x --durations=3) > gpurun_out/c15_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/c15_pytest.log | tail -2
timeout 300 python tools/qr_one64.py > gpurun_out/c15_qr_one.log 2>&1; cat gpurun_out/c15_qr_one.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_test_qr_ft --launch-skip 1 -c 1 -o /tmp/c15_qr -f python tools/qr_one64.py > gpurun_out/c15_ncu_qr.log 2>&1
ncu -i /tmp/c15_qr.ncu-rep --page raw --csv > gpurun_out/c15_qr_raw.csv 2>/dev/null
ncu -i /tmp/c15_qr.ncu-rep --page source --csv > gpurun_out/c15_qr_source.csv 2>/dev/null
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu > gpurun_out/c15_bench.json 2> gpurun_out/c15_bench.err
grep "^{" gpurun_out/c15_bench.json | cut -c1-200; tail -2 gpurun_out/c15_bench.err
ls -la gpurun_out/c15_*
