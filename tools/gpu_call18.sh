#!/bin/bash
# round-2 GPU call 18: final build -- full gpu test suite, the default bench line, ncu --set full of the product QR kernel
# inside a timed step (traffic + DMMA pipe share of the reworked kernel)
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q -x --durations=3) > gpurun_out/c18_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/c18_pytest.log | tail -2
timeout 900 python bench.py > gpurun_out/c18_bench.json 2> gpurun_out/c18_bench.err
grep "^{" gpurun_out/c18_bench.json | cut -c1-220; tail -2 gpurun_out/c18_bench.err
MPBP_PROFILER_RANGE=1 timeout 700 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_qr_ft --launch-skip 700 -c 1 -o /tmp/c18_qr -f python bench.py --steps 1 --warmup 4 --no-cpu --no-profile > gpurun_out/c18_ncu_qr.log 2>&1
ncu -i /tmp/c18_qr.ncu-rep --page raw --csv > gpurun_out/c18_qr_raw.csv 2>/dev/null
ncu -i /tmp/c18_qr.ncu-rep --page source --csv > gpurun_out/c18_qr_source.csv 2>/dev/null
ls -la gpurun_out/c18_*
