#!/bin/bash
# round-2 GPU call 4: R prefetch in the QR, hoisted DMMA Kronecker carry, hub lane at high stream priority, SVD variants
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q --durations=5) > gpurun_out/c4_pytest.log 2>&1
timeout 300 python tools/qr_bench.py > gpurun_out/c4_qr_bench.log 2>&1
MPBP_SVD_PHASES=1 MPBP_SVD_MODE=0 timeout 300 python tools/svd_bench.py > gpurun_out/c4_svd_bench_m0.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu --set outlier_split=1.7 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu --no-profile --set outlier_split=1.7 --set hub_lane=0 > gpurun_out/c4_bench_hub0.json 2> gpurun_out/c4_bench_hub0.err
grep -E "passed|failed" gpurun_out/c4_pytest.log | tail -2
cat gpurun_out/c4_qr_bench.log gpurun_out/c4_svd_bench_m0.log
for f in c4_bench c4_bench_hub0; do cut -c1-220 gpurun_out/$f.json; done
