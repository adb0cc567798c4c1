#!/bin/bash
# round-2 GPU call 3: parity after the SVD/Householder/bulk-copy changes, hub lane A/B, ncu captures of the three kernel families
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q -x --durations=5) > gpurun_out/c3_pytest.log 2>&1
MPBP_SVD_PHASES=1 timeout 300 python tools/svd_bench.py > gpurun_out/c3_svd_bench.log 2>&1
timeout 300 python tools/qr_bench.py > gpurun_out/c3_qr_bench.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu --set outlier_split=1.7 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu --no-profile --set outlier_split=1.7 --set hub_lane=0 > gpurun_out/c3_bench_hub0.json 2> gpurun_out/c3_bench_hub0.err
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu --no-profile --set outlier_split=1.7 --nodes-per-gpu 384 > gpurun_out/c3_bench_n384.json 2> gpurun_out/c3_bench_n384.err
# ncu: one launch of each family inside a full-bond step (after the plain runs above exited)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_kron_carry_mma --launch-skip 3000 -c 1 -o gpurun_out/c3_ncu_kron -f python bench.py --steps 1 --warmup 4 --no-cpu --no-profile > gpurun_out/c3_ncu_kron.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_jacobi_project --launch-skip 3000 -c 1 -o gpurun_out/c3_ncu_svd -f python bench.py --steps 1 --warmup 4 --no-cpu --no-profile > gpurun_out/c3_ncu_svd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_qr_ft --launch-skip 3000 -c 1 -o gpurun_out/c3_ncu_qr -f python bench.py --steps 1 --warmup 4 --no-cpu --no-profile > gpurun_out/c3_ncu_qr.log 2>&1
grep -E "passed|failed" gpurun_out/c3_pytest.log | tail -2
cat gpurun_out/c3_svd_bench.log
for f in c3_bench c3_bench_hub0 c3_bench_n384; do cut -c1-220 gpurun_out/$f.json; done
ls -la gpurun_out/*.ncu-rep
