#!/bin/bash
# round-2 GPU call 8: in-place Cholesky-QR inverse (SVD), ncu --set full of the product QR and Kronecker kernels inside a
# full-bond step (CSV pages only), first runs of configs 2 and 5
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q -x --durations=3) > gpurun_out/c8_pytest.log 2>&1
MPBP_SVD_PHASES=1 timeout 300 python tools/svd_bench.py > gpurun_out/c8_svd_bench.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu --no-profile --set outlier_split=1.7 > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err
timeout 900 python bench.py --config 2 --steps 2 --warmup 4 --no-cpu > gpurun_out/c8_cfg2.json 2> gpurun_out/c8_cfg2.err
timeout 900 python bench.py --config 5 --steps 2 --warmup 4 --no-cpu --nodes-per-gpu 512 > gpurun_out/c8_cfg5.json 2> gpurun_out/c8_cfg5.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_qr_ft --launch-skip 9000 -c 1 -o /tmp/c8_qr -f python bench.py --steps 1 --warmup 4 --no-cpu --no-profile --set outlier_split=1.7 > gpurun_out/c8_ncu_qr.log 2>&1
ncu -i /tmp/c8_qr.ncu-rep --page raw --csv > gpurun_out/c8_qr_raw.csv 2>/dev/null
ncu -i /tmp/c8_qr.ncu-rep --page source --csv > gpurun_out/c8_qr_source.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_kron_carry_mma --launch-skip 9000 -c 1 -o /tmp/c8_kron -f python bench.py --steps 1 --warmup 4 --no-cpu --no-profile --set outlier_split=1.7 > gpurun_out/c8_ncu_kron.log 2>&1
ncu -i /tmp/c8_kron.ncu-rep --page raw --csv > gpurun_out/c8_kron_raw.csv 2>/dev/null
ncu -i /tmp/c8_kron.ncu-rep --page source --csv > gpurun_out/c8_kron_source.csv 2>/dev/null
grep -E "passed|failed" gpurun_out/c8_pytest.log | tail -2
cat gpurun_out/c8_svd_bench.log
for f in c8_bench c8_cfg2 c8_cfg5; do cut -c1-260 gpurun_out/$f.json; tail -2 gpurun_out/$f.err; done
ls -la gpurun_out/c8_*
