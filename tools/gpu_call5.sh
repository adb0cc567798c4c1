#!/bin/bash
# round-2 GPU call 5: memcheck of the bipartite overflow repro, parity suite, Cholesky-QR SVD variant, launch list + kron ncu
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/repro_bipartite.py > gpurun_out/c5_memcheck.log 2>&1
(time timeout 900 python -m pytest tests -m gpu -q --durations=5) > gpurun_out/c5_pytest.log 2>&1
MPBP_SVD_PHASES=1 MPBP_SVD_MODE=2 timeout 300 python tools/svd_bench.py > gpurun_out/c5_svd_bench_m2.log 2>&1
MPBP_SVD_PHASES=1 MPBP_SVD_MODE=1 timeout 300 python tools/svd_bench.py > gpurun_out/c5_svd_bench_m1.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu --set outlier_split=1.7 > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
timeout 600 python bench.py --steps 2 --warmup 4 --no-cpu --no-profile --set outlier_split=1.7 --set svd_mode=2 > gpurun_out/c5_bench_svd2.json 2> gpurun_out/c5_bench_svd2.err
# launch list of one full-bond step (serialised by ncu: shares, not absolute times)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 51500 -c 12500 --csv --log-file gpurun_out/c5_launches.csv python bench.py --steps 1 --warmup 4 --no-cpu --no-profile --set outlier_split=1.7 > gpurun_out/c5_ncu_launch.log 2>&1
# one --set full capture of the DMMA Kronecker carry inside a full-bond step; CSV pages only (the .ncu-rep is 25 MB)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_kron_carry_mma --launch-skip 2500 -c 1 -o /tmp/c5_kron -f python bench.py --steps 1 --warmup 4 --no-cpu --no-profile > gpurun_out/c5_ncu_kron.log 2>&1
ncu -i /tmp/c5_kron.ncu-rep --page raw --csv > gpurun_out/c5_kron_raw.csv 2>/dev/null
ncu -i /tmp/c5_kron.ncu-rep --page source --csv > gpurun_out/c5_kron_source.csv 2>/dev/null
tail -5 gpurun_out/c5_memcheck.log
grep -E "passed|failed" gpurun_out/c5_pytest.log | tail -2
cat gpurun_out/c5_svd_bench_m2.log
for f in c5_bench c5_bench_svd2; do cut -c1-200 gpurun_out/$f.json; done
ls -la gpurun_out/c5_*
