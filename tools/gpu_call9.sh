#!/bin/bash
# round-2 GPU call 9: ncu evidence restricted to the TIMED step (cudaProfilerStart/Stop range): launch list, --set full of the
# product QR (H=64) and SVD kernels mid-step; stream-count A/B; config 4 at reduced T
mkdir -p gpurun_out
timeout 500 python bench.py --steps 2 --warmup 4 --no-cpu --no-profile --set nstreams=2 > gpurun_out/c9_bench_s2.json 2> gpurun_out/c9_bench_s2.err
timeout 500 python bench.py --steps 2 --warmup 4 --no-cpu --no-profile --nodes-per-gpu 320 > gpurun_out/c9_bench_n320.json 2> gpurun_out/c9_bench_n320.err
timeout 400 python bench.py --config 4 --T 20 --steps 1 --warmup 3 --no-cpu --no-profile > gpurun_out/c9_cfg4_T20.json 2> gpurun_out/c9_cfg4_T20.err
MPBP_PROFILER_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 4 --no-cpu --no-profile > gpurun_out/c9_ncu_launch.log 2>&1
MPBP_PROFILER_RANGE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_qr_ft --launch-skip 700 -c 1 -o /tmp/c9_qr -f python bench.py --steps 1 --warmup 4 --no-cpu --no-profile > gpurun_out/c9_ncu_qr.log 2>&1
ncu -i /tmp/c9_qr.ncu-rep --page raw --csv > gpurun_out/r02_qr_ft_raw.csv 2>/dev/null
ncu -i /tmp/c9_qr.ncu-rep --page source --csv > gpurun_out/r02_qr_ft_source.csv 2>/dev/null
MPBP_PROFILER_RANGE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_jacobi_project --launch-skip 700 -c 1 -o /tmp/c9_svd -f python bench.py --steps 1 --warmup 4 --no-cpu --no-profile > gpurun_out/c9_ncu_svd.log 2>&1
ncu -i /tmp/c9_svd.ncu-rep --page raw --csv > gpurun_out/r02_svd_raw.csv 2>/dev/null
ncu -i /tmp/c9_svd.ncu-rep --page source --csv > gpurun_out/r02_svd_source.csv 2>/dev/null
for f in c9_bench_s2 c9_bench_n320 c9_cfg4_T20; do cut -c1-260 gpurun_out/$f.json; tail -2 gpurun_out/$f.err; done
ls -la gpurun_out/r02_* gpurun_out/c9_*
