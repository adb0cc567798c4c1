#!/bin/bash
# round-2 GPU call 20: final library -- the default bench line (with the CPU arm), ncu launch list of the periodic path,
# ncu --set full of the product QR kernel inside a timed step (refreshes profiles/qr_traffic.json for the reworked kernel)
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/c20_bench.json 2> gpurun_out/c20_bench.err
grep "^{" gpurun_out/c20_bench.json | cut -c1-200; tail -2 gpurun_out/c20_bench.err
timeout 120 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_periodic -c 60 --csv --log-file gpurun_out/c20_periodic_launches.csv python -m pytest tests/test_gpu_periodic.py -q -x -k "tree or infinite" > gpurun_out/c20_periodic_ncu.log 2>&1
tail -3 gpurun_out/c20_periodic_ncu.log; wc -l gpurun_out/c20_periodic_launches.csv
MPBP_PROFILER_RANGE=1 timeout 420 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_qr_ft --launch-skip 700 -c 1 -o /tmp/c20_qr -f python bench.py --steps 1 --warmup 4 --no-cpu --no-profile > gpurun_out/c20_ncu_qr.log 2>&1
tail -3 gpurun_out/c20_ncu_qr.log
ncu -i /tmp/c20_qr.ncu-rep --page raw --csv > gpurun_out/c20_qr_raw.csv 2>/dev/null
ncu -i /tmp/c20_qr.ncu-rep --page source --csv > gpurun_out/c20_qr_source.csv 2>/dev/null
ls -la gpurun_out/c20_*
