#!/bin/bash
# round-2 GPU call 13: A/B of the lean panel (dots broadcast through shared memory), QR parity, lane mode test + bench
mkdir -p gpurun_out
timeout 300 python tools/qr_variants.py tools/_variants/old.so tools/_variants/lean.so tools/_variants/lean2.so > gpurun_out/c13_qr_ab.log 2>&1
cat gpurun_out/c13_qr_ab.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -x -q -m gpu -k "qr or knobs or headline_node" > gpurun_out/c13_pytest.log 2>&1
tail -3 gpurun_out/c13_pytest.log
for L in 0 4 8; do
  timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --no-profile --set lanes=$L > gpurun_out/c13_bench_lanes$L.json 2> gpurun_out/c13_bench_lanes$L.err
  grep "^{" gpurun_out/c13_bench_lanes$L.json | cut -c1-200; tail -2 gpurun_out/c13_bench_lanes$L.err
done
