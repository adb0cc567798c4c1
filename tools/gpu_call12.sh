#!/bin/bash
# round-2 GPU call 12: A/B of the QR panel factorisation (shuffle-reduce vs DMMA Gram), QR parity tests on the new build
mkdir -p gpurun_out
timeout 300 python tools/qr_variants.py tools/_variants/old.so tools/_variants/new.so > gpurun_out/c12_qr_ab.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "qr or svd or tree or loopy" > gpurun_out/c12_pytest.log 2>&1
cat gpurun_out/c12_qr_ab.log; tail -5 gpurun_out/c12_pytest.log
