#!/bin/bash
# round-2 GPU call 16: A/B of the branch-free scalar path of the panel chain (lean4) against lean3; QR parity on lean4
mkdir -p gpurun_out
timeout 300 python tools/qr_variants.py tools/_variants/lean3.so tools/_variants/lean4.so > gpurun_out/c16_qr_ab.log 2>&1
cat gpurun_out/c16_qr_ab.log
