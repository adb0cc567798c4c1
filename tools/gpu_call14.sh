#!/bin/bash
# round-2 GPU call 14: A/B of panel variants (lean2: T one row per lane; lean3: + shared-memory reduction, staged R rows)
mkdir -p gpurun_out
timeout 300 python tools/qr_variants.py tools/_variants/lean2.so tools/_variants/lean3.so > gpurun_out/c14_qr_ab.log 2>&1
cat gpurun_out/c14_qr_ab.log
