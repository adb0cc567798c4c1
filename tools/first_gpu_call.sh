#!/bin/bash
# First GPU call of the next round: run the two kernel prototypes next to the product kernels, then the parity suite.
# Build the binaries here first (nvcc cross-compiles without a GPU), then:
#   gpurun --timeout 400 -- 'bash tools/first_gpu_call.sh > gpurun_out/first_call.log 2>&1; tail -30 gpurun_out/first_call.log'
set -x
for shape in "1600 400 148" "4000 400 148" "3200 400 296" "700 100 148" "257 17 8"; do
  timeout 60 ./tools/qr_chain2 $shape
done
timeout 60 ./tools/kron_carry2 32
timeout 60 python tools/qr_variants.py matrixproductbp.jl_b200/libmpbp_b200.so
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
