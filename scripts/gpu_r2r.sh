#!/bin/bash
# round 2, session r (1 GPU): lanes per row of the row-major SpMM on the power-law matrix; spmm parity
for lpr in 1 2 4 8; do
  echo "=== C5 SpMM, PB200_SPMM_LPR=$lpr"
  PB200_SPMM_LPR=$lpr timeout 300 python scripts/kernel_bench.py --config c5 --only "spmm" 2>&1 | grep "^spmm"
done
echo "=== spmm parity"
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_zkernels_gpu.py -m gpu -q --timeout 120 -k "spmm" 2>&1 | tail -3
