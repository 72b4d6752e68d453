#!/bin/bash
# round 2, session p (1 GPU): nnz-balanced consumer of the row-major SpMM: parity, power-law timing
echo "=== spmm parity"
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "spmm" 2>&1 | tail -4
for bal in 1 0; do
  echo "=== C5 SpMM, PB200_SPMM_BAL=$bal"
  PB200_SPMM_BAL=$bal timeout 300 python scripts/kernel_bench.py --config c5 --only "spmm" 2>&1 | grep "^spmm"
done
echo "=== C5 SpMM, default heuristic"
timeout 300 python scripts/kernel_bench.py --config c5 --only "spmm" 2>&1 | grep "^spmm"
