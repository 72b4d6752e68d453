#!/bin/bash
# round 2, session n (1 GPU): long-row threshold of the SpMM schedule on the power-law matrix (C5 per-rank shape),
# SpMM parity with a low threshold, windowed SpMM tests
for t in 2048 1024 512 256 128; do
  echo "=== C5 SpMM, PB200_SPMM_LONGROW=$t"
  PB200_SPMM_LONGROW=$t timeout 300 python scripts/kernel_bench.py --config c5 --only "spmm" 2>&1 | grep "^spmm"
done
echo "=== spmm parity, threshold 128"
PB200_SPMM_LONGROW=128 timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "spmm" -x 2>&1 | tail -4
echo "=== spmm parity, default"
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "spmm" -x 2>&1 | tail -4
