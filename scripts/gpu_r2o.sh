#!/bin/bash
# round 2, session o (1 GPU): rows of the gather buffer shared by adjacent lanes (one L1 wavefront per gathered
# row): parity (real, complex, windowed) and the power-law timing
echo "=== spmm parity"
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "spmm" 2>&1 | tail -4
for t in 2048 1024 512; do
  echo "=== C5 SpMM, PB200_SPMM_LONGROW=$t"
  PB200_SPMM_LONGROW=$t timeout 300 python scripts/kernel_bench.py --config c5 --only "spmm" 2>&1 | grep "^spmm"
done
echo "=== C2 SpMM forced row-major"
PB200_SPMM_V3=1 timeout 300 python scripts/kernel_bench.py --config c2 --only "spmm" 2>&1 | grep "^spmm"
