#!/bin/bash
# round 2, session m (1 GPU): windowed SpMM (v4): parity, timing at the C2 shape against the two gather layouts
echo "=== spmm parity"
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "spmm" -x 2>&1 | tail -8
for v in 2 3 0 1; do
  echo "=== C2 SpMM, PB200_SPMM_V3=$v"
  PB200_DEBUG=1 PB200_SPMM_V3=$v timeout 300 python scripts/kernel_bench.py --config c2 --only "spmm" 2>&1 | grep -i "primme_b200: SpMM\|^spmm"
done
