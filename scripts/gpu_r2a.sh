#!/bin/bash
# round 2, session a: SpMM v3 parity + v2/v3 timing at the C2 and C5 shapes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "spmm" > gpurun_out/pytest_spmm.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_spmm.log
tail -5 gpurun_out/pytest_spmm.log
for cfg in c2 c5; do
  for v in 1 0; do
    echo "=== $cfg v3=$v"
    PB200_SPMM_V3=$v timeout 300 python scripts/kernel_bench.py --reps 20 --config $cfg --only spmm 2>&1 | grep -v "^{" | tee -a gpurun_out/kernel_bench_spmm_r2a.txt
  done
done
