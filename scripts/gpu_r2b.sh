#!/bin/bash
# round 2, session b (1 GPU): SpMM after the long-row pass + layout timing; full GPU test suite; bench with the c5 block
mkdir -p gpurun_out
export PB200_DEBUG=1
for cfg in c2 c5; do
  echo "=== $cfg auto"
  timeout 300 python scripts/kernel_bench.py --reps 20 --config $cfg --only spmm 2>&1 | grep -v "^{" | tee -a gpurun_out/kernel_bench_spmm_r2b.txt
done
PB200_SPMM_V3=1 timeout 300 python scripts/kernel_bench.py --reps 20 --config c5 --only spmm 2>&1 | grep -v "^{" | tee -a gpurun_out/kernel_bench_spmm_r2b.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2b.log
tail -5 gpurun_out/pytest_gpu_r2b.log
echo "=== bench"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -5 gpurun_out/bench_r2b.err; cat gpurun_out/bench_r2b.json
