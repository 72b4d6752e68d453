#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel bench c2"
PB200_DEBUG=1 timeout 200 python scripts/kernel_bench.py --reps 10 2>&1 | grep -v "^{" | tee gpurun_out/kernel_bench_c2_v8.txt
