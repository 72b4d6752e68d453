#!/bin/bash
# One-GPU check: GPU parity tests, per-kernel microbenchmark at the C2 shapes, bench line.
# usage (under gpurun): bash scripts/gpu_check.sh TAG [extra bench args]
TAG=${1:-x}; shift
mkdir -p gpurun_out
export PB200_DEBUG=1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error|exit" gpurun_out/pytest_gpu.log | head -20
echo "=== kernel bench c2"
timeout 200 python scripts/kernel_bench.py --reps 10 2>&1 | grep -v "^{" | tee gpurun_out/kernel_bench_c2_$TAG.txt
echo "=== bench"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
