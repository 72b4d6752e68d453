#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_tma.json 2> gpurun_out/bench_tma.err; echo "bench exit $?"
cat gpurun_out/bench_tma.json; tail -5 gpurun_out/bench_tma.err
PB200_NO_TMA=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_notma.json 2> gpurun_out/bench_notma.err
cat gpurun_out/bench_notma.json
