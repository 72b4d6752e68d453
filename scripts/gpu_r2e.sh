#!/bin/bash
# round 2, session e (1 GPU): whole GPU suite with exact-count assertions; bench with and without alternating sweeps
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 900 > gpurun_out/pytest_gpu_r2e.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2e.log
grep -E "passed|failed|FAILED|exit" gpurun_out/pytest_gpu_r2e.log | head -30
grep -E "counts .* reference" gpurun_out/pytest_gpu_r2e.log | head -40
echo "=== bench alternate"
PB200_DEBUG=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2e_alt.json 2> gpurun_out/bench_r2e_alt.err; tail -3 gpurun_out/bench_r2e_alt.err; cut -c1-1800 gpurun_out/bench_r2e_alt.json
echo "=== bench no alternate"
PB200_NO_ALTERNATE=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2e_noalt.json 2> gpurun_out/bench_r2e_noalt.err; cut -c1-1800 gpurun_out/bench_r2e_noalt.json
