#!/bin/bash
mkdir -p gpurun_out
export PB200_DEBUG=1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "PRIMME-B200|primme_b200:|passed|failed|FAILED" gpurun_out/pytest_gpu.log | head -30
echo "=== kernel bench c2"; timeout 300 python scripts/kernel_bench.py --reps 10 2>&1 | head -7 | tee gpurun_out/kernel_bench_c2_v3.txt
echo "=== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; tail -3 gpurun_out/bench_v3.err; cat gpurun_out/bench_v3.json
echo "=== bench no poll"; PB200_NO_POLL=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_v3_nopoll.json 2> gpurun_out/bench_np.err; tail -3 gpurun_out/bench_np.err; cat gpurun_out/bench_v3_nopoll.json
