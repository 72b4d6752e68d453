#!/bin/bash
mkdir -p gpurun_out
export PB200_DEBUG=1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "PRIMME-B200|primme_b200:|passed|failed|FAILED" gpurun_out/pytest_gpu.log | head -30
echo "=== bench TMA on"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_tma.json 2> gpurun_out/bench_tma.err; tail -3 gpurun_out/bench_tma.err; cat gpurun_out/bench_tma.json
echo "=== bench TMA off"; PB200_NO_TMA=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_notma.json 2> gpurun_out/bench_notma.err; tail -3 gpurun_out/bench_notma.err; cat gpurun_out/bench_notma.json
echo "=== bench only ortho TMA"; PB200_NO_TMA_VWXR=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_tma_ortho_only.json 2> gpurun_out/bench_o.err; tail -3 gpurun_out/bench_o.err; cat gpurun_out/bench_tma_ortho_only.json
