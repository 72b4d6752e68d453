#!/bin/bash
mkdir -p gpurun_out
export PB200_DEBUG=1
echo "=== kernels"; timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q 2>&1 | tail -15
echo "=== failing solver tests (TMA on)"; timeout 600 python -m pytest tests/test_solver_gpu.py tests/test_reference_golden_gpu.py -m gpu -q -x 2>&1 | grep -E "PRIMME-B200|passed|failed|FAILED" | head -20
echo "=== same, TMA vwxr off"; PB200_NO_TMA_VWXR=1 timeout 600 python -m pytest tests/test_solver_gpu.py tests/test_reference_golden_gpu.py -m gpu -q 2>&1 | grep -E "PRIMME-B200|passed|failed|FAILED" | head -20
echo "=== same, all TMA off"; PB200_NO_TMA=1 timeout 600 python -m pytest tests/test_solver_gpu.py tests/test_reference_golden_gpu.py -m gpu -q 2>&1 | grep -E "PRIMME-B200|passed|failed|FAILED" | head -20
echo "=== bench TMA on"; timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_tma.json 2> gpurun_out/bench_tma.err; tail -3 gpurun_out/bench_tma.err; cat gpurun_out/bench_tma.json
echo "=== bench TMA off"; PB200_NO_TMA=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_notma.json 2> gpurun_out/bench_notma.err; tail -3 gpurun_out/bench_notma.err; cat gpurun_out/bench_notma.json
