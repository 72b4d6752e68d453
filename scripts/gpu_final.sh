#!/bin/bash
# Round-end verification on one GPU: smoke, the GPU test suite, the default bench line (with the
# CPU baseline) and the reference arm.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error|exit" gpurun_out/pytest_gpu.log | head -10
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"; cat gpurun_out/bench_final.json | cut -c1-2500
timeout 600 python bench.py --impl reference > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err; echo "reference arm exit $?"; cat gpurun_out/bench_final_reference.json | cut -c1-600
