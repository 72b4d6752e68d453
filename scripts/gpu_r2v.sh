#!/bin/bash
# round 2, session v (1 GPU): full GPU suite after the device dlarnv
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_r2v.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2v.log
grep -E "passed|failed|FAILED|ERROR|exit" gpurun_out/pytest_gpu_r2v.log | head -30
