#!/bin/bash
# round 2, session aa (1 GPU): the driver's round-end sequence: smoke(), then the whole GPU suite
mkdir -p gpurun_out
timeout 900 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_r2aa.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2aa.log
grep -E "passed|failed|FAILED|ERROR|exit" gpurun_out/pytest_gpu_r2aa.log | head -20
