#!/bin/bash
# round 2, session d (2 GPUs): multi-GPU tests incl. the row-partitioned SVD operator; complex reference driver; C3
mkdir -p gpurun_out
export PB200_DEBUG=1
timeout 1200 python -m pytest tests/test_multi_gpu.py tests/test_driver_gpu.py tests/test_zprimme_gpu.py -m gpu -q -s --timeout 900 > gpurun_out/pytest_r2d.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r2d.log
grep -E "passed|failed|FAILED|Error|exit|C3 gpu" gpurun_out/pytest_r2d.log | head -40
tail -25 gpurun_out/pytest_r2d.log | cut -c1-400
