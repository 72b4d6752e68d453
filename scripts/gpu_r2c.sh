#!/bin/bash
# round 2, session c (1 GPU): complex kernels + zprimme + complex reference driver on the GPU
mkdir -p gpurun_out
export PB200_DEBUG=1
timeout 1500 python -m pytest tests/test_zkernels_gpu.py tests/test_zprimme_gpu.py tests/test_driver_gpu.py -m gpu -q -x -s --timeout 900 > gpurun_out/pytest_z_r2c.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_z_r2c.log
grep -E "passed|failed|FAILED|Error|exit|gpu \[|C3 gpu" gpurun_out/pytest_z_r2c.log | head -40
tail -30 gpurun_out/pytest_z_r2c.log | cut -c1-300
