#!/bin/bash
# compute-sanitizer passes over the kernel parity tests (memcheck + racecheck on the TMA/DMMA kernels)
mkdir -p gpurun_out
SEL='test_vwxr or test_ortho_sweep'
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file gpurun_out/sanitizer_memcheck.log \
   python -m pytest tests/test_kernels_gpu.py -q -x -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_memcheck_pytest.log 2>&1
echo "memcheck exit $?"; tail -3 gpurun_out/sanitizer_memcheck_pytest.log; grep -c "Invalid\|Error" gpurun_out/sanitizer_memcheck.log; tail -3 gpurun_out/sanitizer_memcheck.log
timeout 700 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 3 --log-file gpurun_out/sanitizer_racecheck.log \
   python -m pytest tests/test_kernels_gpu.py -q -x -k "test_vwxr and (candp or restart or cand) or (test_ortho_sweep and 28)" -p no:cacheprovider > gpurun_out/sanitizer_racecheck_pytest.log 2>&1
echo "racecheck exit $?"; tail -3 gpurun_out/sanitizer_racecheck_pytest.log; grep -c "Race\|hazard" gpurun_out/sanitizer_racecheck.log; tail -5 gpurun_out/sanitizer_racecheck.log
