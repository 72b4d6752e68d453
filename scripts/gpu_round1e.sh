#!/bin/bash
# staged GPU check: every stage has its own short timeout so a hung kernel costs minutes, not the call
mkdir -p gpurun_out
export PB200_DEBUG=1
kt() {  # label, env assignments...
   local label=$1; shift
   env "$@" timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 45 --timeout-method thread \
      > gpurun_out/kt_$label.log 2>&1
   local rc=$?
   echo "kernel tests [$label] exit $rc: $(grep -E 'passed|failed' gpurun_out/kt_$label.log | tail -1)"
   grep -E "^FAILED|Timeout|primme_b200:" gpurun_out/kt_$label.log | head -5
   return $rc
}
CFG=""
if kt default; then CFG="";
else
   kt nospmm PB200_NO_TMA_SPMM=1 && CFG="PB200_NO_TMA_SPMM=1"
   if [ -z "$CFG" ]; then kt novw PB200_NO_TMA_VWXR=1 && CFG="PB200_NO_TMA_VWXR=1"; fi
   if [ -z "$CFG" ]; then kt nows PB200_NO_WS=1 && CFG="PB200_NO_WS=1"; fi
   if [ -z "$CFG" ]; then kt nows_nospmm PB200_NO_WS=1 PB200_NO_TMA_SPMM=1 && CFG="PB200_NO_WS=1 PB200_NO_TMA_SPMM=1"; fi
   if [ -z "$CFG" ]; then kt notma PB200_NO_TMA=1 && CFG="PB200_NO_TMA=1"; fi
   if [ -z "$CFG" ]; then echo "no working configuration"; exit 1; fi
fi
echo "=== working configuration: '${CFG:-default}'"
echo "=== all gpu tests"
env $CFG timeout 600 python -m pytest tests -m gpu -q --timeout 120 --timeout-method thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "PRIMME-B200|primme_b200:|passed|failed|FAILED|pytest exit" gpurun_out/pytest_gpu.log | head -30
echo "=== kernel bench c2"
env $CFG timeout 200 python scripts/kernel_bench.py --reps 10 2>&1 | head -7 | tee gpurun_out/kernel_bench_c2_v4.txt
echo "=== kernel bench c2, row-block spmm"
env $CFG PB200_NO_TMA_SPMM=1 timeout 200 python scripts/kernel_bench.py --reps 10 2>&1 | head -1 | tee gpurun_out/kernel_bench_c2_v4_spmm1.txt
echo "=== bench"
env $CFG timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err
tail -3 gpurun_out/bench_v4.err; cat gpurun_out/bench_v4.json
