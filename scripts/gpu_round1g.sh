#!/bin/bash
mkdir -p gpurun_out
export PB200_DEBUG=1
kt() {
   local label=$1; shift
   env "$@" timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 45 --timeout-method thread \
      > gpurun_out/kt_$label.log 2>&1
   local rc=$?
   echo "kernel tests [$label] exit $rc: $(grep -E 'passed|failed' gpurun_out/kt_$label.log | tail -1)"
   grep -E "^FAILED|Timeout|primme_b200:" gpurun_out/kt_$label.log | head -5
   return $rc
}
kt default || exit 1
kt optin PB200_WS=1 PB200_NARROW=1 PB200_CAND_TMA=1
kt staged PB200_NO_INLINE_COEF=1
echo "=== all gpu tests"
timeout 600 python -m pytest tests -m gpu -q --timeout 120 --timeout-method thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "primme_b200:|passed|failed|FAILED|pytest exit" gpurun_out/pytest_gpu.log | head -30
echo "=== kernel bench c2"
timeout 200 python scripts/kernel_bench.py --reps 10 2>&1 | grep -v "^{" | tee gpurun_out/kernel_bench_c2_v6.txt
echo "=== kernel bench c2, coefficients staged by memcpy"
PB200_NO_INLINE_COEF=1 timeout 200 python scripts/kernel_bench.py --reps 10 2>&1 | grep -v "^{" | tee gpurun_out/kernel_bench_c2_v6_staged.txt
echo "=== bench"
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err
grep "phases" gpurun_out/bench_v6.err | tail -3; cat gpurun_out/bench_v6.json
echo "=== bench, memcpy+sync panels"
PB200_NO_POLL=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v6_nopoll.json 2> gpurun_out/bench_v6_np.err
grep "phases" gpurun_out/bench_v6_np.err | tail -1; cut -c1-400 gpurun_out/bench_v6_nopoll.json
