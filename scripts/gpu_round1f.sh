#!/bin/bash
# staged GPU check (every stage has its own short timeout)
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
kt default || exit 1
kt optin PB200_WS=1 PB200_NARROW=1 PB200_CAND_TMA=1
kt v1 PB200_NO_TMA=1
echo "=== all gpu tests"
timeout 600 python -m pytest tests -m gpu -q --timeout 120 --timeout-method thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "PRIMME-B200|primme_b200:|passed|failed|FAILED|pytest exit" gpurun_out/pytest_gpu.log | head -30
echo "=== kernel bench c2 (CUDA events around each launch)"
timeout 200 python scripts/kernel_bench.py --reps 10 2>&1 | head -6 | tee gpurun_out/kernel_bench_c2_v5.txt
echo "=== kernel bench c2, 1 CTA/SM ortho plan"
PB200_ORTHO_1CTA=1 timeout 200 python scripts/kernel_bench.py --reps 10 2>&1 | head -4 | tee gpurun_out/kernel_bench_c2_v5_1cta.txt
echo "=== ncu launch list of the kernel bench (kernel-only durations)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_kb_r01.csv \
   python scripts/kernel_bench.py --reps 3 > gpurun_out/kb_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_kb_r01.csv | head -30
echo "=== bench"
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err
tail -3 gpurun_out/bench_v5.err; cat gpurun_out/bench_v5.json
echo "=== ncu full on the kernel bench"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmm_tma|ortho_sweep_tma|vwxr_kernel|vwxr_wide" -c 16 -f \
   -o gpurun_out/prof_r01_v5 python scripts/kernel_bench.py --reps 1 > gpurun_out/ncu_full_v5.log 2>&1
tail -2 gpurun_out/ncu_full_v5.log
ncu -i gpurun_out/prof_r01_v5.ncu-rep --page raw --csv > gpurun_out/prof_r01_v5_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
