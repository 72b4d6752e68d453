#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel bench c2"; timeout 300 python scripts/kernel_bench.py --reps 10 2>&1 | tee gpurun_out/kernel_bench_c2.txt
echo "=== kernel bench c2 no TMA"; PB200_NO_TMA=1 timeout 300 python scripts/kernel_bench.py --reps 10 2>&1 | tee gpurun_out/kernel_bench_c2_notma.txt
echo "=== ncu full (kernel bench, 1 rep)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tma_kernel|spmm_kernel" -c 14 -f -o gpurun_out/prof_r01_kernels python scripts/kernel_bench.py --reps 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
echo "=== ncu launch list of one bench solve"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
tail -2 gpurun_out/launches_r01.csv | cut -c1-300
ls -la gpurun_out
