#!/bin/bash
# round-1 profile of the v9 kernels: (1) ncu --set full of one launch of each hot kernel at the C2
# shapes, (2) ncu launch list of the first 3500 launches of one bench solve
mkdir -p gpurun_out
echo "=== ncu full (kernel bench, 1 rep)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ortho_sweep|spmm|vwxr" -c 12 -f -o gpurun_out/prof_r01b_kernels python scripts/kernel_bench.py --reps 1 > gpurun_out/ncu_full_b.log 2>&1
tail -3 gpurun_out/ncu_full_b.log
echo "=== ncu launch list of one bench solve (first 3500 launches)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/bench_ncu_b.log 2>&1
tail -2 gpurun_out/launches_r01b.csv | cut -c1-300
ls -la gpurun_out
