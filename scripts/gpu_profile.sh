#!/bin/bash
# Profile pass (one GPU): (1) ncu --set full of one launch of each hot kernel at the C2 shapes,
# (2) ncu launch list (gpu__time_duration) of the first launches of one bench solve.
# usage (under gpurun): bash scripts/gpu_profile.sh TAG
TAG=${1:-r01}
mkdir -p gpurun_out
echo "=== ncu full (kernel bench, 1 rep)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ortho_sweep|spmm|vwxr" -c 21 -f \
   -o gpurun_out/prof_${TAG}_kernels python scripts/kernel_bench.py --reps 1 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
echo "=== ncu launch list of one bench solve (first 4000 launches)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --sampler none > gpurun_out/bench_ncu_$TAG.log 2>&1
tail -2 gpurun_out/launches_$TAG.csv | cut -c1-300
ls -la gpurun_out | tail -8
