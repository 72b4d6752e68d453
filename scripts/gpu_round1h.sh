#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel bench c2 (default build, 8 KB coefficient block)"
timeout 200 python scripts/kernel_bench.py --reps 10 2>&1 | grep -v "^{" | tee gpurun_out/kernel_bench_c2_v7.txt
echo "=== kernel bench c2 (2 KB coefficient block)"
PB200_LIB=$PWD/build/var256/libprimme_b200_c256.so timeout 200 python scripts/kernel_bench.py --reps 10 2>&1 | grep -v "^{" | tee gpurun_out/kernel_bench_c2_v7_c256.txt
echo "=== ncu launch list, default"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_kb_v7.csv \
   python scripts/kernel_bench.py --reps 2 > /dev/null 2>&1
grep -E "ortho_sweep|vwxr|spmm" gpurun_out/launches_kb_v7.csv | awk -F'","' '{print substr($5,1,50), $NF}' | head -40
echo "=== ncu launch list, c256"
PB200_LIB=$PWD/build/var256/libprimme_b200_c256.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_kb_v7_c256.csv \
   python scripts/kernel_bench.py --reps 2 > /dev/null 2>&1
grep -E "ortho_sweep|vwxr|spmm" gpurun_out/launches_kb_v7_c256.csv | awk -F'","' '{print substr($5,1,50), $NF}' | head -40
