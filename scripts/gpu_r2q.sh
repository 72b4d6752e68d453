#!/bin/bash
# round 2, session q (1 GPU): ncu of the row-major SpMM on the power-law matrix, grouped and nnz-balanced consumer
mkdir -p gpurun_out /tmp/prof
for bal in 0 1; do
  PB200_SPMM_BAL=$bal timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmm_rm" -s 2 -c 1 -f \
     -o /tmp/prof/spmm_bal$bal python scripts/kernel_bench.py --reps 1 --config c5 --only spmm > gpurun_out/ncu_spmm_bal$bal.log 2>&1
  tail -2 gpurun_out/ncu_spmm_bal$bal.log
  ncu -i /tmp/prof/spmm_bal$bal.ncu-rep --page raw --csv > gpurun_out/ncu_spmm_bal${bal}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/spmm_bal$bal.ncu-rep --page source --csv > gpurun_out/ncu_spmm_bal${bal}_source.csv 2>/dev/null
  ncu -i /tmp/prof/spmm_bal$bal.ncu-rep --page details 2>/dev/null | grep -v "^ *$" | head -300 > gpurun_out/ncu_spmm_bal${bal}_details.txt
  ls -la gpurun_out/ncu_spmm_bal${bal}_*
done
