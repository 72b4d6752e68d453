#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ortho_sweep_mma" -s 3 -c 3 -f -o gpurun_out/prof_mma_update python scripts/kernel_bench.py --reps 1 > gpurun_out/ncu_mma.log 2>&1
tail -2 gpurun_out/ncu_mma.log | cut -c1-300
