#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel bench c2: fused finish, polled"
timeout 200 python scripts/kernel_bench.py --reps 10 2>&1 | grep -v "^{" | tee gpurun_out/kb_a.txt
echo "=== kernel bench c2: separate reduce kernel"
PB200_NO_FUSED_FINISH=1 timeout 200 python scripts/kernel_bench.py --reps 10 2>&1 | grep -v "^{" | tee gpurun_out/kb_b.txt
echo "=== kernel bench c2: fused finish to device memory + memcpy"
PB200_NO_POLL=1 timeout 200 python scripts/kernel_bench.py --reps 10 2>&1 | grep -v "^{" | tee gpurun_out/kb_c.txt
