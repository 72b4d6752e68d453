#!/bin/bash
# round 2, session l (1 GPU): restart kernel timings, column-split kernel vs the previous one
for cfg in c2 c5; do
  for cg in 1 0; do
    echo "=== restart kernel, $cfg, PB200_VWXR_CG=$cg"
    PB200_DEBUG=1 PB200_VWXR_CG=$cg timeout 300 python scripts/kernel_bench.py --config $cfg --only "vwxr restart" 2>&1 | grep -i "vwxr_cg\|vwxr restart\|vwxr_mma<[2-9]"
  done
done
