#!/bin/bash
# round 2, session k (1 GPU): the column-split restart kernel (vwxr_cg_kernel): parity + timing against the
# one-warp-per-8-rows kernel at the C2 and C5 restart shapes, then the C2 solve
mkdir -p gpurun_out
echo "=== vwxr parity"
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "vwxr" 2>&1 | tail -6
for cfg in c2 c5; do
  for cg in 1 0; do
    echo "=== restart kernel, $cfg, PB200_VWXR_CG=$cg"
    PB200_DEBUG=1 PB200_VWXR_CG=$cg timeout 300 python scripts/kernel_bench.py --config $cfg --only "vwxr restart" 2>&1 | grep -v "^{" | tail -4
  done
done
echo "=== C2 solve"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --c5-n 0 --c3-n 0 --c4-m 0 > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2k.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['all_kernels'], d['config']['outer_iterations'], d['config']['matvecs_per_solve'])
PY
