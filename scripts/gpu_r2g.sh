#!/bin/bash
# round 2, session g (1 GPU): persisting-L2 window experiment on C2 (resident solve only)
mkdir -p gpurun_out
for mb in 0 48 80 110; do
  echo "=== L2 persist $mb MB"
  PB200_DEBUG=1 PB200_L2_PERSIST_MB=$mb timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --c5-n 0 --c3-n 0 > gpurun_out/bench_r2g_$mb.json 2> gpurun_out/bench_r2g_$mb.err
  grep "L2 persisting" gpurun_out/bench_r2g_$mb.err | head -1
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r2g_$mb.json') if l.startswith('{')][-1])
print('ms', round(d['ms_per_step'],1), 'frac', round(d['roofline']['frac'],4), {k:(v['GBps'], round(v['ms'],1)) for k,v in d['roofline']['all_kernels'].items() if k in ('spmm','ortho_sweep','vwxr')}, 'its', d['config']['outer_iterations'])
PY
done
