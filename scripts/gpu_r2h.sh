#!/bin/bash
# round 2, session h (1 GPU): fused QMR utilities + inline sweep coefficients (parity), finer L2 window sweep, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -k "zkernels or zprimme or jdqmr or utilities or driver or solver or host_contract" > gpurun_out/pytest_gpu_r2h.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2h.log
grep -E "passed|failed|FAILED|exit" gpurun_out/pytest_gpu_r2h.log | head -20
for mb in 32 40 56 64; do
  echo "=== L2 persist $mb MB"
  PB200_L2_PERSIST_MB=$mb timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --c5-n 0 --c3-n 0 > gpurun_out/bench_r2h_$mb.json 2> gpurun_out/bench_r2h_$mb.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r2h_$mb.json') if l.startswith('{')][-1])
print('ms', round(d['ms_per_step'],1), 'frac', round(d['roofline']['frac'],4), {k:(v['GBps'], round(v['ms'],1)) for k,v in d['roofline']['all_kernels'].items() if k in ('spmm','ortho_sweep','vwxr')}, 'its', d['config']['outer_iterations'])
PY
done
echo "=== bench default"
PB200_DEBUG=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r2h.json') if l.startswith('{')][-1])
print('C2 ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], d['roofline']['all_kernels'])
print('c5', {k: d['c5'][k] for k in ('ms_per_solve','matvecs_per_s','outer_iterations','kernels_rank0')} if d.get('c5') and 'error' not in d['c5'] else d.get('c5'))
print('c3', {k: d['c3'][k] for k in ('ms_per_solve','matvecs_per_s','matvecs_per_solve','gpu_launches_per_solve','kernels')} if d.get('c3') and 'error' not in d['c3'] else d.get('c3'))
PY
