#!/bin/bash
# round 2, session s (1 GPU): the whole GPU suite with the new restart kernel / windowed SpMM / lane-shared gathers,
# then the default bench line (C2 + c5 + c3 + c4 blocks)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu_r2s.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2s.log
grep -E "passed|failed|FAILED|exit|Error" gpurun_out/pytest_gpu_r2s.log | head -20
echo "=== bench default"
PB200_DEBUG=1 timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2s.json 2> gpurun_out/bench_r2s.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r2s.json') if l.startswith('{')][-1])
print('C2 ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], d['roofline']['all_kernels'], 'share', d['roofline'].get('device_time_share_of_solve'))
print('counts', d['config']['outer_iterations'], d['config']['restarts'], d['config']['matvecs_per_solve'])
print('c5', {k: d['c5'][k] for k in ('ms_per_solve','matvecs_per_s','outer_iterations','kernels_rank0')} if d.get('c5') and 'error' not in d['c5'] else d.get('c5'))
print('c3', {k: d['c3'][k] for k in ('ms_per_solve','matvecs_per_s','matvecs_per_solve','gpu_launches_per_solve','kernels')} if d.get('c3') and 'error' not in d['c3'] else d.get('c3'))
print('c4', d.get('c4'))
print('cpu_baseline', d.get('cpu_baseline'))
PY
grep "SpMM b=\|window" gpurun_out/bench_r2s.err | head
