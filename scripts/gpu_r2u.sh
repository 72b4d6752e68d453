#!/bin/bash
# round 2, session u (1 GPU): dlarnv on the device (bit parity), solver parity incl. exact counts, C2 bench with profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "dlarnv" 2>&1 | tail -3
timeout 1500 python -m pytest tests/test_solver_gpu.py tests/test_zprimme_gpu.py tests/test_driver_gpu.py -m gpu -q --timeout 900 2>&1 | tail -3
PB200_HOST_PROFILE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --c3-n 0 --c4-m 0 > gpurun_out/bench_r2u.json 2> gpurun_out/bench_r2u.err
grep "host profile" gpurun_out/bench_r2u.err | tail -4
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r2u.json') if l.startswith('{')][-1])
print('C2 ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'share', d['roofline'].get('device_time_share_of_solve'), 'counts', d['config']['outer_iterations'], d['config']['matvecs_per_solve'])
print('c5', {k: d['c5'][k] for k in ('ms_per_solve','matvecs_per_s','outer_iterations','kernels_rank0')} if d.get('c5') and 'error' not in d['c5'] else d.get('c5'))
PY
