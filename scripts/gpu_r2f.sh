#!/bin/bash
# round 2, session f (2 GPUs): 2-GPU tests (row-partitioned SVD operator), then bench at N=1 (e2e with pooled storage, c3/c5 blocks)
mkdir -p gpurun_out
export PB200_DEBUG=1
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_kernels_gpu.py -m gpu -q -k "two_gpu or reproducible" --timeout 900 > gpurun_out/pytest_r2f.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r2f.log
grep -E "passed|failed|FAILED|Error|exit" gpurun_out/pytest_r2f.log | head; tail -12 gpurun_out/pytest_r2f.log | cut -c1-300
echo "=== bench N=1"
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; grep dprimme_csr gpurun_out/bench_r2f.err | tail -4; cat gpurun_out/bench_r2f.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('C2 ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'])
print('c5', {k: d['c5'][k] for k in ('ms_per_solve','matvecs_per_s','outer_iterations')} if d.get('c5') and 'error' not in d['c5'] else d.get('c5'))
print('c3', d.get('c3'))
"
