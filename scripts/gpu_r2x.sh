#!/bin/bash
# round 2, session x (N GPUs, under gpurun --gpus N): [N = 2: the 2-GPU tests first] then the sharded bench at N ranks
# with the host profile of every solve: bash scripts/gpu_r2x.sh N
N=${1:-8}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 600 2>&1 | tail -4
fi
if [ -n "$DT" ]; then
  # push / gather kernel times of the row-sharded SpMM (debug events, one synchronisation per block: not a bench number)
  PB200_DIST_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
     bench.py --gpus $N --steps 1 --warmup 1 --c5-steps 1 --c4-m 0 > gpurun_out/bench_dt_n$N.json 2> gpurun_out/bench_dt_n$N.err
  grep "dist timing" gpurun_out/bench_dt_n$N.err | sort | uniq | head -20
fi
PB200_HOST_PROFILE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit $?"; grep "host profile: solve" gpurun_out/bench_n$N.err | tail -3
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1])
print('C2 N=$N ms', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'share', d['roofline'].get('device_time_share_of_solve'), 'its', d['config']['outer_iterations'], d['config']['matvecs_per_solve'])
print(d['roofline']['all_kernels'])
c5=d.get('c5'); print('c5', {k: c5.get(k) for k in ('ms_per_solve','matvecs_per_s','outer_iterations','largest_eval','max_resnorm','kernels_rank0','halo_rank0')} if c5 else None)
print('c4', d.get('c4'))
PY
