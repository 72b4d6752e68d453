#!/bin/bash
# round 2, session y (N GPUs): push / gather kernel times of the row-sharded SpMM on C5 (debug events, one
# synchronisation per block: not a bench number): bash scripts/gpu_r2y.sh N
N=${1:-2}
mkdir -p gpurun_out
PB200_DIST_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
   bench.py --gpus $N --steps 1 --warmup 1 --c5-steps 1 --c4-m 0 > gpurun_out/bench_dt_n$N.json 2> gpurun_out/bench_dt_n$N.err
grep "dist timing" gpurun_out/bench_dt_n$N.err | sort | uniq | head -40
