#!/bin/bash
# sharded bench only, at N ranks (under gpurun --gpus N): bash scripts/gpu_multi_bench.sh N
N=${1:-8}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit $?"; grep -v "^W\|^\*" gpurun_out/bench_n$N.err | tail -5; grep "^{" gpurun_out/bench_n$N.json | cut -c1-1200
