#!/bin/bash
# N-GPU check: the 2-rank row-sharded parity test, then the sharded bench at N ranks.
# usage (under gpurun --gpus N): bash scripts/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/pytest_multi.log 2>&1
echo "pytest multi exit $?"; tail -15 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit $?"; grep -v "^W\|^\*" gpurun_out/bench_n$N.err | tail -8; cat gpurun_out/bench_n$N.json
