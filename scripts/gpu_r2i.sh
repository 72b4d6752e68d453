#!/bin/bash
# round 2, session i (1 GPU): evict_first SpMM, C5-shape kernel bench, bench line, ncu captures for profiles/
mkdir -p gpurun_out
echo "=== kernel bench c5 (per-GPU shapes of C5 at N = 8)"
timeout 400 python scripts/kernel_bench.py --reps 10 --config c5 2>&1 | grep -v "^{" | tee gpurun_out/kernel_bench_c5_r2i.txt
echo "=== spmm without evict_first"
PB200_NO_EVICT_FIRST=1 timeout 300 python scripts/kernel_bench.py --reps 10 --config c5 --only spmm 2>&1 | grep -v "^{" | tee -a gpurun_out/kernel_bench_c5_r2i.txt
echo "=== kernel bench c2"
timeout 300 python scripts/kernel_bench.py --reps 10 2>&1 | grep -v "^{" | tee gpurun_out/kernel_bench_c2_r2i.txt
echo "=== bench"
PB200_DEBUG=1 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r2i.json') if l.startswith('{')][-1])
print('C2 ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'cpu', d['cpu_baseline'])
print('c5', {k: d['c5'][k] for k in ('ms_per_solve','matvecs_per_s','outer_iterations','kernels_rank0')} if d.get('c5') and 'error' not in d['c5'] else d.get('c5'))
print('c3', {k: d['c3'][k] for k in ('ms_per_solve','matvecs_per_s','matvecs_per_solve','gpu_launches_per_solve','kernels')} if d.get('c3') and 'error' not in d['c3'] else d.get('c3'))
PY
echo "=== ncu full, C2 shapes"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ortho_sweep|spmm|vwxr" -c 24 -f \
   -o gpurun_out/prof_r02_c2_kernels python scripts/kernel_bench.py --reps 1 > gpurun_out/ncu_full_r02_c2.log 2>&1
tail -2 gpurun_out/ncu_full_r02_c2.log
echo "=== ncu full, C5 shapes"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ortho_sweep|spmm|vwxr" -c 24 -f \
   -o gpurun_out/prof_r02_c5_kernels python scripts/kernel_bench.py --reps 1 --config c5 > gpurun_out/ncu_full_r02_c5.log 2>&1
tail -2 gpurun_out/ncu_full_r02_c5.log
echo "=== ncu launch list of one bench solve"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r02.csv \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --sampler none --c5-n 0 --c3-n 0 > gpurun_out/bench_ncu_r02.log 2>&1
tail -1 gpurun_out/launches_r02.csv | cut -c1-200
ls -la gpurun_out/*.ncu-rep
