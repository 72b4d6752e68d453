#!/bin/bash
# round 2, session z (1 GPU): the default bench line with the host profile of every solve, the reference arm once as a
# FULL solve (time-to-converge of the unmodified reference on this box's host cores), the ncu launch list of one solve
mkdir -p gpurun_out
PB200_HOST_PROFILE=1 timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2z.json 2> gpurun_out/bench_r2z.err
grep "host profile" gpurun_out/bench_r2z.err | tail -12 | cut -c1-330
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r2z.json') if l.startswith('{')][-1])
print('C2 ms', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'share', d['roofline'].get('device_time_share_of_solve'), d['roofline']['all_kernels'])
for k in ('c5','c3','c4'):
    c=d.get(k); print(k, {kk: c.get(kk) for kk in ('ms_per_solve','matvecs_per_s','applications_per_s','outer_iterations','matvecs_per_solve','kernels_rank0','kernels')} if c else None)
print('cpu_baseline', d.get('cpu_baseline')); print('clocks', d.get('clocks'))
PY
echo "=== reference arm, full solve"
timeout 1500 python bench.py --impl reference --steps 1 --warmup 0 --ref-matvecs 0 > gpurun_out/bench_r2z_reference_full.json 2> gpurun_out/bench_r2z_reference_full.err
cut -c1-900 gpurun_out/bench_r2z_reference_full.json
echo "=== ncu launch list of one bench solve"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r02.csv \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --sampler none --c5-n 0 --c3-n 0 --c4-m 0 > gpurun_out/bench_ncu_r02.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_r02.csv > gpurun_out/launches_r02_summary.txt; head -16 gpurun_out/launches_r02_summary.txt
gzip -f gpurun_out/launches_r02.csv
