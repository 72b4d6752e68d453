#!/bin/bash
# round 2, session j (1 GPU): ncu captures for profiles/ (reports converted to CSV summaries on the box: the
# .ncu-rep files exceed what gpurun copies back), plus the LOBPCG block > 8 GPU test
mkdir -p gpurun_out /tmp/prof
timeout 600 python -m pytest tests/test_solver_gpu.py -m gpu -q --timeout 300 -k "panel_width" 2>&1 | tail -3
for cfg in c2 c5; do
  echo "=== ncu full, $cfg shapes"
  timeout 900 ncu --set full --clock-control none -k regex:"ortho_sweep|spmm|vwxr|dist_push" -c 24 -f \
     -o /tmp/prof/prof_r02_${cfg} python scripts/kernel_bench.py --reps 1 --config $cfg > gpurun_out/ncu_full_r02_$cfg.log 2>&1
  tail -1 gpurun_out/ncu_full_r02_$cfg.log
  ncu -i /tmp/prof/prof_r02_${cfg}.ncu-rep --page raw --csv > /tmp/prof/raw_$cfg.csv 2>/dev/null
  python scripts/summarize_ncu.py /tmp/prof/raw_$cfg.csv > gpurun_out/ncu_full_r02_${cfg}_summary.md
  # a few extra columns for the SpMM rows: L2 hit rate and throughput
  python - <<PY
import csv
rows=list(csv.reader(open('/tmp/prof/raw_$cfg.csv')))
hdr=rows[0]; col={h:i for i,h in enumerate(hdr)}
want=[h for h in hdr if h in ('Kernel Name','gpu__time_duration.sum','lts__t_sector_hit_rate.pct','lts__throughput.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active')]
with open('gpurun_out/ncu_r02_${cfg}_extra.csv','w') as f:
    w=csv.writer(f); w.writerow(want); w.writerow([rows[1][col[h]] for h in want])
    for r in rows[2:]:
        w.writerow([r[col[h]][:60] for h in want])
PY
  head -30 gpurun_out/ncu_full_r02_${cfg}_summary.md | cut -c1-330
done
echo "=== ncu launch list of one bench solve"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r02.csv \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --sampler none --c5-n 0 --c3-n 0 --c4-m 0 > gpurun_out/bench_ncu_r02.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_r02.csv > gpurun_out/launches_r02_summary.txt; head -12 gpurun_out/launches_r02_summary.txt
gzip -f gpurun_out/launches_r02.csv
ls -la gpurun_out | tail -12; du -sh gpurun_out
