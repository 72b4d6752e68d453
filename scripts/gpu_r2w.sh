#!/bin/bash
# round 2, session w (1 GPU): spmm parity after the per-warp long-row partials, power-law timing vs long-row threshold,
# then ncu --set full of every hot kernel at the C2 and C5 shapes (summaries for profiles/), source-level stalls of the
# column-split restart kernel
mkdir -p gpurun_out /tmp/prof
echo "=== spmm parity"
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_zkernels_gpu.py -m gpu -q --timeout 120 -k "spmm or dlarnv" 2>&1 | tail -3
for t in 1024 512 256; do
  echo "=== C5 SpMM, PB200_SPMM_LONGROW=$t"
  PB200_SPMM_LONGROW=$t timeout 300 python scripts/kernel_bench.py --config c5 --only "spmm" 2>&1 | grep "^spmm"
done
for cfg in c2 c5; do
  echo "=== ncu full, $cfg shapes"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ortho_sweep|spmm|vwxr|dist_push|larnv" -c 30 -f \
     -o /tmp/prof/prof_r02_${cfg} python scripts/kernel_bench.py --reps 1 --config $cfg > gpurun_out/ncu_full_r02_$cfg.log 2>&1
  tail -1 gpurun_out/ncu_full_r02_$cfg.log
  ncu -i /tmp/prof/prof_r02_${cfg}.ncu-rep --page raw --csv > /tmp/prof/raw_$cfg.csv 2>/dev/null
  python scripts/summarize_ncu.py /tmp/prof/raw_$cfg.csv > gpurun_out/ncu_full_r02_${cfg}_summary.md
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
  cut -c1-260 gpurun_out/ncu_full_r02_${cfg}_summary.md | head -36
done
ncu -i /tmp/prof/prof_r02_c2.ncu-rep --page source --csv -k regex:vwxr_cg > gpurun_out/ncu_vwxr_cg_source.csv 2>/dev/null
ncu -i /tmp/prof/prof_r02_c2.ncu-rep --page details -k regex:vwxr_cg 2>/dev/null | grep -v "^ *$" | head -260 > gpurun_out/ncu_vwxr_cg_details.txt
ls -la gpurun_out | tail -8
