#!/bin/bash
# round 2, session t (1 GPU): where the wall time of the C2 solve goes on the host (PB200_HOST_PROFILE), with and
# without the tensor-map cache; profiling events off (--no-... keeps them on: bench needs them) 
for c in 1; do
  echo "=== C2 solve, tensor-map cache $c"
  PB200_TMAP_CACHE=$c PB200_HOST_PROFILE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --c5-n 0 --c3-n 0 --c4-m 0 2>&1 | grep "host profile\|^{" | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'share', d['roofline']['device_time_share_of_solve'])
    else: print(l.strip())
" | tail -5
done
