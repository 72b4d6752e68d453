#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel (name, grid)
launch count, total and mean duration, share of the summed device time."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) < 15:
        continue
    name = re.sub(r"\(.*", "", r[4]).replace("void <unnamed>::", "")
    agg[name][0] += 1
    agg[name][1] += float(r[-1])
tot = sum(v[1] for v in agg.values())
fam = collections.defaultdict(lambda: [0, 0.0])
for k, v in agg.items():
    f = ("ortho_sweep" if "ortho_sweep" in k else "vwxr" if "vwxr" in k else "spmm" if "spmm" in k
         else "panel_reduce" if "reduce" in k else "utils")
    fam[f][0] += v[0]
    fam[f][1] += v[1]
print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.2f} ms summed kernel time (ncu: serialised, cold caches)")
print("## by kernel family")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:14s} launches={v[0]:6d} total_ms={v[1] / 1e6:9.2f} share={v[1] / tot:6.3f}")
print("## by kernel instantiation")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{k[:60]:60s} n={v[0]:6d} total_ms={v[1] / 1e6:9.2f} avg_us={v[1] / v[0] / 1e3:8.1f} share={v[1] / tot:6.3f}")
