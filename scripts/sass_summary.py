#!/usr/bin/env python
"""Blackwell evidence from the built product library: per hot kernel, the counts of the SASS mnemonics that matter
(UTMALDG = tensor-map TMA loads, UBLKCP = bulk copies, SYNCS = mbarrier ops, DMMA = fp64 tensor-core MMA, LDG.E.*.256 =
32-byte gathers, DFMA, RED/ATOM, ...) and the first lines of the listing.  Usage: python scripts/sass_summary.py [out.md]"""
import collections
import re
import subprocess
import sys

LIB = "primme_b200/libprimme_b200.so"
HOT = ["ortho_sweep_mma_exact_kernel", "ortho_sweep_mma_kernel", "vwxr_mma_kernel", "vwxr_cg_kernel", "spmm_rm_kernel", "spmm_tma_kernel",
       "spmm_win_kernel", "spmm_win_analyze", "larnv_kernel",
       "dist_push_kernel", "zsweep_kernel", "ztall_kernel", "zdots_kernel", "spmm_pack_kernel"]
KEYS = ["UTMALDG", "UBLKCP", "SYNCS", "DMMA", "DFMA", "LDG.E.ENL2.256", "LDG.E.128", "LDG.E.64", "STG.E.128", "STG.E.64",
        "LDS", "STS", "SHFL", "BAR", "ATOM", "RED", "STG.E.64.STRONG.SYS", "LDG.E.64.STRONG.SYS", "MEMBAR", "HMMA", "UTCMMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)[1:]
    stats = collections.OrderedDict()
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        base = next((h for h in HOT if h in dem), None)
        if not base:
            continue
        ins = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f)
        c = collections.Counter()
        for i in ins:
            # longest key first so that the system-scope accesses are not counted as plain LDG/STG
            for k in sorted(KEYS, key=len, reverse=True):
                if i.startswith(k):
                    c[k] += 1
                    break
        m = re.search(re.escape(base) + r"(<[^>]*>)?", dem)
        short = m.group(0) if m else base
        stats[short] = (len(ins), c)
    lines = ["# SASS summary of the hot kernels (cuobjdump -sass primme_b200/libprimme_b200.so, sm_100a)", "",
             "`UTMALDG` = cp.async.bulk.tensor (TMA tensor maps), `UBLKCP` = cp.async.bulk, `SYNCS` = mbarrier, `DMMA` = fp64 "
             "tensor-core MMA (tcgen05 has no fp64 kind), `LDG.E.ENL2.256` = 32-byte gather of the row-major SpMM, "
             "`*.STRONG.SYS` = system-scope release/acquire of the peer-memory halo protocol.", "",
             "| kernel instance | instr | " + " | ".join(KEYS) + " |", "|---|---|" + "---|" * len(KEYS)]
    for name, (n, c) in stats.items():
        lines.append(f"| `{name}` | {n} | " + " | ".join(str(c.get(k, 0)) if c.get(k, 0) else "" for k in KEYS) + " |")
    txt = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(txt)
    print(txt[:3000])


if __name__ == "__main__":
    main()
