#!/usr/bin/env python
"""Summarise an `ncu --set full` report: one row per captured launch with the metrics the roofline
argument uses (duration, DRAM bytes = traffic, DRAM / fp64-pipe / L1 utilisation, occupancy, shared
memory bank conflicts, top warp stall reasons).
Usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python scripts/summarize_ncu.py raw.csv"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, body = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def scaled(r, name):
    """durations in us, byte counts in MB, whatever unit ncu chose for the column"""
    return f(r, name) * SCALE.get(units[col[name]], 1.0)


def f(r, name):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:
        return float("nan")


stall = [h for h in hdr if re.match(r"smsp__average_warps?_issue_stalled_.*_per_issue_active\.ratio$", h)
         or re.match(r"smsp__average_warp_latency_issue_stalled_.*\.ratio$", h)]
print("| kernel | grid x block | regs | us | DRAM rd MB | DRAM wr MB | DRAM GB/s | DFMA pipe % | DMMA pipe % | L1/LSU % | issue % | warps active % | smem ld conflicts | top stalls |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in body:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void <unnamed>::", "")
    st = sorted(((f(r, h), re.sub(r".*stalled_(.*?)(_per_issue_active)?\.ratio", r"\1", h)) for h in stall), reverse=True)[:3]
    print("| %s | %s x %s | %d | %.1f | %.1f | %.1f | %.0f | %.1f | %.1f | %.1f | %.1f | %.1f | %.0f | %s |" % (
        name, int(f(r, "launch__grid_size")), int(f(r, "launch__block_size")),
        f(r, "launch__registers_per_thread"), scaled(r, "gpu__time_duration.sum"),
        scaled(r, "dram__bytes_read.sum"), scaled(r, "dram__bytes_write.sum"),
        (scaled(r, "dram__bytes_read.sum") + scaled(r, "dram__bytes_write.sum")) / scaled(r, "gpu__time_duration.sum") * 1e3,
        f(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        f(r, "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
        f(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        f(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
        ", ".join("%s %.1f" % (n, v) for v, n in st)))
