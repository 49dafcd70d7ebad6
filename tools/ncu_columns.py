#!/usr/bin/env python
"""Collect selected counters of several .ncu-rep files into one CSV (one column per capture):
    python tools/ncu_columns.py out.csv label=path.ncu-rep [label=path.ncu-rep ...]"""
import csv
import subprocess
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
cols, names = {}, []
for arg in sys.argv[2:]:
    label, path = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            d[h] = (v + " " + u).strip()
    cols[label] = d
    names.append(label)
keys = []
for d in cols.values():
    for k in d:
        if k not in keys:
            keys.append(k)
with open(sys.argv[1], "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["metric"] + names)
    for k in keys:
        w.writerow([k] + [cols[n].get(k, "") for n in names])
print("wrote", sys.argv[1], len(keys), "metrics x", len(names), "captures")
