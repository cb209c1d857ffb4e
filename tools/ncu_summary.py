#!/usr/bin/env python
"""Condense an ncu report into the JSON summaries kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep "how it was captured" > profiles/rNN_X_ncu_summary.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_active.avg"]

rep, how = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = {"capture": how, "report": rep.split("/")[-1], "kernels": []}
for r in rows[2:]:
    d = {"Kernel Name": r[hdr.index("Kernel Name")]}
    for k in KEYS:
        if k in hdr:
            d[k] = ("%s %s" % (r[hdr.index(k)], units[hdr.index(k)])).strip()
    stalls = {}
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(r[i])
    if stalls:
        d["warp_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
    out["kernels"].append(d)
print(json.dumps(out, indent=1))
