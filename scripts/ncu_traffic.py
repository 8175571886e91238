#!/usr/bin/env python
"""scripts/ncu_traffic.py REPORT.ncu-rep KERNEL BATCH STEPS_PER_LAUNCH SUBSET.csv -- reads one `ncu --set full` capture (no GPU
needed), writes the subset of raw metrics the round's write-up quotes to SUBSET.csv and records the measured DRAM traffic per
step in profiles/traffic.json, which bench.py's roofline.traffic reads (a measurement of the committed capture, never a
constant in the source)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kernel, batch, steps, subset = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals))
u = dict(zip(hdr, units))
KEEP = ("Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__cycles_active", "gpu__dram_throughput", "sm__warps_active.avg.pct", "smsp__issue_active.avg.pct",
        "sm__pipe_fma_cycles_active.avg.pct", "sm__inst_executed_pipe_fma.sum.pct", "sm__pipe_tensor_cycles_active", "launch__registers_per_thread",
        "launch__cluster", "launch__grid_size", "launch__block_size", "launch__occupancy_limit", "smsp__inst_executed.sum",
        "smsp__pcsamp_warps_issue_stalled", "smsp__average_warp", "smsp__warp_issue_stalled", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate")
with open(subset, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    for k in hdr:
        if any(k.startswith(p) for p in KEEP):
            w.writerow([k, u[k], d[k]])


def to_bytes(key):
    v, unit = float(d[key].replace(",", "")), u[key].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]


per_step = (to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")) / steps
tj = os.path.join(ROOT, "profiles", "traffic.json")
t = json.load(open(tj)) if os.path.exists(tj) else {}
t.setdefault(kernel, {})[str(batch)] = {"dram_bytes_per_step": per_step, "steps_per_launch": steps, "capture": os.path.relpath(subset, ROOT)}
json.dump(t, open(tj, "w"), indent=1, sort_keys=True)
print(f"{kernel} B={batch}: {per_step:.0f} DRAM bytes per step over {steps} steps -> {tj}; subset -> {subset}")
