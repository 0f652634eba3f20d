#!/usr/bin/env python3
"""Per-launch summary of an `ncu --set full` report as JSON (what profiles/*.json hold).
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rXX_ncu_full.json [label ...]
Labels (optional, one per captured launch, in order) name the launches; the metrics are read with
`ncu -i <rep> --page raw --csv`, so this runs on the CPU box."""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",          # tcgen05 (UTC*MMA) pipe
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",   # legacy mma.sync pipe
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
rep, out = sys.argv[1], sys.argv[2]
labels = sys.argv[3:]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)],
                     capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
res = []
for i, r in enumerate(data):
    d = {"launch": labels[i] if i < len(labels) else None}
    for h, u, v in zip(hdr, units, r):
        if h in ("Kernel Name", "Block Size", "Grid Size") or h in METRICS:
            d[f"{h} [{u}]" if u else h] = v
    res.append(d)
json.dump(res, open(out, "w"), indent=1)
print(f"{len(res)} launches -> {out}")
