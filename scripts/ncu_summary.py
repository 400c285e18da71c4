#!/usr/bin/env python
"""Compact per-kernel summary of an `ncu --set full` report: python scripts/ncu_summary.py X.ncu-rep [out.json]

Prints (and optionally stores as JSON) for every profiled launch: duration, DRAM bytes read / written, DRAM and SM
throughput (% of peak), tensor-pipe activity, achieved occupancy, registers, grid - the counters DESIGN.md / bench.py cite.
"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_hmma_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_xu.sum": "xu_inst",
    "sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed": "xu_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__cycles_active.avg": "sm_cycles_active",
}
UNIT_SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    d = {"kernel": r[hdr.index("Kernel Name")][:120]}
    for i, h in enumerate(hdr):
        key = WANT.get(h)
        if key is None or r[i] == "":
            continue
        try:
            v = float(r[i].replace(",", ""))
        except ValueError:
            continue
        v *= UNIT_SCALE.get(units[i], 1.0) if key.endswith(("_MB", "_us")) else 1.0
        d[key] = round(v, 3)
    if "dram_read_MB" in d and "dram_write_MB" in d:
        d["dram_traffic_MB"] = round(d["dram_read_MB"] + d["dram_write_MB"], 3)
    out.append(d)
for d in out:
    print(json.dumps(d))
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as f:
        json.dump(out, f, indent=1)
