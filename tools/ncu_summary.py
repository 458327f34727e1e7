#!/usr/bin/env python
"""Markdown table of the headline metrics of every launch in an .ncu-rep (input of profiles/*_ncu_summary.md)."""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max.per_second", "launch__grid_size",
        "launch__block_size", "launch__cluster_dim_x", "launch__registers_per_thread", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active"]

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("\n| metric | value |\n|---|---|")
        for name in WANT:
            if name in hdr:
                i = hdr.index(name)
                print("| `%s` | %s %s |" % (name, vals[i], units[i]))
