#!/usr/bin/env python
"""Summarise an ``ncu --set full`` report (.ncu-rep) into a small CSV for profiles/.

    python tools/ncu_summary.py gpurun_out/r01_attn.ncu-rep > profiles/r01_ncu_attn.csv

One line per profiled launch with the counters the roofline discussion uses
(/opt/skills/guides/B200_PROFILING.md): duration, DRAM bytes read/written, DRAM throughput %,
tensor-pipe active %, SM throughput %, L2 bytes, registers, grid, achieved occupancy.
"""
import csv
import subprocess
import sys

WANT = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    out = csv.writer(sys.stdout)
    out.writerow([hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in idx])
    for r in data:
        out.writerow([r[i] for i in idx])


if __name__ == "__main__":
    main()
