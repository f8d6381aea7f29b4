#!/usr/bin/env python
"""Summarise an `ncu --set full` report into the small CSV kept under profiles/ (the .ncu-rep files stay in
gpurun_out/, which is scratch).  Usage: python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/out.csv"""
import csv
import io
import subprocess
import sys

KEEP = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    print(f"{out}: {len(rows) - 2} launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
