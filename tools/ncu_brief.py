#!/usr/bin/env python
"""Print a compact per-launch summary (time, DRAM bytes, throughput, issue/pipe utilisation, top stall reasons)
of an ncu report: python tools/ncu_brief.py gpurun_out/x.ncu-rep"""
import csv, io, subprocess, sys

def num(x):
    try: return float(x.replace(',', ''))
    except Exception: return float('nan')

def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    stall = [i for i, h in enumerate(hdr) if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h]
    for r in rows[2:]:
        g = lambda k: num(r[col[k]]) if k in col else float('nan')
        name = r[col['Kernel Name']].replace('void unnamed>::', '').split('(')[0]
        t = g('gpu__time_duration.sum')
        rd, wr = g('dram__bytes_read.sum'), g('dram__bytes_write.sum')
        print(f"{name:34s} {t:8.1f} us regs {g('launch__registers_per_thread'):.0f} smem {g('launch__shared_mem_per_block_dynamic'):.1f}KB "
              f"occ_lim(reg/smem) {g('launch__occupancy_limit_registers'):.0f}/{g('launch__occupancy_limit_shared_mem'):.0f} "
              f"DRAM rd {rd:.3f} wr {wr:.3f} ({r[col['dram__bytes_read.sum']+0] and ''}units as reported) dram% {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} "
              f"issue% {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} fp64% {g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.1f} "
              f"warps {g('smsp__warps_active.avg.per_cycle_active'):.2f} elig {g('smsp__warps_eligible.avg.per_cycle_active'):.2f} inst {g('smsp__inst_executed.sum')/1e6:.1f}M "
              f"bankconf {g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum')/1e6:.2f}M")
        vals = sorted(((hdr[i].split('issue_stalled_')[-1].replace('_per_issue_active.ratio', ''), num(r[i])) for i in stall), key=lambda t: -t[1])[:7]
        print("      stalls/issue: " + "  ".join(f"{a} {b:.2f}" for a, b in vals))

if __name__ == "__main__":
    main(sys.argv[1])
