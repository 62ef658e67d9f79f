#!/usr/bin/env python
"""Condenses an .ncu-rep into the handful of metrics the design argues from (run here, no GPU)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        print("kernel:", d.get("Kernel Name", ("", "?"))[1][:110])
        for k in KEYS:
            if k in d:
                print(f"  {k:82s} {d[k][1]:>18s} {d[k][0]}")
        st = [(float(v.replace(",", "")), h) for h, (u, v) in d.items() if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued")]
        tot = sum(x for x, _ in st) or 1.0
        print("  warp stall samples:")
        for x, h in sorted(st, reverse=True)[:6]:
            print(f"    {100 * x / tot:6.2f}%  {h.replace('smsp__pcsamp_warps_issue_stalled_', '')}")


if __name__ == "__main__":
    main(sys.argv[1])
