#!/usr/bin/env python
"""Markdown summary of one or more .ncu-rep captures (ncu -i ... --page raw --csv): the metrics DESIGN.md and the bench
line quote.  Usage: python tools/ncu_summary.py title=path.ncu-rep [...] > profiles/rN_ncu_summary.md  (no GPU needed)"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
]


def main():
    print("# ncu --set full summaries (tools/ncu_summary.py)\n")
    print("Captured under gpurun with `ncu --set full --clock-control none --import-source on` after the same command exited 0 "
          "without ncu; per-launch values of ONE launch (cold caches, serialised: compare shares and ratios, not absolute times).\n")
    for arg in sys.argv[1:]:
        title, path = arg.rsplit("=", 1)
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"## {title}\n\n(no data in {path})\n")
            continue
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            print(f"## {title}: `{name}`\n")
            print("| metric | value | unit |\n|---|---|---|")
            for i, h in enumerate(hdr):
                if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                    print(f"| {h} | {vals[i]} | {units[i]} |")
            print()


if __name__ == "__main__":
    main()
