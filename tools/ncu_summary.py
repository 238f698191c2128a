#!/usr/bin/env python
"""Summarise an .ncu-rep (one `ncu --set full` capture) as a markdown table for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title / command line" > profiles/rN_x_ncu_summary.md
"""
import csv
import io
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct")
PREFIX = ("smsp__average_warps_issue_stalled_", "sm__inst_executed_pipe_")


def main():
    rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print("# %s\n" % title)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("## %s  (grid %s, block %s)\n" % (d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
        print("| metric | unit | value |\n|---|---|---|")
        for k, u in zip(hdr, units):
            if k in KEEP or (k.startswith(PREFIX) and "not_issued" not in k and k.endswith((".ratio", ".avg.pct_of_peak_sustained_active"))):
                print("| %s | %s | %s |" % (k, u, d[k]))
        print()


if __name__ == "__main__":
    main()
