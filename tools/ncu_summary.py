#!/usr/bin/env python
"""Markdown table of the headline metrics of every kernel in an ncu report.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.md
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = []
    for r in data:
        n = r[hdr.index("Kernel Name")]
        n = n.split("(")[0].split("::")[-1]
        names.append(n)
    print("| metric | unit | " + " | ".join(names) + " |")
    print("|---|---|" + "---|" * len(names))
    for w in WANT:
        if w not in hdr:
            continue
        i = hdr.index(w)
        vals = []
        for r in data:
            try:
                vals.append(f"{float(r[i].replace(',', '')):.4g}")
            except ValueError:
                vals.append(r[i])
        print(f"| {w} | {units[i]} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
