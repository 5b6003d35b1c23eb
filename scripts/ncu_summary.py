#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/:
key raw metrics of the first captured kernel plus the hottest source lines (needs -lineinfo and
--import-source on at capture time).   python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("# ncu summary of", rep)
    for i, h in enumerate(hdr):
        if h in ("Kernel Name", "Block Size", "Grid Size"):
            print(f"{h}: {vals[i]}")
    for i, h in enumerate(hdr):
        if h in KEYS or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            print(f"{h} [{units[i]}] = {vals[i]}")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]))))
    head, cur, tot, smp, text = None, None, collections.OrderedDict(), {}, {}
    for r in src:
        if r and r[0] == "Line No":
            head = r; ii = head.index("Instructions Executed"); si = head.index("# Samples"); continue
        if head is None or len(r) <= ii:
            continue
        if r[0].strip():
            try:
                cur = int(r[0]); text[cur] = r[1]
            except ValueError:
                cur = None
            continue
        if cur is None:
            continue
        try:
            tot[cur] = tot.get(cur, 0) + int(r[ii]); smp[cur] = smp.get(cur, 0) + int(r[si])
        except ValueError:
            pass
    T, S = sum(tot.values()) or 1, sum(smp.values()) or 1
    print(f"\n# source lines with >= 1.5 % of executed warp instructions ({T} total) or stall samples ({S} total)")
    for ln in sorted(tot):
        if tot[ln] >= 0.015 * T or smp[ln] >= 0.015 * S:
            print(f"{ln:5d} {100 * tot[ln] / T:5.1f}% inst {100 * smp[ln] / S:5.1f}% samples | {text[ln].strip()[:110]}")


if __name__ == "__main__":
    main()
