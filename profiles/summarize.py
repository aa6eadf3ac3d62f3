#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line, and key raw metrics.
usage: python profiles/summarize.py <src.csv> <raw.csv> [source.cu]"""
import collections
import csv
import sys


def main(src_csv, raw_csv, cu="ddrl_b200/csrc/fcnet.cu", top=40):
    rows = list(csv.reader(open(raw_csv)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
    for i, h in enumerate(hdr):
        if h in want:
            print(f"{h:90s} {units[i]:10s} {[d[i] for d in data]}")
    rows = list(csv.reader(open(src_csv)))
    agg = collections.Counter()
    exc = collections.Counter()
    cur, hdr = None, None
    for r in rows:
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if r and r[0].isdigit() and cur and hdr:
            d = dict(zip(hdr, r))
            try:
                s = int(d.get("# Samples", "0") or 0)
            except ValueError:
                s = 0
            agg[(cur, int(r[0]))] += s
            try:
                exc[(cur, int(r[0]))] += int(d.get("L1 Wavefronts Shared Excessive", "0") or 0)
            except ValueError:
                pass
    tot = sum(agg.values())
    print("total samples", tot)
    src = open(cu).read().split("\n")
    name = cu.split("/")[-1]
    for (f, l), s in agg.most_common(top):
        text = src[l - 1].strip()[:95] if f == name and l - 1 < len(src) else ""
        print(f"{f}:{l:4d} {s:6d} {100 * s / max(tot, 1):5.1f}% excwf={exc[(f, l)]:8d} | {text}")


if __name__ == "__main__":
    main(*sys.argv[1:])
