#!/usr/bin/env python
"""Key metrics of one `ncu --set full` capture (raw csv) followed by the hottest source lines (source csv).
usage: python profiles/ncu_summary.py <raw.csv> <src.csv> <kernel source .cu> "<header line>" > profiles/rNN_ncu_<kernel>_summary.txt"""
import csv
import io
import sys
from contextlib import redirect_stdout

sys.path.insert(0, __file__.rsplit("/", 1)[0])
import summarize  # noqa: E402

WANT = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum")


def main(raw_csv, src_csv, cu, header=""):
    rows = list(csv.reader(open(raw_csv)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    if header:
        print("# " + header)
    for i, h in sorted(enumerate(hdr), key=lambda t: t[1]):
        if h in WANT or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            print(f"{h:95s} {units[i]:10s} {data[0][i]}")
    print()
    buf = io.StringIO()
    with redirect_stdout(buf):
        summarize.main(src_csv, raw_csv, cu)
    print("# warp-state samples per source line (top 40)")
    print("\n".join(l for l in buf.getvalue().split("\n") if l.startswith("total samples") or "%" in l and "excwf" in l))


if __name__ == "__main__":
    main(*sys.argv[1:])
