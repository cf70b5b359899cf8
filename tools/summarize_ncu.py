#!/usr/bin/env python
"""Summarise `ncu --set full` reports (read here, no GPU needed) into a markdown table for profiles/.
   python tools/summarize_ncu.py gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...] > profiles/xxx.md"""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("smsp__inst_executed.sum", "warp insts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__inst_executed_op_tma_ld.sum", "TMA loads"),
]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    return r[0], r[1], r[2:]


def fmt(v, u):
    try:
        f = float(v.replace(",", ""))
    except ValueError:
        return v
    if u in ("%",):
        return "%.1f" % f
    if f >= 1e6 and u in ("inst", ""):
        return "%.1fM" % (f / 1e6)
    return ("%.3f %s" % (f, u)).strip() if u else ("%.3g" % f)


def main():
    print("| report | kernel | " + " | ".join(n for _, n in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    for rep in sys.argv[1:]:
        hdr, units, rows = rows_of(rep)
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows:
            name = r[idx["Kernel Name"]].replace("void ", "").split("(")[0]
            cells = []
            for m, _ in COLS:
                cells.append(fmt(r[idx[m]], units[idx[m]]) if m in idx else "-")
            print("| %s | `%s` | %s |" % (rep.split("/")[-1].replace(".ncu-rep", ""), name, " | ".join(cells)))


if __name__ == "__main__":
    main()
