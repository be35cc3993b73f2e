#!/usr/bin/env python
"""Markdown summary of `ncu --page raw --csv` exports (one launch each): the metrics DESIGN.md and bench.py quote.
usage: tools/ncu_summary.py out.md title=file.raw.csv ..."""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe cycles active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed.sum", "thread instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instruction"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long scoreboard (cycles/issue)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 sectors read (from L1)"),
]


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    return {n: (v, u) for n, u, v in zip(names, units, vals)}, dict(zip(names, vals)).get("Kernel Name", "?")


out = open(sys.argv[1], "w")
for spec in sys.argv[2:]:
    title, path = spec.split("=", 1)
    try:
        m, kname = load(path)
    except Exception as e:                       # noqa: BLE001
        out.write(f"## {title}\n\n(capture missing: {e})\n\n")
        continue
    out.write(f"## {title}\n\nkernel: `{kname[:160]}`\n\n| metric | value |\n|---|---:|\n")
    for key, label in WANT:
        if key in m and m[key][0] != "":
            out.write(f"| {label} (`{key}`) | {m[key][0]} {m[key][1]} |\n")
    out.write("\n")
out.close()
