#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export into one line per captured launch.
usage: python scripts/ncu_summary.py gpurun_out/prof_TAG_raw.csv [--md]"""
import csv, sys
COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("launch__grid_size", "grid"), ("launch__block_size", "blk"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smemwave"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__inst_executed.sum", "inst"),
        ("smsp__sass_inst_executed_op_local_ld.sum", "lld"), ("smsp__sass_inst_executed_op_local_st.sum", "lst")]

def conv(v, unit, want):
    try: x = float(v.replace(",", ""))
    except ValueError: return v
    if want in ("rdMB", "wrMB"):
        f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
        return f"{x * f:.2f}"
    if want == "us":
        f = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
        return f"{x * f:.1f}"
    return f"{x:.4g}"

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [(c, n) for c, n in COLS if c in idx]
print("kernel | " + " | ".join(n for _, n in cols))
for r in data:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
    print(name + " | " + " | ".join(conv(r[idx[c]], units[idx[c]], n) for c, n in cols))
