#!/usr/bin/env python
"""Turn an .ncu-rep (brought back from the GPU box in gpurun_out/) into a small markdown summary for profiles/.

  python tools/ncu_summary.py gpurun_out/fill_r1_c.ncu-rep profiles/r1_fill_summary.md [--traffic-json out.json --workload cfg2 --cells N]
"""
import argparse
import csv
import io
import json
import subprocess

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), CTAs/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (shared memory), CTAs/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy, % of 64 warps/SM"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy, % (1 warp-instruction/clk/SMSP)"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction (of 32)"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe, % of peak"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe, % of peak"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy pipe (IMAD/IDP) cycles active, %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe, % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1/shared data-pipe wavefronts, % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate, %"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate, %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput, % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput, % of peak"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (warps per issue)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
]


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, zip(units, r))) for r in rows[2:]]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep"); ap.add_argument("out")
    ap.add_argument("--title", default=None); ap.add_argument("--note", default="")
    ap.add_argument("--traffic-json"); ap.add_argument("--workload"); ap.add_argument("--cells", type=int)
    a = ap.parse_args()
    launches = load(a.rep)
    lines = [f"# {a.title or a.rep}", "", f"source: `{a.rep}` (ncu --set full --clock-control none --import-source on), read with "
             "`ncu -i ... --page raw --csv` by tools/ncu_summary.py.", ""]
    if a.note:
        lines += [a.note, ""]
    for k, L in enumerate(launches):
        name = L.get("Kernel Name", ("", "?"))[1]
        lines += [f"## launch {k}: `{name}`", "", "| metric | value |", "|---|---|"]
        for key, label in KEYS:
            if key in L:
                u, v = L[key]
                lines.append(f"| {label} (`{key}`) | {v} {u} |")
        lines.append("")
    open(a.out, "w").write("\n".join(lines))
    if a.traffic_json:
        L = launches[0]
        def b(key):
            u, v = L[key]
            return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        json.dump({"workload": a.workload, "cells": a.cells, "kernel": L["Kernel Name"][1], "source": a.rep,
                   "dram_bytes_read": b("dram__bytes_read.sum"), "dram_bytes_write": b("dram__bytes_write.sum"),
                   "duration_ms_under_ncu": L["gpu__time_duration.sum"]}, open(a.traffic_json, "w"), indent=1)
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
