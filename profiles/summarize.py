#!/usr/bin/env python
"""Summarise an .ncu-rep (from `ncu --set full`) into the handful of numbers DESIGN.md / bench.py cite.
usage: python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import io
import re
import subprocess
import sys
import collections

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "sm__cycles_active.avg"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for kr in rows[2:]:
        name = kr[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"== kernel: {name}")
        for i, h in enumerate(hdr):
            if h in KEYS or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                try:
                    v = float(kr[i])
                    if "issue_stalled" in h and v < 0.05:
                        continue
                    print(f"{h:80s} {v:18.4f} {units[i]}")
                except ValueError:
                    print(f"{h:80s} {kr[i]:>18s} {units[i]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = next((i for i, r in enumerate(rows) if "Instructions Executed" in r), None)
    if h is not None:
        hdr, data = rows[h], rows[h + 1:]
        iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
        ops = collections.Counter()
        for r in data:
            if len(r) <= iE:
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS])
            try:
                ops[m.group(2).split(".")[0] if m else "?"] += float(r[iE] or 0)
            except ValueError:          # a second kernel's header row inside the same csv
                continue
        tot = sum(ops.values())
        print(f"== SASS: {len(data)} static instructions, {tot:.0f} executed warp-instructions")
        for op, c in ops.most_common(16):
            print(f"   {op:8s} {100 * c / tot:5.1f}%  {c:.0f}")


if __name__ == "__main__":
    main()
