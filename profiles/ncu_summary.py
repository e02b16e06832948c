#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + SASS pages) into the few numbers the profile notes quote.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [n_cells]"""
import collections
import csv
import io
import subprocess
import sys


def page(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--csv", *args], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    ncells = float(sys.argv[2]) if len(sys.argv) > 2 else 8388608.0
    rows = page(rep, "--page", "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
            "lts__t_sectors_srcunit_tex_op_write.sum", "launch__shared_mem_per_block_dynamic",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "Kernel Name"]
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                print(f"{w:70s} {units[i]:10s} {[r[i] for r in data]}")
    for r in data[:1]:
        for i, h in enumerate(hdr):
            if h == "smsp__inst_executed.sum":
                print("instructions per cell:", float(r[i]) * 32 / ncells)
    rows = page(rep, "--page", "source", "--print-source", "sass")
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not starts:
        return
    h = rows[starts[0]]
    end = starts[1] - 1 if len(starts) > 1 else len(rows)
    ix = {n: i for i, n in enumerate(h)}
    byop, stalls = collections.Counter(), collections.Counter()
    for r in rows[starts[0] + 1:end]:
        if len(r) < len(h):
            continue
        try:
            n = int(r[ix["Instructions Executed"]])
        except ValueError:
            continue
        src = r[ix["Source"]].split()
        op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
        byop[op] += n
        for s in h:
            if s.startswith("stall_") and "Not Issued" not in s:
                stalls[s] += int(r[ix[s]] or 0)
    print("per-cell instruction mix:", ", ".join(f"{op} {n * 32 / ncells:.0f}" for op, n in byop.most_common(22)))
    fp64 = sum(n for op, n in byop.items() if op in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX")) * 32 / ncells
    print(f"FP64 thread-instructions per cell (DADD+DMUL+DFMA+DSETP+DMNMX): {fp64:.1f}")
    print("stall samples:", ", ".join(f"{k[6:]} {v}" for k, v in stalls.most_common(8)))


if __name__ == "__main__":
    main()
