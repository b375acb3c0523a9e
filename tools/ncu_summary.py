#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): per-kernel headline metrics, SASS opcode mix weighted by
executed count, and the instructions with the most stall samples.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex] > profiles/xxx.txt
"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
       "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
       "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "launch__occupancy_limit_registers",
       "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else None
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    rows = [r for r in rows if r and not r[0].startswith("==")]
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    seen = set()
    for r in rows[2:]:
        name = r[kn].split("(")[0]
        if pat and not re.search(pat, name):
            continue
        if name in seen:
            continue
        seen.add(name)
        print(f"=== {name}  (launch id {r[0]})")
        for m in RAW:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:75s} {r[i]:>18s} {units[i]}")
    for name in sorted(seen):
        src = run(["-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + name.split("<")[0].split("::")[-1].split()[-1], "--launch-count", "1"]) if False else \
            run(["-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + name.split("<")[0].split("::")[-1].split()[-1]])
        lines = [l for l in src.splitlines() if l and not l.startswith("==")]
        recs = list(csv.reader(io.StringIO("\n".join(lines))))
        try:
            h = next(i for i, r in enumerate(recs) if r and r[0] == "Address")
        except StopIteration:
            continue
        cols = recs[h]
        ci, cs, ce = cols.index("Source"), cols.index("Warp Stall Sampling (All Samples)"), cols.index("Instructions Executed")
        ops, stalls, total = Counter(), [], 0
        for r in recs[h + 1:]:
            if len(r) <= max(ci, cs, ce) or r[0] == "Address":
                break                                   # first launch only
            try:
                n = int(r[ce]); s = int(r[cs])
            except ValueError:
                continue
            sass = r[ci].strip()
            op = re.sub(r"^@!?U?P\d+\s+", "", sass).split()[0] if sass else "?"
            ops[op.split(".")[0]] += n
            total += n
            stalls.append((s, sass))
        print(f"=== {name}: SASS opcode mix (warp-level executed, first launch), total {total}")
        for op, n in ops.most_common(22):
            print(f"  {op:12s} {n:12d} {100.0 * n / max(1, total):6.2f}%")
        print("  -- top stall-sample instructions")
        for s, sass in sorted(stalls, reverse=True)[:14]:
            print(f"  {s:7d}  {sass}")


if __name__ == "__main__":
    main()
