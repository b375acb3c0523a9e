#!/usr/bin/env python
"""Per-source-line executed warp instructions and stall samples of one kernel from an .ncu-rep (needs -lineinfo).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_gray_blur5 [top]
"""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, agg, src = None, collections.defaultdict(lambda: [0, 0]), {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        cur_file = r[1].split("/")[-1]
    if len(r) < 8 or not r[0].isdigit():
        continue
    key = (cur_file, int(r[0]))
    src.setdefault(key, r[1].strip())
    try:
        agg[key][0] += int(r[7]); agg[key][1] += int(r[4])
    except ValueError:
        pass
tot = sum(v[0] for v in agg.values()) or 1
print(f"{kern}: {tot} warp instructions (all captured launches)")
for key, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100.0 * n / tot:5.1f}%  stall {s:6d}  {key[0]}:{key[1]:<4d} {src[key][:110]}")
