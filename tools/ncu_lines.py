#!/usr/bin/env python3
"""Per-source-line instruction and stall-sample totals of one kernel from an ncu report captured with --import-source on.

usage: ncu_lines.py <report.ncu-rep> <kernel regex> [launch skip] [top N]
(reads `ncu --page source --csv --print-source cuda,sass`; rows of SASS follow the source line they belong to)
"""
import csv, io, subprocess, sys, collections, os
rep, rx = sys.argv[1:3]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{rx}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
fname, line, text = "", 0, ""
inst = collections.Counter(); samp = collections.Counter(); src = {}
hdr = None
for row in csv.reader(io.StringIO(out)):
    if len(row) == 2 and row[0] == "File Name":
        fname = os.path.basename(row[1]); continue
    if row and row[0] == "Line No":
        hdr = row; continue
    if hdr is None or len(row) < 8: continue
    if row[0].isdigit():
        line = int(row[0]); src[(fname, line)] = row[1]
    if row[2].startswith("0x"):
        try:
            inst[(fname, line)] += int(row[7]); samp[(fname, line)] += int(row[6])
        except ValueError:
            pass
ti, ts = sum(inst.values()), sum(samp.values())
print(f"total warp instructions {ti}, samples {ts}")
byfile = collections.Counter()
for (f, l), v in inst.items(): byfile[f] += v
print("by file:", [(f, round(100 * v / ti, 1)) for f, v in byfile.most_common(8)])
for (f, l), v in inst.most_common(top):
    print(f"{100 * v / ti:5.1f}% inst {100 * samp[(f, l)] / max(ts, 1):5.1f}% samp  {f}:{l}  {src.get((f, l), '')[:110].strip()}")
