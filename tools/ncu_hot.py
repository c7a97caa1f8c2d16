#!/usr/bin/env python3
"""Aggregates an ncu report's source page by CUDA source line: warp-stall samples and instruction counts."""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and any("Sampling" in c for c in r):
        hdr = r; rows = rows[i+1:]; break
if hdr is None:
    print(out[:2000]); sys.exit(1)
def col(name):
    for i, c in enumerate(hdr):
        if c.strip() == name: return i
    for i, c in enumerate(hdr):
        if name in c: return i
    return None
ci_src = col("Source"); ci_samp = col("# Samples") if col("# Samples") is not None else col("Warp Stall Sampling (All Samples)")
ci_inst = col("Instructions Executed")
print("columns:", [c for c in hdr][:12])
agg = collections.Counter(); inst = collections.Counter()
tot = 0
for r in rows:
    if len(r) <= max(ci_src, ci_samp): continue
    try: s = int(r[ci_samp].replace(",", ""))
    except: continue
    agg[r[ci_src].strip()[:110]] += s; tot += s
    if ci_inst is not None:
        try: inst[r[ci_src].strip()[:110]] += int(r[ci_inst].replace(",", ""))
        except: pass
print("total samples", tot)
for k, v in agg.most_common(top):
    print(f"{v:8d} {100.0*v/max(tot,1):5.1f}%  inst={inst.get(k,0):>10d}  {k}")
