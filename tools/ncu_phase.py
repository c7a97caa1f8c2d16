#!/usr/bin/env python3
"""Aggregates warp-stall samples and executed instructions of an ncu report by CUDA source line (or line ranges).
usage: ncu_phase.py report.ncu-rep [--fn substr] [--file substr] [lo-hi=name ...]"""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]
args = sys.argv[2:]
fn_f = file_f = None
ranges = []
i = 0
while i < len(args):
    if args[i] == "--fn": fn_f = args[i + 1]; i += 2; continue
    if args[i] == "--file": file_f = args[i + 1]; i += 2; continue
    m = re.match(r"(\d+)-(\d+)=(.*)", args[i])
    if m: ranges.append((int(m.group(1)), int(m.group(2)), m.group(3)))
    i += 1
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = cur_fn = None
hdr = None
acc = collections.defaultdict(lambda: [0, 0, ""])  # (fn, file, line) -> samples, instr, text
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No" and "# Samples" in r: hdr = r; si, ii = r.index("# Samples"), r.index("Instructions Executed"); continue
    if hdr is None or not r[0].strip(): continue
    try: ln, s, n = int(r[0]), int(r[si]), int(r[ii])
    except Exception: continue
    a = acc[(cur_fn, cur_file, ln)]; a[0] += s; a[1] += n; a[2] = r[1]
fns = sorted({k[0] for k in acc})
for fn in fns:
    if fn_f and fn_f not in fn: continue
    items = [(k, v) for k, v in acc.items() if k[0] == fn]
    tot_s = sum(v[0] for _, v in items); tot_i = sum(v[1] for _, v in items)
    print(f"== {fn[:110]}\n   samples {tot_s}  warp-instr {tot_i}")
    if file_f: items = [(k, v) for k, v in items if file_f in k[1]]
    if ranges:
        for lo, hi, nm in ranges:
            s = sum(v[0] for k, v in items if lo <= k[2] <= hi); n = sum(v[1] for k, v in items if lo <= k[2] <= hi)
            print(f"  {nm:30s} {lo:4d}-{hi:4d}  samples {100*s/max(tot_s,1):5.1f}%  instr {100*n/max(tot_i,1):5.1f}%")
        s = sum(v[0] for k, v in items if not any(lo <= k[2] <= hi for lo, hi, _ in ranges))
        print(f"  {'(other lines / other files)':30s}            samples {100*(tot_s - sum(v[0] for k, v in items) + s)/max(tot_s,1):5.1f}%")
    else:
        for k, v in sorted(items, key=lambda kv: -kv[1][0])[:40]:
            print(f"  {k[1].split('/')[-1][:18]:18s}:{k[2]:4d} {100*v[0]/max(tot_s,1):5.1f}% inst {100*v[1]/max(tot_i,1):5.1f}%  {v[2].strip()[:90]}")
