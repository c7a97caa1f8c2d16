#!/usr/bin/env python3
"""Pretty-prints a bench.py JSON line (file or stdin)."""
import json, sys
txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
d = json.loads(txt.strip().splitlines()[-1])
print(f"value {d['value']:.0f} {d['unit']}  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:.0f}  p50 {d.get('p50_latency_ms', 0):.3f} ms  gpus {d['n_gpus']}")
r = d.get("roofline") or {}
for k in r.get("kernels", []):
    g = k["gbs"]
    print(f"  {k['kernel']:16s} {k['ms']:8.3f} ms  {('%7.0f GB/s' % g) if g else ''}")
if r:
    print(f"  dominant {r['kernel']} frac {r['frac']}, whole-path {r['whole_path']['achieved']:.1f} GB/s")
cb = d.get("cpu_baseline")
if cb: print("  cpu", round(cb["value"], 1), cb["unit"], cb["cores"], "threads")
print("  clocks", d.get("clocks"))
mj = d.get("e2e_mjpg")
if mj and "error" in mj: print("  e2e_mjpg failed:", mj["error"])
elif mj: print(f"  e2e_mjpg {mj['value']:.0f} frames/s  ({mj['jpeg_bytes_per_frame']} B/frame, q{mj['quality']}, decoder {mj['decoder']}, parallel {mj['frames_decoded_by_parallel_kernels']}/{mj.get('frames_per_step_per_gpu', mj.get('frames_per_step'))}, tags {mj['tags_found']}/{mj['tags_present']})")
fe = r.get("front_end") or {}
if fe: print(f"  front end {fe['ms']:.3f} ms  frac {fe['frac']:.3f} (per-kernel bytes)  compulsory frac {fe.get('compulsory', {}).get('frac')}")
print("  step_stats", d.get("step_stats"))
for k, v in (d.get("extra") or {}).items():
    if "error" in v: print(f"  {k}: failed: {v['error']}")
    else: print(f"  {k}: {v['value']:.0f} frames/s  e2e {v['e2e']:.0f}  p50 {v['p50_latency_ms']:.3f} ms  hbm frac {v['whole_path_hbm_frac']:.3f}")
