#!/usr/bin/env python3
"""Builds profiles/summary_<tag>.md and profiles/traffic_<round>.json from gpurun_out artefacts.

usage: make_profile_summary.py <tag> <bench.json> <launches.csv> <skip launches> <steps in list> [ncu-rep ...]
The launch list is the `ncu --metrics gpu__time_duration.sum` pass of the same bench command; its per-launch
times are cold-cache and serialised, so only each kernel's SHARE of the step is compared with the CUDA-event
times bench.py measured.
"""
import collections, csv, json, os, re, subprocess, sys

tag, bench_path, launches_path = sys.argv[1:4]
reps = sys.argv[4:]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
line = json.loads(open(bench_path).read().strip().splitlines()[-1])
kern = line["roofline"]["kernels"]
peak = line["roofline"]["peak"]
B = line["config"].get("frames_per_batch", line["config"]["frames_per_step_per_gpu"])

# launch list -> per kernel mean device time, only launches of full batches (grid z or y == B is not visible here,
# so take the launches of the LAST complete step: the final `len(kern)` pipeline launches before the latency runs)
rows = [r for r in csv.reader(open(launches_path)) if len(r) > 14 and r[0].isdigit()]
names = [re.sub(r"\(.*", "", r[4]).replace("void ", "").split("<")[0] for r in rows]
times = [float(r[14]) for r in rows]
alias = {"k_pre_yuyv_dec2": "pre_yuyv_dec2", "k_pre_gray_dec2": "pre_gray_dec2", "k_pre_bgr_dec2": "pre_bgr_dec2", "k_pre_bgr_dec1": "pre_bgr_dec1",
         "k_pre_generic": "pre_generic", "k_blur": "blur", "k_tile_minmax": "tile_minmax", "k_ccl_local": "ccl_local", "k_ccl_merge": "ccl_merge",
         "k_ccl_handoff": "ccl_handoff", "k_ccl_final": "ccl_final", "k_boundary": "boundary", "k_select": "select", "k_scatter": "scatter",
         "k_fit_small": "fit_small", "k_decode": "decode", "k_quads": "quads"}
# batched launches are the slow ones: per kernel name take the maximum-duration launches' median
per = collections.defaultdict(list)
fit_cta_seen = collections.defaultdict(int)
for r, n, t in zip(rows, names, times):
    key = alias.get(n)
    if n == "k_fit_cta":
        key = "fit_medium" if "(128, 1, 1)" in r[7] else ("fit_huge" if "(512, 1, 1)" in r[7] else "fit_large")
    if key: per[key].append(t)
ncu_ms = {}
for k, v in per.items():
    v = sorted(v, reverse=True)
    top = v[: max(1, len(v) // 20)] if len(v) > 40 else v[:3]   # the batched launches (single-frame latency runs are tiny)
    ncu_ms[k] = sum(top) / len(top) / 1e6
tot_ev = sum(k["ms"] for k in kern)
tot_ncu = sum(ncu_ms.get(k["kernel"], 0.0) for k in kern)

out = []
out.append(f"# {tag}: {line['config']['workload']}\n")
out.append(f"bench.py, B200, {B} frames per batch: **{line['value']:.0f} frames/s device-resident, "
           f"{line['e2e']['value']:.0f} frames/s end to end, p50 single-frame latency {line.get('p50_latency_ms', 0):.3f} ms**"
           + (f", CPU classic detector (port) {line['cpu_baseline']['value']:.0f} frames/s on {line['cpu_baseline']['cores']} threads" if line.get("cpu_baseline") and "value" in line["cpu_baseline"] else "")
           + (f", reference kernels recompiled for sm_100a (decode excluded) {line['reference_gpu']['value']:.0f} frames/s" if line.get("reference_gpu") and "value" in line["reference_gpu"] else "")
           + ".\n")
out.append("Per-kernel CUDA-event time per step (bench.py `roofline.kernels`) vs the ncu launch list "
           f"(`{os.path.basename(launches_path)}`, share of summed kernel time; absolute ncu times are cold-cache and serialised):\n")
out.append("| kernel | ms / step (events) | share (events) | ms (ncu list) | share (ncu list) | algorithmic GB/s | of measured HBM peak |")
out.append("|---|---|---|---|---|---|---|")
for k in kern:
    n = k["kernel"]
    g = k["gbs"]
    out.append(f"| {n} | {k['ms']:.3f} | {100 * k['ms'] / tot_ev:.1f} % | {ncu_ms.get(n, float('nan')):.3f} | "
               f"{100 * ncu_ms.get(n, 0) / tot_ncu if tot_ncu else 0:.1f} % | {('%.0f' % g) if g else '-'} | {('%.1f %%' % (100 * g / peak)) if g else '-'} |")
out.append("")
traffic = {}
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units = rr[0], rr[1]
    for last in rr[2:]:
        def g(n):
            return last[hdr.index(n)] if n in hdr else None
        def gb(n):  # bytes
            v, u = float(g(n).replace(",", "")), units[hdr.index(n)]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        name = re.sub(r"\(.*", "", g("Kernel Name")).replace("void ", "")
        block = g("launch__block_size")
        key = alias.get(name.split("<")[0], name)
        if name.startswith("k_fit_cta"): key = "fit_medium" if block == "128" else ("fit_huge" if block == "512" else "fit_large")
        rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
        traffic[key] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "report": os.path.basename(rep),
                        "issue_active_pct": float(g("smsp__issue_active.avg.pct_of_peak_sustained_active")),
                        "warps_active_pct": float(g("sm__warps_active.avg.pct_of_peak_sustained_active")),
                        "duration_us": float(g("gpu__time_duration.sum")), "frames_per_launch": B, "workload": line["config"]["workload"]}
        st = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try: st.append((int(last[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except Exception: pass
        tot = sum(v for v, _ in st) or 1
        alg = next((k["alg_bytes"] for k in kern if k["kernel"] == key), None)
        out.append(f"`ncu --set full` on **{name}** ({os.path.basename(rep)}): duration {g('gpu__time_duration.sum')} {units[hdr.index('gpu__time_duration.sum')]}, "
                   f"{g('launch__registers_per_thread')} registers/thread, CTA limits regs/smem/warps = {g('launch__occupancy_limit_registers')}/"
                   f"{g('launch__occupancy_limit_shared_mem')}/{g('launch__occupancy_limit_warps')}, warps active "
                   f"{float(g('sm__warps_active.avg.pct_of_peak_sustained_active')):.1f} %, issue active "
                   f"{float(g('smsp__issue_active.avg.pct_of_peak_sustained_active')):.1f} %, DRAM throughput "
                   f"{float(g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):.1f} % of peak; "
                   f"dram read {rd / 1e6:.1f} MB + write {wr / 1e6:.1f} MB per launch"
                   + (f" (algorithmic {alg / 1e6:.1f} MB)" if alg else "") + "; stalls: "
                   + ", ".join(f"{n} {100 * v / tot:.0f} %" for v, n in sorted(st, reverse=True)[:6]) + ".\n")
open(os.path.join(ROOT, "profiles", f"summary_{tag}.md"), "w").write("\n".join(out) + "\n")
tpath = os.path.join(ROOT, "profiles", os.environ.get("TRAFFIC_JSON", "traffic.json"))
old = json.load(open(tpath)) if os.path.exists(tpath) else {}
for k, v in traffic.items():
    v["tag"] = tag
    old[k] = v
json.dump(old, open(tpath, "w"), indent=1, sort_keys=True)
print("\n".join(out))
