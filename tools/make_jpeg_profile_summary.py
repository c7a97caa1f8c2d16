#!/usr/bin/env python3
"""profiles/summary_r01_jpeg.md from an `ncu --set full` report of the JPEG decode kernels (tools/bench_mjpg.py run).
usage: make_jpeg_profile_summary.py <report.ncu-rep> [note ...]"""
import csv, subprocess, sys, os
rep = sys.argv[1]
note = " ".join(sys.argv[2:])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
     "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
     "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
     "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(M)], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = hdr.index


def num(r, n):
    v, u = float(r[col(n)].replace(",", "")), units[col(n)]
    return v / 1000 if u == "ns" else v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


md = ["# Round 1: `ncu --set full` on the JPEG luminance decode kernels, one 128-frame batch (config-2 scenes, 4:2:2, quality 75, 114 KB/frame)\n",
      "Capture: `ncu --set full --clock-control none --import-source on -k regex:k_jpeg -s 34 -c 17` around `python tools/bench_mjpg.py "
      "--quality 75 --steps 2 --warmup 2 --lanes 1` (after the same command ran clean without ncu; the .ncu-rep stays in gpurun_out/).  "
      "Durations are the profiler's (serialised, cold cache).\n",
      "| kernel | grid x block | regs | us | DRAM read MB | DRAM write MB | warps active % | issue active % | threads / instr | warp instr (M) | L2 hit % | stalls per issue: long sb / branch / wait / barrier |",
      "|---|---|---|---|---|---|---|---|---|---|---|---|"]
seen, total = 0, 0.0
for r in rows[2:]:
    name = r[col("Kernel Name")]
    name = name[name.index("k_jpeg"):].split("(")[0]
    total += num(r, "gpu__time_duration.sum")
    if name == "k_jpeg_sync":
        seen += 1
        if seen > 3:
            continue
        name += f" (round {seen - 1})"
    md.append(f"| `{name}` | {r[col('launch__grid_size')]} x {r[col('launch__block_size')]} | {r[col('launch__registers_per_thread')]} | "
              f"{num(r, 'gpu__time_duration.sum'):.0f} | {num(r, 'dram__bytes_read.sum') / 1e6:.1f} | {num(r, 'dram__bytes_write.sum') / 1e6:.1f} | "
              f"{num(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | {num(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{num(r, 'smsp__thread_inst_executed_per_inst_executed.ratio'):.1f} | {num(r, 'smsp__inst_executed.sum') / 1e6:.1f} | "
              f"{num(r, 'lts__t_sector_hit_rate.pct'):.0f} | "
              f"{num(r, 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'):.1f} / "
              f"{num(r, 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio'):.1f} / "
              f"{num(r, 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio'):.1f} / "
              f"{num(r, 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio'):.1f} |")
md.append(f"\nSum over the batch's JPEG kernels: {total / 1000:.2f} ms (later `k_jpeg_sync` rounds, a few microseconds each, are not listed).\n")
if note:
    md.append(note + "\n")
open(os.path.join(ROOT, "profiles", "summary_r01_jpeg.md"), "w").write("\n".join(md))
print("\n".join(md))
