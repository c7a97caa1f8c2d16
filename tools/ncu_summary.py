#!/usr/bin/env python3
"""One-screen summary of an ncu report: duration, occupancy, DRAM traffic, stall breakdown, hot lines."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr = rows[0]; last = rows[-1]
def g(n):
    return last[hdr.index(n)] if n in hdr else None
print("kernel", g("Kernel Name"), "grid", g("launch__grid_size"), "block", g("launch__block_size"))
for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__registers_per_thread",
          "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]:
    if k in hdr: print(f"  {k:62s} {g(k)} {rows[1][hdr.index(k)] if len(rows)>2 else ''}")
st = []
for i, h in enumerate(hdr):
    if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
        try: st.append((int(last[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        except: pass
tot = sum(v for v, _ in st) or 1
print("  stalls:", ", ".join(f"{n} {100*v/tot:.0f}%" for v, n in sorted(st, reverse=True)[:7]))
