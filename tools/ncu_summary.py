"""Summarise an .ncu-rep: headline metrics, stall reasons and the hottest SASS blocks. usage: ncu_summary.py rep [kernel-index]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]
want = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:70s} {vals[i]} {units[i]}")
print("-- stalls (warps per issue-active cycle)")
for i, h in enumerate(hdr):
    if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h:
        try:
            v = float(vals[i])
        except ValueError:
            continue
        if v > 0.05:
            print(f"  {v:6.2f} {h.split('stalled_')[1].split('_per_issue')[0]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))[2:]
rows = [r for r in rows if len(r) > 5 and r[5].isdigit()]
tot = sum(int(r[5]) for r in rows)
print(f"-- SASS: {len(rows)} instructions, {tot} warp-level executions")
blocks, cur = [], None
for r in rows:
    c = int(r[5])
    if cur is None or c != cur[0]:
        cur = [c, 0, [], 0]
        blocks.append(cur)
    cur[1] += 1
    cur[2].append(r[1].strip())
    cur[3] += int(r[4])
for b in sorted(blocks, key=lambda b: -b[0] * b[1])[:12]:
    ops = collections.Counter((x.split()[1] if x.startswith("@") else x.split()[0]) for x in b[2])
    print(f"  exec {b[0]:>10} x {b[1]:>4} = {b[0] * b[1] / tot * 100:5.1f}%  samples {b[3]:>6}  {dict(ops.most_common(7))}")
