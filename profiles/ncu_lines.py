#!/usr/bin/env python
"""Per-source-line and per-opcode attribution of one kernel in an ncu report (needs -lineinfo):

    python profiles/ncu_lines.py gpurun_out/x.ncu-rep [top_n]
"""
import collections
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
col = {h: i for i, h in enumerate(rows[0])}
for k in ("gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum"):
    if k in col:
        print(f"{k:70s} {rows[1][col[k]]:10s} {rows[2][col[k]]}")
for k in sorted(col):
    if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k and I(rows[2][col[k]].split(".")[0]) > 2000:
        print(f"{k:70s} {rows[2][col[k]]}")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
lines, ops = [], collections.Counter()
cur = None
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) >= 8 and r[0].isdigit():
        lines.append((cur, int(r[0]), r[1], I(r[7]), I(r[6])))
    elif len(r) >= 8 and r[0] == "" and r[2].startswith("0x"):
        toks = r[3].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        ops[op] += I(r[7])
tot = sum(x[3] for x in lines)
print("instructions attributed:", tot)
for f, l, src, n, s in sorted(lines, key=lambda o: -o[3])[:top]:
    print(f"{f}:{l:4d} {100 * n / tot:5.1f}% samples {s:6d}  {src.strip()[:110]}")
t2 = sum(ops.values())
print("opcodes:", "  ".join(f"{o} {100 * n / t2:.1f}%" for o, n in ops.most_common(16)))
