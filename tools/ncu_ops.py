"""Aggregate an ncu source page (ncu -i X.ncu-rep --page source --csv) by SASS opcode and by source-line range:
warp-stall samples and executed instructions.  usage: python tools/ncu_ops.py report.ncu-rep [top_n]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, v = r[0], (r[2] if len(r) > 2 else r[1])
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
for a, b in zip(h, v):
    if a in want or ("issue_stalled" in a and a.endswith("per_issue_active.ratio") and float(b or 0) > 0.3):
        print(f"{a} = {b}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))[2:]
ops = collections.defaultdict(lambda: [0, 0, 0])
tot = 0
for row in rows:
    if len(row) < 6:
        continue
    sass = row[1].strip()
    parts = sass.split()
    op = parts[1] if sass.startswith("@") else parts[0]
    base = ".".join(op.rstrip(";").split(".")[:2])
    s, ex = int(row[2] or 0), int(row[5] or 0)
    ops[base][0] += s
    ops[base][1] += ex
    ops[base][2] += 1
    tot += s
print("total samples", tot)
for k, v in sorted(ops.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k:28s} samples {v[0]:7d} ({100 * v[0] / max(tot, 1):5.1f}%)  executed {v[1]:10d}  sites {v[2]}")
