"""Summarise an .ncu-rep: headline metrics + warp-stall samples aggregated by SASS opcode (first kernel in the report).
usage: python tools/ncu_ops.py report.ncu-rep"""
import csv, subprocess, sys, io
from collections import defaultdict
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:75s} {rows[1][i]:10s}", [r[i] for r in rows[2:]])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hi[0]]
body = rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]
ci = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = defaultdict(lambda: defaultdict(int))
tot = 0
for r in body:
    if len(r) <= ci["# Samples"]:
        continue
    toks = r[1].split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    n = int(r[ci["# Samples"]])
    tot += n
    agg[op]["n"] += n
    for s in stalls:
        agg[op][s] += int(r[ci[s]])
print("total samples", tot)
for op, d in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:12]:
    print(f"{op:8s} {d['n']:6d} {100 * d['n'] / tot:5.1f}%", {s[6:]: d[s] for s in stalls if d[s] >= 10})
