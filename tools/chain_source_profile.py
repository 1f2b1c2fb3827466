"""Per-instruction stall samples of k_ans_chain's straight-line step, from an ncu report captured with
--set full --import-source on.  Usage: python tools/chain_source_profile.py <report.ncu-rep> > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
cnt = [int(r[ix["Instructions Executed"]]) for r in data]
total = sum(int(r[ix["# Samples"]]) for r in data)
# the unrolled batch is the largest group of instructions sharing one execution count
mode = max((c for c in Counter(cnt).items() if c[0] > 1000), key=lambda kv: kv[1])[0]
idx = [i for i, c in enumerate(cnt) if abs(c - mode) <= mode // 20]
reg = sum(int(data[i][ix["# Samples"]]) for i in idx)
print(f"report {rep}")
print(f"kernel samples {total}; straight-line chain region: {len(idx)} instructions, {reg} samples ({100 * reg / total:.1f} %)")
agg = {}
for i in idx:
    r = data[i]
    op = r[ix["Source"]].split()[0]
    if op.startswith("@"):
        op = r[ix["Source"]].split()[1]
    a = agg.setdefault(op, [0, 0, Counter()])
    a[0] += 1
    a[1] += int(r[ix["# Samples"]])
    for sc in stalls:
        v = int(r[ix[sc]])
        if v:
            a[2][sc[6:]] += v
print("\nby opcode (samples, share of region, top stall reasons)")
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"  {op:24s} n={a[0]:4d} {a[1]:7d} {100 * a[1] / reg:5.1f}%  {dict(a[2].most_common(3))}")
# one step, instruction by instruction
j = idx[len(idx) // 2]
while "LDS.U16" not in data[j][ix["Source"]]:
    j += 1
k = j + 1
while "LDS.U16" not in data[k][ix["Source"]]:
    k += 1
base = min(int(data[i][ix["stall_selected"]]) for i in range(j, k + 1) if int(data[i][ix["stall_selected"]]) > 0)
print(f"\none step (samples per instruction; ~{base} samples = one issue cycle)")
for r in data[j:k + 1]:
    s = int(r[ix["# Samples"]])
    why = {sc[6:]: int(r[ix[sc]]) for sc in stalls if int(r[ix[sc]])}
    print(f"  {r[ix['Source']].strip()[:56]:56s} {s:5d}  ~{s / base:4.1f} cyc  {why}")
print(f"  step total ~{sum(int(r[ix['# Samples']]) for r in data[j:k]) / base:.1f} cycles")
