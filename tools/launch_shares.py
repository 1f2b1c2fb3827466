"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python tools/launch_shares.py gpurun_out/<list>.csv [> profiles/<name>.txt]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)   # -> microseconds
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
total = sum(a[1] for a in agg.values())
print(f"{sys.argv[1]}: {sum(a[0] for a in agg.values())} launches, {total / 1e3:.3f} ms "
      "(serialised, cold-cache per-launch times: read the shares, not the absolutes)")
for name, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{name:28s} launches {n:4d}  total {t / 1e3:9.3f} ms  {100 * t / total:5.1f}%  per launch {t / n / 1e3:8.3f} ms")
