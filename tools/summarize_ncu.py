"""Turn gpurun_out/*.ncu-rep + launches.csv into the small, committed summaries under profiles/.
Usage: python tools/summarize_ncu.py <tag>     (reads gpurun_out/<tag>_*.ncu-rep / <tag>_launches.csv)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
summary = {}
traffic = {}
for name in sorted(os.listdir(os.path.join(ROOT, "gpurun_out"))):
    if not (name.startswith(tag + "_prof_") and name.endswith(".ncu-rep")):
        continue
    kern = name[len(tag) + 6:-8]
    raw = subprocess.run(["ncu", "-i", os.path.join(ROOT, "gpurun_out", name), "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in WANT or h == "Kernel Name":
            d[h] = f"{v} {u}".strip()
    summary[kern] = d
    def num(key):
        i = hdr.index(key)
        x = float(vals[i].replace(",", ""))
        u = units[i]
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    traffic[f"k_{kern}_dram_bytes_per_launch"] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    # stall breakdown from the source page
    src = subprocess.run(["ncu", "-i", os.path.join(ROOT, "gpurun_out", name), "--page", "source", "--csv"],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    sh = srows[1]
    reasons = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
    agg = {r: 0 for r in reasons}
    for r in srows[2:]:
        for k in reasons:
            v = r[sh.index(k)]
            if v:
                agg[k] += int(v)
    tot = sum(agg.values()) or 1
    d["warp_stall_samples_pct"] = {k[6:]: round(100.0 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:6]}
launch_csv = os.path.join(ROOT, "gpurun_out", tag + "_launches.csv")
if os.path.exists(launch_csv):
    lines = [l for l in open(launch_csv) if l.startswith('"')]
    per = {}
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0]
        per.setdefault(k, []).append(float(row["Metric Value"].replace(",", "")) * (1e-3 if row["Metric Unit"] in ("ns", "nsecond") else 1.0))
    summary["launch_list_us"] = {k: {"launches": len(v), "median_us": sorted(v)[len(v) // 2]} for k, v in per.items()}
    steady = {k: v["median_us"] for k, v in summary["launch_list_us"].items() if k.startswith("k_") and "build" not in k and "synth" not in k}
    tot = sum(steady.values()) or 1
    summary["share_of_step_pct"] = {k: round(100 * v / tot, 1) for k, v in steady.items()}
    with open(os.path.join(out_dir, tag + "_launches.csv"), "w") as f:
        f.writelines(lines)
with open(os.path.join(out_dir, tag + "_ncu_summary.json"), "w") as f:
    json.dump(summary, f, indent=1)
with open(os.path.join(out_dir, "ncu_traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)
print(json.dumps(summary, indent=1)[:3000])
