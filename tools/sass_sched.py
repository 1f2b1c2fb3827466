"""tools/sass_sched.py -- decode the scheduling control fields of a kernel's SASS (stall count, yield, write /
read scoreboard, wait mask) and run the single-warp issue model of /opt/skills/guides/B300_MICROARCH.md over an
address range: a local estimate of the cycles one warp needs per rANS chain step, before spending GPU time.

    python tools/sass_sched.py OBJ KERNEL_SUBSTRING [--from ADDR --to ADDR] [--lds 29] [--marker LDS.U16]
"""
import argparse
import re
import subprocess


def load(obj, kernel):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    out, on, cur = [], False, None
    for line in txt.splitlines():
        if "Function :" in line:
            on = kernel in line
            continue
        if not on:
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", line)
        if m:
            cur = {"addr": int(m.group(1), 16), "text": " ".join(m.group(2).split()), "lo": int(m.group(3), 16)}
            continue
        m = re.match(r"\s*/\* (0x[0-9a-f]+) \*/", line)
        if m and cur:
            hi = int(m.group(1), 16)
            cur.update(stall=(hi >> 41) & 0xF, yld=(hi >> 45) & 1, wbar=(hi >> 46) & 7, rbar=(hi >> 49) & 7,
                       wait=(hi >> 52) & 0x3F)
            out.append(cur)
            cur = None
    return out


FMA_OPS = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2")
ALU_OPS = ("IADD3", "VIADD", "LOP3", "SHF", "SEL", "ISETP", "PRMT", "LEA", "MOV", "IABS", "FLO", "POPC", "VIMNMX", "IMNMX", "FSEL", "PLOP3")
LSU_OPS = ("LDS", "STS", "LDG", "STG", "ATOMS", "ATOMG", "RED", "LDSM")


def simulate(ins, lat, occ=None):
    """occ: pipe occupancy per warp instruction {"fma": 2, "wide": 4, "alu": 2, "lsu": 2, "lsu128": 4} or None"""
    t, sb = 0, [0] * 6
    busy = {"fma": 0, "alu": 0, "lsu": 0}
    times = []
    for i in ins:
        arm = max([sb[s] for s in range(6) if i["wait"] >> s & 1] or [0])
        t = max(t, arm)
        op = i["text"].split()[0] if not i["text"].startswith("@") else i["text"].split()[1]
        if occ:
            pipe, n = None, 0
            if op.startswith(FMA_OPS):
                pipe, n = "fma", occ["wide"] if ".WIDE" in op else (occ.get("hi", occ["wide"]) if ".HI" in op else occ["fma"])
            elif op.startswith(ALU_OPS):
                pipe, n = "alu", occ["alu"]
            elif op.startswith(LSU_OPS):
                pipe, n = "lsu", occ["lsu128"] if ".128" in op else (occ.get("sts", occ["lsu"]) if op.startswith("STS") else occ["lsu"])
            if pipe:
                t = max(t, busy[pipe])
                busy[pipe] = t + n
        times.append(t)
        l = next((v for k, v in lat.items() if op.startswith(k)), 20)
        if i["wbar"] < 6:
            sb[i["wbar"]] = max(sb[i["wbar"]], t + l)
        if i["rbar"] < 6:
            sb[i["rbar"]] = max(sb[i["rbar"]], t + 8)
        t += max(1, i["stall"])
    return times


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj")
    ap.add_argument("kernel")
    ap.add_argument("--from", dest="lo", default=None)
    ap.add_argument("--to", dest="hi", default=None)
    ap.add_argument("--lds", type=int, default=29)
    ap.add_argument("--marker", default=r"LDS\.U16 R\d+, \[R\d+\]$")
    ap.add_argument("--show", action="store_true")
    ap.add_argument("--occ", default=None, help="pipe occupancy fma,wide,alu,lsu,lsu128[,hi,sts] e.g. 2.6,8,1.5,2.5,7.8,6.5,4.5 (profiles/r01_ubench_issue_throughput.txt)")
    a = ap.parse_args()
    ins = load(a.obj, a.kernel)
    if a.lo:
        ins = [i for i in ins if int(a.lo, 16) <= i["addr"] <= int(a.hi, 16)]
    lat = {"LDS": a.lds, "STS": 8, "LDG": 400, "CREDUX": 30, "REDUX": 30}
    occ = dict(zip(("fma", "wide", "alu", "lsu", "lsu128", "hi", "sts"), map(float, a.occ.split(",")))) if a.occ else None
    times = simulate(ins, lat, occ)
    marks = [t for i, t in zip(ins, times) if re.search(a.marker, i["text"])]
    if a.show:
        for i, t in zip(ins, times):
            print(f"{t:8.1f} /*{i['addr']:04x}*/ st={i['stall']:2d} y={i['yld']} w={i['wbar']} r={i['rbar']} wait={i['wait']:02x}  {i['text']}")
    if len(marks) > 1:
        d = [b - a_ for a_, b in zip(marks, marks[1:])]
        d = [x for x in d if x < 4 * sorted(d)[len(d) // 2]]   # leave out the jumps between unrolled regions
        print(f"{len(marks)} markers, mean distance {sum(d) / len(d):.2f} cycles (min {min(d)}, max {max(d)}), {len(ins) / len(marks):.1f} instructions per marker")


if __name__ == "__main__":
    main()
