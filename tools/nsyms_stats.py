import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydrium_b200 import engine as E
W = H = 4096
with E.Engine(device=0, max_batch_tiles=256) as eng:
    d_in = eng.device_alloc(W * H * 3); cap = E.output_bound(W, H); d_out = eng.device_alloc(cap)
    eng.synth_fill(d_in, W, H, bits=8, seed=0)
    eng.enable_taps(True)
    eng.encode_image_device(d_in, W, H, 3, d_out=d_out, d_out_cap=cap)
    ns = np.array([int(eng.read_tap(E.TAP_NSYMS, t, np.uint32)[0]) for t in range(256)])
    print("nsyms min/mean/max", ns.min(), ns.mean(), ns.max(), "sum", ns.sum())
    print(np.sort(ns)[-10:])
    clk = np.array([eng.read_tap(E.TAP_CLK, t, np.uint32) for t in range(256)])
    print("prologue cycles min/mean/max", clk[:, 0].min(), clk[:, 0].mean(), clk[:, 0].max())
    cps = clk[:, 1] / ns
    print("chain cycles/step min/mean/max", cps.min(), cps.mean(), cps.max())
    # co-residency: tiles sharing an SM
    from collections import defaultdict
    by = defaultdict(list)
    for t in range(256):
        by[int(clk[t, 2])].append((t, int(clk[t, 3]), float(cps[t])))
    solo = [v[0][2] for v in by.values() if len(v) == 1]
    pair = [x[2] for v in by.values() if len(v) == 2 for x in v]
    print("SMs", len(by), "solo tiles", len(solo), "mean c/step", np.mean(solo) if solo else None, "paired tiles", len(pair), "mean c/step", np.mean(pair) if pair else None)
    print("examples", list(by.items())[:4])
