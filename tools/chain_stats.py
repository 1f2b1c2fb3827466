"""Per-tile statistics of the rANS chain kernels on a launch that fills every SM several times
(a band of the 65536-wide image): cycles per symbol, prologue cycles, how the chain warps were spread
over SMs and sub-partitions, per-kernel device time.  usage: chain_stats.py [table|compact] [tiles]"""
import os
import sys
from collections import Counter

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydrium_b200 import engine as E

mode = {"table": 1, "compact": 2, "auto": 0}[sys.argv[1] if len(sys.argv) > 1 else "compact"]
tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
taps = "--no-taps" not in sys.argv
W = 65536
rows = max(1, tiles // 256)
H = rows * 256
T = rows * 256
with E.Engine(device=0, max_batch_tiles=T) as eng:
    eng.set_chain_kernel(mode)
    d_in = eng.device_alloc(W * H * 3)
    cap = E.output_bound(W, H)
    d_out = eng.device_alloc(cap)
    eng.synth_fill(d_in, W, H, bits=8, seed=0, full_width=65536, full_height=65536)
    if taps:
        eng.enable_taps(True)
    eng.enable_timing(True)
    for _ in range(2):
        eng.stage_ms()
        n = eng.encode_image_device(d_in, 65536, 65536, 3, tile_rows=(0, rows), d_out=d_out, d_out_cap=cap)
    st = eng.stage_ms()
    print(f"mode {sys.argv[1] if len(sys.argv) > 1 else 'compact'}: {T} tiles, {n} bytes; stage ms {({k: round(v, 3) for k, v in st.items()})}")
    if taps:
        ns = np.array([int(eng.read_tap(E.TAP_NSYMS, t, np.uint32)[0]) for t in range(T)])
        clk = np.array([eng.read_tap(E.TAP_CLK, t, np.uint32) for t in range(T)]).astype(np.int64)
        cps = clk[:, 1] / ns
        print("symbols/tile min/mean/max", ns.min(), round(ns.mean()), ns.max(), "total", ns.sum())
        print("prologue cycles min/mean/max", clk[:, 0].min(), round(clk[:, 0].mean()), clk[:, 0].max())
        print("chain cycles/symbol min/mean/max", round(cps.min(), 1), round(cps.mean(), 1), round(cps.max(), 1))
        chain_ms = st["ans_chain"]
        print(f"aggregate: {ns.sum() / chain_ms / 1e6:.2f} Gsym/s over the kernel's {chain_ms:.2f} ms; "
              f"sum of chain cycles / (kernel cycles at 1.965 GHz) = {clk[:, 1].sum() / (chain_ms * 1.965e6):.1f} chains in flight on average")
        sm = Counter(int(x) for x in clk[:, 2])
        print("tiles per SM min/max", min(sm.values()), max(sm.values()), "SMs used", len(sm))
        if mode == 2:
            sub = Counter(int(x) >> 8 for x in clk[:, 3])
            print("chain warps per sub-partition (all SMs):", dict(sorted(sub.items())), "chain = warp 1 in", int((clk[:, 3] & 1).sum()), "tiles")
