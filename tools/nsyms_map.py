"""Symbols per tile of BASELINE config 2, as a 16x16 map (thousands) -- shows where the long rANS chains sit."""
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
    ns = np.array([int(eng.read_tap(E.TAP_NSYMS, t, np.uint32)[0]) for t in range(256)]).reshape(16, 16)
    np.set_printoptions(linewidth=200)
    print((ns // 1000))
    print("row max (k):", (ns.max(axis=1) // 1000))
