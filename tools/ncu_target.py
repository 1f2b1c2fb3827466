"""Minimal workload for ncu: BASELINE configs[1] (4096x4096 sRGB8, 256 tiles) encoded N times.
Usage: ncu ... python tools/ncu_target.py [iterations]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydrium_b200.engine import Engine, output_bound  # noqa: E402

W = H = 4096
it = int(sys.argv[1]) if len(sys.argv) > 1 else 3
with Engine(device=0, max_batch_tiles=256) as eng:
    d_in = eng.device_alloc(W * H * 3)
    cap = output_bound(W, H)
    d_out = eng.device_alloc(cap)
    eng.synth_fill(d_in, W, H, bits=8, seed=0)
    eng.enable_timing(True)   # plain single-stream sequence: one launch of every kernel per image, as bench.py times them
    for _ in range(it):
        n = eng.encode_image_device(d_in, W, H, 3, d_out=d_out, d_out_cap=cap)
    print("bytes", n)
