import os, sys
sys.path.insert(0, "/root/repo")
from hydrium_b200 import engine as E
W = H = 4096
with E.Engine(device=0, max_batch_tiles=256) as eng:
    d_in = eng.device_alloc(W * H * 3); cap = E.output_bound(W, H); d_out = eng.device_alloc(cap)
    eng.synth_fill(d_in, W, H, bits=8, seed=0)
    for i in range(3):
        print("run", i, file=sys.stderr)
        eng.encode_image_device(d_in, W, H, 3, d_out=d_out, d_out_cap=cap)
