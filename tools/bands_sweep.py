"""Wall-clock step of config 2 (device-resident) for HYDRIUM_B200_BANDS=1..4: how many bands the image-level
encode cuts the tile rows into (engine.cu::launch_bands)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hydrium_b200 import engine as E
W = H = 4096
with E.Engine(device=0, max_batch_tiles=256) as eng:
    d_in = eng.device_alloc(W * H * 3); cap = E.output_bound(W, H); d_out = eng.device_alloc(cap)
    eng.synth_fill(d_in, W, H, bits=8, seed=0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(13):
        flush.fill_(i); torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.encode_image_device(d_in, W, H, 3, d_out=d_out, d_out_cap=cap)
        ts.append(1e3 * (time.perf_counter() - t0))
    ts = sorted(ts[3:])
    print(os.environ.get("HYDRIUM_B200_BANDS"), "bands: median %.3f best %.3f ms (wall)" % (ts[len(ts)//2], ts[0]))
