"""Throughput of large batches on one GPU (BASELINE configs 3-5 in spirit): W x H image, tile mode.
Usage: python tools/throughput.py [width height bits batch_tiles reps]
HYDRIUM_B200_CHAIN=latency|throughput picks the chain kernel (default: by batch size)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydrium_b200.engine import Engine
W = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
H = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
bits = int(sys.argv[3]) if len(sys.argv) > 3 else 8
T = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
item = bits // 8
mode = os.environ.get("HYDRIUM_B200_CHAIN", "auto")
with Engine(device=0, max_batch_tiles=T) as eng:
    d_in = eng.device_alloc(W * H * 3 * item)
    cap = W * H * 2 + (1 << 20)
    d_out = eng.device_alloc(cap)
    eng.synth_fill(d_in, W, H, bits=bits, seed=0)
    kw = dict(sample_fmt=0 if bits == 8 else 1, linear_light=int(bits == 16), d_out=d_out, d_out_cap=cap)
    for r in range(reps + 1):
        t0 = time.perf_counter()
        n = eng.encode_image_device(d_in, W, H, 3, **kw)
        dt = time.perf_counter() - t0
        print(f"[{mode}] {W}x{H} u{bits} T={T}: {n} bytes in {dt * 1e3:.1f} ms -> {W * H / dt / 1e6:.0f} Mpx/s ({n * 8 / (W * H):.2f} bpp)", flush=True)
    eng.enable_timing(True)
    eng.stage_ms()
    eng.encode_image_device(d_in, W, H, 3, **kw)
    print(f"[{mode}] stage_ms", {k: round(v, 3) for k, v in eng.stage_ms().items()}, flush=True)
