"""Per-GPU shares of BASELINE configs 3, 4 and 5 on ONE B200, device-resident, for several engine batch
sizes (tiles in flight per launch).  What an 8-GPU run gives each rank:
  config 3: 16384x16384 16-bit linear -> 512 tiles (8 tile rows of 64), packed RGB16
  config 4: 1024 frames of 1920x1080 sRGB8 -> 128 frames = 5120 tiles, each frame its own codestream
  config 5: 65536x65536 sRGB8 -> 8192 tiles (32 tile rows of 256), level-10 container on rank 0
Times are wall clock around the synchronous C-ABI calls (descriptors prebuilt), best of 3."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydrium_b200 import engine as E
from hydrium_b200.abi import HYD_UINT8, HYD_UINT16
from hydrium_b200.lib import HydbTile


def best_of(fn, reps=3):
    fn()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def config3(eng):
    w, rows = 16384, 8
    n_in = w * rows * 256 * 3 * 2
    d_in = eng.device_alloc(n_in)
    eng.synth_fill(d_in, w, rows * 256, bits=16, x0=0, y0=0, full_width=16384, full_height=16384)
    cap = 2 * E.output_bound(w, rows * 256)
    d_out = eng.device_alloc(cap)
    dt, n = best_of(lambda: eng.encode_image_device(d_in, 16384, 16384, 3, sample_fmt=HYD_UINT16, linear_light=1,
                                                    tile_rows=(0, rows), d_out=d_out, d_out_cap=cap))
    eng.device_free(d_in); eng.device_free(d_out)
    return w * rows * 256, dt, n


def config5(eng):
    w, rows = 65536, 32
    n_in = w * rows * 256 * 3
    d_in = eng.device_alloc(n_in)
    eng.synth_fill(d_in, w, rows * 256, bits=8, x0=0, y0=0, full_width=65536, full_height=65536)
    cap = E.output_bound(w, rows * 256)
    d_out = eng.device_alloc(cap)
    dt, n = best_of(lambda: eng.encode_image_device(d_in, 65536, 65536, 3, tile_rows=(0, rows), d_out=d_out, d_out_cap=cap), reps=2)
    eng.device_free(d_in); eng.device_free(d_out)
    return w * rows * 256, dt, n


def config4(eng):
    w, h, count = 1920, 1080, 128
    ntx, nty = 8, 5
    img_bytes = w * h * 3
    d_in = eng.device_alloc(count * img_bytes)
    for k in range(count):
        eng.synth_fill(d_in + k * img_bytes, w, h, bits=8, seed=k)
    cap = count * E.output_bound(w, h)
    d_out = eng.device_alloc(cap)
    per_launch = max(1, eng.max_batch // (ntx * nty))
    launches = []
    for first in range(0, count, per_launch):
        tiles = []
        for k in range(first, min(count, first + per_launch)):
            base = d_in + k * img_bytes
            for ty in range(nty):
                for tx in range(ntx):
                    t = HydbTile()
                    p = base + (ty * 256 * w + tx * 256) * 3
                    t.plane = (C.c_void_p * 3)(p, p + 1, p + 2)
                    t.row_stride, t.pixel_stride = w * 3, 3
                    t.x0, t.y0 = tx * 256, ty * 256
                    t.width, t.height = min(256, w - tx * 256), min(256, h - ty * 256)
                    t.image_width, t.image_height = w, h
                    t.is_last = int(tx == ntx - 1 and ty == nty - 1)
                    t.sample_fmt, t.linear_light = HYD_UINT8, 0
                    t.with_image_header = int(tx == 0 and ty == 0)
                    tiles.append(t)
        launches.append((HydbTile * len(tiles))(*tiles))

    def run():
        pos = 0
        for arr in launches:
            eng._check(eng.lib.hydb_engine_encode_tiles(eng._h, arr, len(arr), d_out, cap, pos))
            n = C.c_uint64(0)
            eng._check(eng.lib.hydb_engine_finish(eng._h, C.byref(n)))
            pos += int(n.value)
        return pos
    dt, n = best_of(run)
    eng.device_free(d_in); eng.device_free(d_out)
    return w * h * count, dt, n


which = sys.argv[1:] or ["3", "4", "5"]
batches = [int(b) for b in os.environ.get("CONFIG_BENCH_BATCHES", "256,1024,4096").split(",")]
print(f"chain kernel: {os.environ.get('HYDRIUM_B200_CHAIN', 'auto')}", flush=True)
for batch in batches:
    with E.Engine(device=0, max_batch_tiles=batch) as eng:
        for name, fn in (("3", config3), ("4", config4), ("5", config5)):
            if name not in which:
                continue
            px, dt, n = fn(eng)
            print(f"config {name} share, batch {batch:5d} tiles: {px / 1e6:8.1f} Mpx in {dt * 1e3:8.2f} ms = {px / dt / 1e6:8.0f} Mpx/s"
                  f"  ({n / px:.3f} B/px out)", flush=True)
