"""Timing of multi-group frames through the nine-symbol API (host buffers): tile_size_shift 3 and
one-frame mode on a 4096x4096 image, next to the reference library on one host thread."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydrium_b200.encoder import encode_cli_loop
from hydrium_b200.lib import load_library
from hydrium_b200.synth import synth_image
from oracle.pyoracle import ref_library, have_ref

W = H = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
img = synth_image(W, H, 8)
lib = load_library()
ref = ref_library("Os") if have_ref() else None
for name, sx in (("tile mode 256^2 (shift 0), sync", 0), ("tile mode 2048^2 (shift 3)", 3), ("one-frame mode", -1)):
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        out = encode_cli_loop(lib, img, shift_x=sx, shift_y=sx)
        best = min(best, time.perf_counter() - t0)
    line = f"{name:34s} {len(out):9d} bytes  ours {best * 1e3:8.1f} ms = {W * H / best / 1e6:7.1f} Mpx/s"
    if ref is not None:
        t0 = time.perf_counter()
        want = encode_cli_loop(ref, img, shift_x=sx, shift_y=sx)
        dt = time.perf_counter() - t0
        line += f"   reference {dt * 1e3:8.1f} ms = {W * H / dt / 1e6:6.1f} Mpx/s   identical {out == want}"
    print(line, flush=True)
