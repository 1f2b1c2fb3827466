"""One 4096x4096 encode per mode (tile_size_shift 3, one-frame) through the nine-symbol API: for an ncu
launch list (which kernels make up a multi-group frame's time) or, with HYDRIUM_B200_APITRACE=1, the
host-side phases of every hyd_send_tile."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydrium_b200.encoder import encode_cli_loop
from hydrium_b200.lib import load_library
from hydrium_b200.synth import synth_image
img = synth_image(4096, 4096, 8)
lib = load_library()
for rep in range(2):
    for sx in (3, -1):
        t0 = time.perf_counter()
        out = encode_cli_loop(lib, img, shift_x=sx, shift_y=sx)
        print(sx, len(out), f"{1e3 * (time.perf_counter() - t0):.1f} ms", file=sys.stderr, flush=True)
