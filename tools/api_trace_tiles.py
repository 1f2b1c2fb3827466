"""Phases of the nine-symbol API in tile mode with batching (HYDRIUM_B200_APITRACE=1 HYDRIUM_B200_BATCH=256)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HYDRIUM_B200_BATCH", "256")
os.environ.setdefault("HYDRIUM_B200_APITRACE", "1")
from hydrium_b200.encoder import encode_cli_loop
from hydrium_b200.lib import load_library
from hydrium_b200.synth import synth_image
img = synth_image(4096, 4096, 8)
lib = load_library()
for rep in range(3):
    t0 = time.perf_counter()
    out = encode_cli_loop(lib, img)
    print(len(out), f"{1e3 * (time.perf_counter() - t0):.1f} ms", file=sys.stderr, flush=True)
