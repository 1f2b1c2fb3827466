"""One small pass over every kernel family, meant to run under compute-sanitizer
(tests/test_gpu_parity.py::test_compute_sanitizer): whole-image device path with both rANS chain kernels,
a multi-group frame, one-frame mode over two LF groups, the asynchronous nine-symbol API with a tiny
output area (re-gather), a recurring chunk (CUDA-graph capture and replay), and float samples.  Exits non-zero when a result differs from the oracle /
reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hydrium_b200.encoder import encode_cli_loop  # noqa: E402
from hydrium_b200.engine import Engine  # noqa: E402
from hydrium_b200.lib import load_library  # noqa: E402
from hydrium_b200.synth import synth_image  # noqa: E402
from oracle.pyoracle import Oracle, have_ref, ref_library  # noqa: E402

lib = load_library()
orc = Oracle()
img = synth_image(300, 260, 8, seed=1)
want = orc.encode_image(img)
bad = 0
with Engine(device=0, max_batch_tiles=8) as eng:
    for mode in (Engine.CHAIN_TABLE, Engine.CHAIN_COMPACT):
        eng.set_chain_kernel(mode)
        bad += eng.encode_image(img) != want
        bad += eng.encode_image_host(img) != want
img16 = synth_image(300, 140, 16, seed=3)
bad += encode_cli_loop(lib, img16, linear_light=1) != orc.encode_image(img16, linear_light=1)
os.environ["HYDRIUM_B200_OUTCAP_KB"] = "64"
os.environ["HYDRIUM_B200_BATCH"] = "3"
bad += encode_cli_loop(lib, img) != want
del os.environ["HYDRIUM_B200_OUTCAP_KB"], os.environ["HYDRIUM_B200_BATCH"]
# the same chunk geometry three times: the second submission captures the job as a CUDA graph, the third replays it
# (engine.cu, hydb_engine_submit_frames); the staging copy runs on the helper-thread pool
wide = synth_image(2304, 256, 8, seed=6)
want_wide = orc.encode_image(wide)
for _ in range(3):
    bad += encode_cli_loop(lib, wide) != want_wide
f32 = (synth_image(270, 130, 16, seed=5).astype(np.float32) / np.float32(65535))
bad += encode_cli_loop(lib, f32) != orc.encode_image(f32)
if have_ref():
    ref = ref_library("O3")
    multi = synth_image(600, 300, 8, seed=2)
    bad += encode_cli_loop(lib, multi, shift_x=1, shift_y=1) != encode_cli_loop(ref, multi, shift_x=1, shift_y=1)
    two = synth_image(2100, 264, 8, seed=4)
    bad += encode_cli_loop(lib, two, shift_x=-1, shift_y=-1) != encode_cli_loop(ref, two, shift_x=-1, shift_y=-1)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
