"""Where k_frame_lf's time goes (cycles of the master CTA's thread 0, from the engine's clock tap): one
2048x2048 frame of 64 groups through hydb_engine_encode_frames."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydrium_b200 import engine as E
from hydrium_b200.abi import HYD_UINT8


class HydbFrame(C.Structure):
    _fields_ = [("plane", C.c_void_p * 3), ("row_stride", C.c_int64), ("pixel_stride", C.c_int64),
                ("width", C.c_uint32), ("height", C.c_uint32), ("x0", C.c_uint32), ("y0", C.c_uint32),
                ("image_width", C.c_uint32), ("image_height", C.c_uint32), ("is_last", C.c_int32),
                ("sample_fmt", C.c_int32), ("linear_light", C.c_int32), ("with_image_header", C.c_int32),
                ("one_frame", C.c_int32), ("lf_part", C.c_int32), ("preset", C.c_uint32), ("preset_bits", C.c_uint32),
                ("alpha_floor", C.c_uint32), ("clusters_per_preset", C.c_uint32)]


W = H = 2048
with E.Engine(device=0, max_batch_tiles=65) as eng:
    d_in = eng.device_alloc(W * H * 3)
    d_out = eng.device_alloc(64 << 20)
    eng.synth_fill(d_in, W, H, bits=8, seed=0)
    eng.enable_taps(True)
    f = HydbFrame()
    f.plane = (C.c_void_p * 3)(d_in, d_in + 1, d_in + 2)
    f.row_stride, f.pixel_stride = W * 3, 3
    f.width, f.height, f.image_width, f.image_height = W, H, W, H
    f.is_last, f.sample_fmt, f.with_image_header, f.one_frame = 1, HYD_UINT8, 1, 1
    lib = eng.lib
    lib.hydb_engine_encode_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_uint64]
    for _ in range(2):
        eng._check(lib.hydb_engine_encode_frames(eng._h, C.byref(f), 1, d_out, 64 << 20, 0))
        n = C.c_uint64(0)
        eng._check(lib.hydb_engine_finish(eng._h, C.byref(n)))
    clk = eng.read_tap(E.TAP_CLK, 0, np.uint32)
    last = int(clk[3])
    us = lambda c: c / 1965.0
    print(f"frame 2048x2048: {n.value} bytes")
    print(f"k_frame_lf master thread: residuals + section head {us(int(clk[0])):.0f} us, symbols + histogram {us(int(clk[1])):.0f} us, "
          f"code lengths + stream header {us(int(clk[2])):.0f} us, symbol bits {us(last & 0xFFFFF):.0f} us, HF metadata image {us((last >> 20) << 8):.0f} us")
