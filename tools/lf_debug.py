"""Debug: compare section L (LF stream bits) of the GPU against the sequential host-compiled coder."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydrium_b200 import engine as E
from hydrium_b200.synth import synth_image
from hydrium_b200.lib import HydbTile
hh = C.CDLL("tests/host_harness/_build/libhost_harness.so")
eng = E.Engine(max_batch_tiles=8)
rng = np.random.default_rng(3)
imgs = [("tiny", synth_image(1, 1, 8)), ("thin", synth_image(257, 3, 8, seed=4)), ("flat", np.full((256, 256, 3), 128, np.uint8)),
        ("black", np.zeros((264, 40, 3), np.uint8)), ("s", synth_image(300, 260, 8))]
for name, img in imgs:
    h, w, ch = img.shape
    d_img = eng.upload(img); d_out = eng.device_alloc(1 << 20)
    for ty in range((h + 255) // 256):
        for tx in range((w + 255) // 256):
            t = HydbTile(); p = d_img + (ty * 256 * w * ch + tx * 256 * ch)
            t.plane = (C.c_void_p * 3)(p, p + 1, p + 2); t.row_stride, t.pixel_stride = w * ch, ch
            t.x0, t.y0 = tx * 256, ty * 256; t.width, t.height = min(256, w - tx * 256), min(256, h - ty * 256)
            t.image_width, t.image_height = w, h; t.is_last = 0; t.sample_fmt = 0; t.linear_light = 0
            try:
                eng.encode_tiles([t], d_out, 1 << 20)
            except Exception as e:
                print(name, tx, ty, "ERR", e)
            lfq = eng.read_tap(E.TAP_LFQ, 0, np.int32)
            bl = int(eng.read_tap(E.TAP_LFBITLEN, 0, np.uint32)[0])
            bits = eng.read_tap(E.TAP_LFBITS, 0, np.uint32)
            vbw, vbh = (t.width + 7) // 8, (t.height + 7) // 8
            out = np.zeros(8192, np.uint32); ebl = C.c_uint32()
            hh.hh_lf_stream(lfq.ctypes.data_as(C.c_void_p), vbw, vbh, out.ctypes.data_as(C.c_void_p), out.size, C.byref(ebl))
            a = np.unpackbits(bits.view(np.uint8), bitorder="little")[:bl]
            b = np.unpackbits(out.view(np.uint8), bitorder="little")[:ebl.value]
            m = min(len(a), len(b)); d = np.nonzero(a[:m] != b[:m])[0]
            print(name, (tx, ty), "gpu bits", bl, "expected", ebl.value, "first diff", d[0] if d.size else None)
            if d.size or bl != ebl.value:
                print("  gpu ", "".join(map(str, a[:120])))
                print("  want", "".join(map(str, b[:120])))
                print("  lfq", lfq.reshape(3, 32, 32)[:, :vbh, :vbw].reshape(3, -1)[:, :12])
