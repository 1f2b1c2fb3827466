"""Stage-by-stage diagnosis of the CUDA pipeline against the oracle (run on a GPU box).

Usage: python tools/gpu_diag.py [--quick]
Prints, per test tile, which stage first diverges and where.  Development tool; the formal
parity tests are tests/test_gpu_parity.py.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from hydrium_b200 import engine as E  # noqa: E402
from hydrium_b200.synth import synth_image  # noqa: E402
from oracle.pyoracle import Oracle, Stages  # noqa: E402

SCAN_V = [0, 1, 0, 0, 1, 2, 3, 2, 1, 0, 0, 1, 2, 3, 4, 5, 4, 3, 2, 1, 0, 0, 1, 2, 3, 4, 5, 6, 7, 6, 5, 4,
          3, 2, 1, 0, 1, 2, 3, 4, 5, 6, 7, 7, 6, 5, 4, 3, 2, 3, 4, 5, 6, 7, 7, 6, 5, 4, 5, 6, 7, 7, 6, 7]
SCAN_H = [0, 0, 1, 2, 1, 0, 0, 1, 2, 3, 4, 3, 2, 1, 0, 0, 1, 2, 3, 4, 5, 6, 5, 4, 3, 2, 1, 0, 0, 1, 2, 3,
          4, 5, 6, 7, 7, 6, 5, 4, 3, 2, 1, 2, 3, 4, 5, 6, 7, 7, 6, 5, 4, 3, 4, 5, 6, 7, 7, 6, 5, 6, 7, 7]


def first_diff(a, b):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    if a.size != b.size:
        return f"size {a.size} vs {b.size}"
    d = np.nonzero(a != b)[0]
    return "equal" if d.size == 0 else f"{d.size} diffs, first at {d[0]}: {a[d[0]]} vs {b[d[0]]}"


def check_tile(eng: E.Engine, orc: Oracle, img: np.ndarray, tx: int, ty: int, linear: int, verbose=True) -> bool:
    h, w, ch = img.shape
    st = Stages()
    ref_frame = orc.encode_tile(img, tx, ty, linear_light=linear, stages=st)
    vbw, vbh = st.vbw, st.vbh
    sw, sh = vbw * 8, vbh * 8
    # encode just this tile through the engine
    from hydrium_b200.lib import HydbTile
    import ctypes as C
    d_img = eng.upload(img)
    cap = 1 << 20
    d_out = eng.device_alloc(cap)
    item = img.dtype.itemsize
    t = HydbTile()
    p = d_img + (ty * 256 * w * ch + tx * 256 * ch) * item
    t.plane = (C.c_void_p * 3)(p, p + item, p + 2 * item)
    t.row_stride, t.pixel_stride = w * ch, ch
    t.x0, t.y0 = tx * 256, ty * 256
    t.width, t.height = min(256, w - tx * 256), min(256, h - ty * 256)
    t.image_width, t.image_height = w, h
    t.is_last = int((tx + 1) * 256 >= w and (ty + 1) * 256 >= h)
    t.sample_fmt = 0 if img.dtype == np.uint8 else 1
    t.linear_light = linear
    ok = True
    try:
        try:
            n = eng.encode_tiles([t], d_out, cap)
            frame = eng.download(d_out, n)
        except Exception as e:  # keep going: taps still tell where it went wrong
            print("   engine error:", e)
            frame = b""
            ok = False
        res = {}
        xyb = eng.read_tap(E.TAP_XYB, 0, np.float32, 1 << 20).reshape(256, 256, 3)[:sh, :sw]
        dct = eng.read_tap(E.TAP_DCT, 0, np.float32, 1 << 20).reshape(256, 256, 3)[:sh, :sw]
        o_xyb = st.xyb[:sh * sw * 3].reshape(sh, sw, 3)
        o_dct = st.dct[:sh * sw * 3].reshape(sh, sw, 3)
        res["xyb"] = first_diff(xyb.view(np.uint32), o_xyb.view(np.uint32))
        res["dct"] = "equal" if np.array_equal(dct, o_dct) else first_diff(dct.view(np.uint32), o_dct.view(np.uint32))
        # coefficients: device scan order [blk(32 stride)][c][64] vs oracle raster ints
        coef = eng.read_tap(E.TAP_COEF, 0, np.int16, 1 << 20).reshape(32, 32, 3, 64)[:vbh, :vbw]
        oq = st.quant[:sh * sw * 3].reshape(vbh, 8, vbw, 8, 3)
        exp = np.zeros((vbh, vbw, 3, 64), np.int64)
        for j in range(1, 64):
            exp[:, :, :, j] = oq[:, SCAN_H[j], :, SCAN_V[j], :]
        res["coef"] = first_diff(coef.astype(np.int64), exp)
        lfq = eng.read_tap(E.TAP_LFQ, 0, np.int32).reshape(3, 32, 32)[:, :vbh, :vbw]
        res["lfq"] = first_diff(lfq, oq[:, 0, :, 0, :].transpose(2, 0, 1))
        nzi = eng.read_tap(E.TAP_NZINFO, 0, np.uint16).reshape(32, 32, 3)[:vbh, :vbw]
        res["nz"] = first_diff(nzi & 0xFF, st.nonzeroes[:vbh * vbw * 3].reshape(vbh, vbw, 3))
        nsyms = int(eng.read_tap(E.TAP_NSYMS, 0, np.uint32)[0])
        syms = eng.read_tap(E.TAP_SYMS, 0, np.uint32, 1 << 20)
        os_ = st.hf_syms[:st.n_syms]
        opacked = (os_[:, 0] | (os_[:, 1] << 8) | (os_[:, 2] << 12) | (os_[:, 3] << 16)).astype(np.uint32)
        res["syms"] = f"n={nsyms} vs {st.n_syms}; " + first_diff(syms, opacked)
        freqs = eng.read_tap(E.TAP_FREQS, 0, np.uint32).reshape(9, 64)
        res["freqs"] = first_diff(freqs, st.freqs[:, :64])
        sect = eng.read_tap(E.TAP_SECT, 0, np.uint32)
        lfbitlen = int(eng.read_tap(E.TAP_LFBITLEN, 0, np.uint32)[0])
        res["sect"] = f"prefixABL={sect[0]} D={sect[1]} E={sect[2]} total={sect[3]} lfbits={lfbitlen} | oracle lf={st.lf_bitlen} pre={st.pre_bitlen} E={st.ans_bitlen}"
        payload = eng.read_tap(E.TAP_PAYLOAD, 0, np.uint8, 1 << 20).tobytes() if frame else b""
        # oracle payload = pre bits + ans bits
        pre = np.unpackbits(np.frombuffer(st.bits("pre"), np.uint8), bitorder="little")[:st.pre_bitlen]
        ans = np.unpackbits(np.frombuffer(st.bits("ans"), np.uint8), bitorder="little")[:st.ans_bitlen]
        full = np.concatenate([pre, ans])
        full = np.concatenate([full, np.zeros((-len(full)) % 8, np.uint8)])
        opay = np.packbits(full, bitorder="little").tobytes()
        if payload:
            pb = np.unpackbits(np.frombuffer(payload, np.uint8), bitorder="little")
            ob = np.unpackbits(np.frombuffer(opay, np.uint8), bitorder="little")
            m = min(len(pb), len(ob))
            d = np.nonzero(pb[:m] != ob[:m])[0]
            res["payload"] = f"{len(payload)} vs {len(opay)} bytes; " + ("equal" if d.size == 0 and len(pb) == len(ob) else f"first bit diff at {d[0] if d.size else m}")
        res["frame"] = "equal" if frame == ref_frame else f"DIFF ({len(frame)} vs {len(ref_frame)} bytes)"
        bad = [k for k, v in res.items() if k not in ("sect",) and "equal" not in v]
        ok = ok and not bad
        if verbose or bad:
            print(f" tile ({tx},{ty}) of {w}x{h} {img.dtype} vb={vbw}x{vbh}: {'OK' if not bad else 'FAIL ' + ','.join(bad)}")
            if bad or verbose:
                for k, v in res.items():
                    print(f"     {k:8s} {v}")
    finally:
        eng.device_free(d_img)
        eng.device_free(d_out)
    return ok


def main():
    quick = "--quick" in sys.argv
    orc = Oracle()
    t0 = time.time()
    eng = E.Engine(max_batch_tiles=64)
    print(f"engine up in {time.time() - t0:.2f}s")
    eng.enable_taps(True)
    rng = np.random.default_rng(3)
    cases = [
        (synth_image(256, 256, 8), [(0, 0)], 0),
        (synth_image(700, 600, 8), [(0, 0), (2, 2), (1, 2), (2, 0)], 0),
        (synth_image(512, 512, 16), [(0, 0), (1, 1)], 1),
        (np.full((256, 256, 3), 128, np.uint8), [(0, 0)], 0),
        ((rng.integers(0, 2, (300, 260, 3)) * 255).astype(np.uint8), [(0, 0), (1, 1), (1, 0)], 0),
        (synth_image(1000, 300, 8, smooth=True), [(0, 0), (3, 1)], 0),
        (synth_image(300, 300, 16, seed=5), [(1, 1)], 0),
    ]
    all_ok = True
    for img, tiles, lin in cases:
        for (tx, ty) in tiles:
            all_ok &= check_tile(eng, orc, img, tx, ty, lin, verbose=not quick)
    eng.enable_taps(False)
    # whole images through the batch path
    for (w, h, bits, lin) in [(700, 600, 8, 0), (512, 512, 16, 1), (1920, 1080, 8, 0), (1024, 1024, 8, 0)]:
        img = synth_image(w, h, bits)
        t0 = time.time()
        mine = eng.encode_image(img, linear_light=lin)
        t1 = time.time()
        ref = orc.encode_image(img, linear_light=lin)
        same = mine == ref
        all_ok &= same
        print(f"image {w}x{h} u{bits}: {'OK' if same else 'FAIL'} {len(mine)} vs {len(ref)} bytes, gpu call {1e3 * (t1 - t0):.1f} ms")
    print("ALL OK" if all_ok else "SOME FAILED")
    return 0 if all_ok else 1


if __name__ == "__main__":
    sys.exit(main())
