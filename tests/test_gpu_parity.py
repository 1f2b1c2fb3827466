"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same
inputs (bit-exact: integer / byte output; the float stages are compared bit-for-bit as well,
except that a zero may differ in sign in the DCT tap, which no later stage can observe).

Everything here needs a B200 (`-m gpu`).  /root/reference is not used at run time."""
import ctypes as C

import numpy as np
import pytest

from hydrium_b200 import engine as E
from hydrium_b200.abi import HYD_API_ERROR, HYD_FLOAT32, HYD_NEED_MORE_OUTPUT, HYD_UINT8, HYD_UINT16
from hydrium_b200.encoder import HYDEncoder, HydriumError, encode_cli_loop, tile_grid
from hydrium_b200.lib import HydbTile
from hydrium_b200.synth import synth_image
from oracle.pyoracle import Stages
from util import GOLDEN, image_set, kat_image, kat_table, sha256

pytestmark = pytest.mark.gpu

SCAN_V = [0, 1, 0, 0, 1, 2, 3, 2, 1, 0, 0, 1, 2, 3, 4, 5, 4, 3, 2, 1, 0, 0, 1, 2, 3, 4, 5, 6, 7, 6, 5, 4,
          3, 2, 1, 0, 1, 2, 3, 4, 5, 6, 7, 7, 6, 5, 4, 3, 2, 3, 4, 5, 6, 7, 7, 6, 5, 4, 5, 6, 7, 7, 6, 7]
SCAN_H = [0, 0, 1, 2, 1, 0, 0, 1, 2, 3, 4, 3, 2, 1, 0, 0, 1, 2, 3, 4, 5, 6, 5, 4, 3, 2, 1, 0, 0, 1, 2, 3,
          4, 5, 6, 7, 7, 6, 5, 4, 3, 2, 1, 2, 3, 4, 5, 6, 7, 7, 6, 5, 4, 3, 4, 5, 6, 7, 7, 6, 5, 6, 7, 7]


def _tile_desc(d_img, img, tx, ty, linear, image_size=None, origin=None):
    h, w, ch = img.shape
    item = img.dtype.itemsize
    t = HydbTile()
    if origin is None:
        p = d_img + (ty * 256 * w * ch + tx * 256 * ch) * item
        t.width, t.height = min(256, w - tx * 256), min(256, h - ty * 256)
    else:   # `img` is the tile window itself
        p = d_img
        t.width, t.height = w, h
    t.plane = (C.c_void_p * 3)(p, p + item, p + 2 * item)
    t.row_stride, t.pixel_stride = w * ch, ch
    t.x0, t.y0 = tx * 256, ty * 256
    iw, ih = (w, h) if image_size is None else image_size
    t.image_width, t.image_height = iw, ih
    t.is_last = int((tx + 1) * 256 >= iw and (ty + 1) * 256 >= ih)
    t.sample_fmt = E._fmt_of(img)
    t.linear_light = linear
    return t


def test_every_stage_matches_oracle(engine, oracle):
    """T0..T7 of SURVEY.md section 4 for a spread of tiles, via the engine's stage taps."""
    engine.enable_taps(True)
    rng = np.random.default_rng(3)
    try:
        for name, img, lin in image_set(rng):
            h, w, ch = img.shape
            d_img = engine.upload(img)
            d_out = engine.device_alloc(1 << 20)
            try:
                for ty in range((h + 255) // 256):
                    for tx in range((w + 255) // 256):
                        st = Stages()
                        want = oracle.encode_tile(img, tx, ty, linear_light=lin, stages=st)
                        n = engine.encode_tiles([_tile_desc(d_img, img, tx, ty, lin)], d_out, 1 << 20)
                        got = engine.download(d_out, n)
                        vbw, vbh = st.vbw, st.vbh
                        sw, sh = vbw * 8, vbh * 8
                        tag = (name, tx, ty)
                        xyb = engine.read_tap(E.TAP_XYB, 0, np.float32).reshape(256, 256, 3)[:sh, :sw]
                        assert np.array_equal(xyb.view(np.uint32), st.xyb[:sh * sw * 3].reshape(sh, sw, 3).view(np.uint32)), (tag, "xyb")
                        dct = engine.read_tap(E.TAP_DCT, 0, np.float32).reshape(256, 256, 3)[:sh, :sw]
                        assert np.array_equal(dct, st.dct[:sh * sw * 3].reshape(sh, sw, 3)), (tag, "dct")
                        coef = engine.read_tap(E.TAP_COEF, 0, np.int16).reshape(32, 32, 3, 64)[:vbh, :vbw]
                        oq = st.quant[:sh * sw * 3].reshape(vbh, 8, vbw, 8, 3)
                        for j in range(1, 64):
                            assert np.array_equal(coef[:, :, :, j], oq[:, SCAN_H[j], :, SCAN_V[j], :]), (tag, "coef", j)
                        lfq = engine.read_tap(E.TAP_LFQ, 0, np.int32).reshape(3, 32, 32)[:, :vbh, :vbw]
                        assert np.array_equal(lfq, oq[:, 0, :, 0, :].transpose(2, 0, 1)), (tag, "lf ints")
                        nz = engine.read_tap(E.TAP_NZINFO, 0, np.uint16).reshape(32, 32, 3)[:vbh, :vbw] & 0xFF
                        assert np.array_equal(nz, st.nonzeroes[:vbh * vbw * 3].reshape(vbh, vbw, 3)), (tag, "nz")
                        syms = engine.read_tap(E.TAP_SYMS, 0, np.uint32)
                        s = st.hf_syms[:st.n_syms]
                        assert np.array_equal(syms, (s[:, 0] | (s[:, 1] << 8) | (s[:, 2] << 12) | (s[:, 3] << 16)).astype(np.uint32)), (tag, "symbols")
                        assert np.array_equal(engine.read_tap(E.TAP_FREQS, 0, np.uint32).reshape(9, 64), st.freqs[:, :64]), (tag, "freqs")
                        sect = engine.read_tap(E.TAP_SECT, 0, np.uint32)
                        assert int(sect[0] + sect[1]) == st.pre_bitlen and int(sect[2]) == st.ans_bitlen, (tag, "section lengths")
                        assert got == want, (tag, "frame")
            finally:
                engine.device_free(d_img)
                engine.device_free(d_out)
    finally:
        engine.enable_taps(False)


@pytest.mark.parametrize("name", [k for k, v in kat_table().items() if "width" in v and "ref_only" not in k and v["shift"] == 0])
def test_golden_known_answers(engine, name):
    """Committed reference outputs (tests/golden/kat.json) -- no oracle involved."""
    e = kat_table()[name]
    out = engine.encode_image(kat_image(e), linear_light=e["linear_light"])
    assert len(out) == e["length"] and sha256(out) == e["sha256"]


def test_golden_files(engine):
    t = kat_table()
    for fname, key in [("synth_40x24.jxl", "T_40x24"), ("synth_300x260.jxl", "Q_300x260")]:
        assert engine.encode_image(kat_image(t[key])) == open(f"{GOLDEN}/{fname}", "rb").read()


def test_whole_images_device_and_host_paths(engine, oracle):
    rng = np.random.default_rng(5)
    for name, img, lin in image_set(rng):
        want = oracle.encode_image(img, linear_light=lin)
        assert engine.encode_image(img, linear_light=lin) == want, (name, "device path")
        assert engine.encode_image_host(img, linear_light=lin) == want, (name, "host path")


def test_compact_chain_kernel(product_lib, oracle):
    """k_ans_chain_compact (sorted alias pieces + warp-wide search, sixteen chains per SM) against the
    oracle and against the table kernel: frequencies, section lengths and bytes, over the image set
    (edge tiles, flat tiles, 16-bit linear), and on a launch large enough to fill every SM several times."""
    rng = np.random.default_rng(5)
    with E.Engine(device=0, max_batch_tiles=64) as eng:
        eng.set_chain_kernel(E.Engine.CHAIN_COMPACT)
        eng.enable_taps(True)
        for name, img, lin in image_set(rng):
            if img.dtype == np.float32:
                continue   # float tiles always take the table kernel (tokens may pass 32)
            want = oracle.encode_image(img, linear_light=lin)
            assert eng.encode_image(img, linear_light=lin) == want, (name, "compact kernel")
            st = Stages()
            oracle.encode_tile(img, 0, 0, linear_light=lin, stages=st)
            d_img, d_out = eng.upload(img), eng.device_alloc(1 << 20)
            try:
                eng.encode_tiles([_tile_desc(d_img, img, 0, 0, lin)], d_out, 1 << 20)
                assert np.array_equal(eng.read_tap(E.TAP_FREQS, 0, np.uint32).reshape(9, 64), st.freqs[:, :64]), (name, "freqs")
                sect = eng.read_tap(E.TAP_SECT, 0, np.uint32)
                assert int(sect[0] + sect[1]) == st.pre_bitlen and int(sect[2]) == st.ans_bitlen, (name, "section lengths")
            finally:
                eng.device_free(d_img)
                eng.device_free(d_out)
    # 1280 tiles in one launch (four times what the table kernel keeps resident): both kernels and
    # the automatic choice give the same bytes, and those are the oracle's
    w, h = 8192, 10 * 256
    outs = []
    img = None
    for mode in (E.Engine.CHAIN_TABLE, E.Engine.CHAIN_COMPACT, E.Engine.CHAIN_AUTO):
        with E.Engine(device=0, max_batch_tiles=(w // 256) * (h // 256)) as eng:
            eng.set_chain_kernel(mode)
            d_in, cap = eng.device_alloc(w * h * 3), E.output_bound(w, h)
            d_out = eng.device_alloc(cap)
            eng.synth_fill(d_in, w, h, bits=8, seed=7)
            outs.append(eng.download(d_out, eng.encode_image_device(d_in, w, h, 3, d_out=d_out, d_out_cap=cap)))
            if img is None:
                img = np.frombuffer(eng.download(d_in, w * h * 3), np.uint8).reshape(h, w, 3)
            eng.device_free(d_in)
            eng.device_free(d_out)
    assert outs[0] == outs[1] == outs[2]
    assert outs[1] == oracle.encode_image(img)
    # 512 tiles: more chains than the table kernel keeps resident (296) but below the compact kernel's
    # threshold -- two rounds of the table kernel (band-pipelined whole image, and one plain batch through
    # hydb_engine_encode_tiles)
    w, h = 4096, 32 * 256
    with E.Engine(device=0, max_batch_tiles=512) as eng:
        d_in, cap = eng.device_alloc(w * h * 3), E.output_bound(w, h)
        d_out = eng.device_alloc(cap)
        eng.synth_fill(d_in, w, h, bits=8, seed=9)
        got = eng.download(d_out, eng.encode_image_device(d_in, w, h, 3, d_out=d_out, d_out_cap=cap))
        img = np.frombuffer(eng.download(d_in, w * h * 3), np.uint8).reshape(h, w, 3)
        want = oracle.encode_image(img)
        assert got == want, "512 tiles, band pipeline"
        tiles = [_tile_desc(d_in, img, tx, ty, 0) for ty in range(32) for tx in range(16)]
        tiles[0].with_image_header = 1
        n = eng.encode_tiles(tiles, d_out, cap)
        assert eng.download(d_out, n) == want, "512 tiles, one batch"
        eng.device_free(d_in)
        eng.device_free(d_out)


def test_batches_larger_than_the_workspace(product_lib, oracle):
    """An image with more tiles than max_batch_tiles goes through several launches."""
    img = synth_image(1300, 1100, 8, seed=2)   # 6 x 5 = 30 tiles
    with E.Engine(device=0, max_batch_tiles=7) as eng:
        assert eng.encode_image(img) == oracle.encode_image(img)


def test_nine_symbol_api_cli_loop(product_lib, oracle):
    """The reference CLI's call sequence (hydrium.c:402-480) against our library.  Default
    (asynchronous) mode: the concatenation is the reference's.  hydb_encoder_set_batch(1): every
    tile's bytes surface in the flush loop that follows it, like the reference."""
    for img, lin in [(synth_image(700, 600, 8), 0), (synth_image(300, 520, 16, seed=3), 1)]:
        assert encode_cli_loop(product_lib, img, linear_light=lin) == oracle.encode_image(img, linear_light=lin)
        per = []
        out = encode_cli_loop(product_lib, img, linear_light=lin, per_tile=per, batch=1)
        assert out == oracle.encode_image(img, linear_light=lin)
        h, w, _ = img.shape
        hdr = oracle.image_header(w, h)
        k = 0
        for ty in range((h + 255) // 256):
            for tx in range((w + 255) // 256):
                want = oracle.encode_tile(img, tx, ty, linear_light=lin)
                assert per[k] == (hdr + want if k == 0 else want), (tx, ty)
                k += 1


def test_nine_symbol_api_small_output_buffer(product_lib, oracle):
    """HYD_NEED_MORE_OUTPUT back-pressure (libhydrium.c:147-166): 64-byte output buffer."""
    img = synth_image(300, 260, 8, seed=1)
    assert encode_cli_loop(product_lib, img, out_buf_size=64) == oracle.encode_image(img)


def test_nine_symbol_api_one_frame_single_group(product_lib, oracle):
    """BASELINE config 1: 256x256 through the CLI-default one-frame mode."""
    e = kat_table()["A_256_oneframe"]
    out = encode_cli_loop(product_lib, kat_image(e), shift_x=-1, shift_y=-1)
    assert sha256(out) == e["sha256"]
    img = synth_image(200, 131, 8, seed=8)
    assert encode_cli_loop(product_lib, img, shift_x=-1, shift_y=-1) == oracle.encode_image(img)


def test_nine_symbol_api_batched_mode(product_lib, oracle):
    """hydb_encoder_set_batch: output is deferred but the concatenation is identical."""
    img = synth_image(1100, 800, 8, seed=6)   # 5 x 4 tiles
    want = oracle.encode_image(img)
    h, w, ch = img.shape
    for batch in (3, 20, 64):
        enc = HYDEncoder(product_lib)
        assert product_lib.hydb_encoder_set_batch(enc._enc, batch) == 0
        enc.check(enc.set_metadata(w, h, 0, 0, 0))
        obuf = np.empty(1 << 16, np.uint8)
        enc.check(enc.provide_output_buffer(obuf))
        out = bytearray()
        _, _, ntx, nty = tile_grid(w, h, 0, 0)
        for ty in range(nty):
            for tx in range(ntx):
                p = img.ctypes.data + (ty * 256 * w + tx * 256) * ch
                enc.check(enc.send_tile((p, p + 1, p + 2), tx, ty, w * ch, ch, -1, HYD_UINT8))
                while True:
                    ret = enc.flush()
                    _, n = enc.release_output_buffer()
                    out += obuf[:n].tobytes()
                    enc.check(enc.provide_output_buffer(obuf))
                    if ret != -2:
                        break
                enc.check(ret)
        enc.destroy()
        assert bytes(out) == want, batch


def _send_all(lib, img, calls, *, shift=0, lin=0, flush_every=1, env=None, remeta=None, pixel_stride=None, planes=None):
    """Drive a libhydrium build with an explicit list of (tile_x, tile_y, is_last) calls; the output buffer is
    drained after every `flush_every`-th tile (0 = only at the end) and once more at the end."""
    import os
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        h, w, ch = img.shape
        item = img.dtype.itemsize
        enc = HYDEncoder(lib)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    out = bytearray()
    obuf = np.empty(1 << 20, np.uint8)   # like the CLI; the reference writes headers straight into it

    def drain():
        while True:
            ret = enc.flush()
            _, n = enc.release_output_buffer()
            out.extend(obuf[:n].tobytes())
            enc.check(enc.provide_output_buffer(obuf))
            if ret != HYD_NEED_MORE_OUTPUT:
                enc.check(ret)
                return
    enc.check(enc.set_metadata(w, h, lin, shift, shift))
    enc.check(enc.provide_output_buffer(obuf))
    tw = 256 << shift
    for i, (tx, ty, last) in enumerate(calls):
        if remeta is not None and i == remeta[0]:
            shift = remeta[1]
            tw = 256 << shift
            enc.check(enc.set_metadata(w, h, lin, shift, shift))
        p = img.ctypes.data + (ty * tw * w * ch + tx * tw * ch) * item
        pl = (p, p + item, p + 2 * item) if planes is None else tuple(p + o * item for o in planes)
        enc.check(enc.send_tile(pl, tx, ty, w * ch, ch if pixel_stride is None else pixel_stride, last, E._fmt_of(img)))
        if flush_every and (i + 1) % flush_every == 0:
            drain()
    drain()
    drain()   # a second poll in a row: whatever is left (no tile was marked last) must come out now
    enc.destroy()
    return bytes(out)


def test_async_pipeline_orders_gaps_and_backpressure(product_lib, reflib):
    """The default (asynchronous) nine-symbol path against the reference library driven with the very same
    calls: raster, reversed and shuffled send orders, the lower-right tile first with is_last = -1, tile
    subsets that never mark a last tile (libhydrium.h:235-240), a caller that never drains until the end
    (finished chunks spill to the heap), tiny chunks, and an output area too small for a chunk (re-gather)."""
    img = synth_image(1100, 800, 8, seed=6)   # 5 x 4 tiles
    tiles = [(x, y) for y in range(4) for x in range(5)]
    rng = np.random.default_rng(11)
    shuffled = [tiles[i] for i in rng.permutation(len(tiles))]
    cases = {
        "raster": [(x, y, -1) for x, y in tiles],
        "reversed": [(x, y, -1) for x, y in reversed(tiles)],           # the lower-right tile comes first
        "shuffled": [(x, y, -1) for x, y in shuffled],
        "subset, no last tile": [(x, y, 0) for x, y in shuffled[:7]],
        "subset, explicit last in the middle": [(x, y, int(i == 3)) for i, (x, y) in enumerate(shuffled[:9])],
    }
    for name, calls in cases.items():
        want = _send_all(reflib, img, calls)
        assert _send_all(product_lib, img, calls) == want, name
        assert _send_all(product_lib, img, calls, flush_every=0) == want, (name, "drained only at the end")
        assert _send_all(product_lib, img, calls, env={"HYDRIUM_B200_BATCH": "3", "HYDRIUM_B200_DEPTH": "2"}, flush_every=0) == want, (name, "chunks of 3, 2 in flight")
        assert _send_all(product_lib, img, calls, env={"HYDRIUM_B200_BATCH": "1"}) == want, (name, "synchronous")
    calls = cases["raster"]
    want = _send_all(reflib, img, calls)
    assert _send_all(product_lib, img, calls, env={"HYDRIUM_B200_OUTCAP_KB": "64"}) == want, "re-gather path"
    # larger tiles and 16-bit samples through the same machinery (one multi-group frame per chunk)
    img16 = synth_image(1100, 800, 16, seed=2)
    calls = [(x, y, -1) for y in range(2) for x in range(3)]
    # (a full-size tile first: the reference sizes its per-group arrays by the first tile it sees and
    # corrupts its heap when a later one has more groups)
    order = [calls[0]] + calls[:0:-1]
    want = _send_all(reflib, img16, order, shift=1, lin=1)
    assert _send_all(product_lib, img16, order, shift=1, lin=1) == want
    assert _send_all(product_lib, img16, order, shift=1, lin=1, env={"HYDRIUM_B200_OUTCAP_KB": "64"}, flush_every=0) == want


def test_metadata_changed_on_a_live_encoder(product_lib, reflib, oracle):
    """hyd_set_metadata again on a used encoder with a larger tile geometry: the staging ring is rebuilt,
    nothing overflows, and the stream continues without a second image header.  (The reference itself
    corrupts its heap on this sequence -- it sizes its group arrays once -- so the expectation is put
    together from two of its encoders.)"""
    img = synth_image(1100, 800, 8, seed=9)
    h, w, _ = img.shape
    first = [(0, 0, 0), (1, 0, 0)]            # two 256-tiles
    second = [(0, 0, 0), (0, 0, 1)]           # then shift 2: 1024-tiles
    hdr = oracle.image_header(w, h)
    part_b = _send_all(reflib, img, second, shift=2)
    assert part_b.startswith(hdr)
    want = _send_all(reflib, img, first) + part_b[len(hdr):]
    assert _send_all(product_lib, img, first + second, remeta=(2, 2)) == want
    assert _send_all(product_lib, img, first + second, remeta=(2, 2), env={"HYDRIUM_B200_BATCH": "1"}) == want


def test_interleaved_layout_at_the_end_of_the_callers_buffer(product_lib, oracle):
    """ARGB: the three planes start one sample into each pixel.  The staging copy must not read past the
    last addressed sample -- the image here ends exactly at an inaccessible page."""
    import ctypes
    import mmap
    w, h = 300, 270
    rgb = synth_image(w, h, 8, seed=12)
    argb = np.concatenate([np.full((h, w, 1), 9, np.uint8), rgb], axis=2)
    nbytes = argb.nbytes - 0   # A R G B per pixel; the last B is the last byte
    page = mmap.PAGESIZE
    total = (nbytes + page - 1) // page * page + page
    m = mmap.mmap(-1, total)
    base = ctypes.addressof(ctypes.c_char.from_buffer(m))
    libc = ctypes.CDLL(None, use_errno=True)
    assert libc.mprotect(ctypes.c_void_p(base + total - page), page, 0) == 0   # PROT_NONE guard page
    start = total - page - nbytes
    view = np.frombuffer(m, np.uint8, count=nbytes, offset=start).reshape(h, w, 4)
    view[:] = argb
    got = _send_all(product_lib, view, [(x, y, -1) for y in range(2) for x in range(2)], pixel_stride=4, planes=(1, 2, 3))
    assert got == oracle.encode_image(rgb)
    del view
    libc.mprotect(ctypes.c_void_p(base + total - page), page, 3)
    m.close()


def test_float_samples_beyond_the_16_bit_coefficient_range(product_lib, reflib):
    """HYD_FLOAT32 samples scaled far outside [0, 1]: quantised coefficients leave int16, which the B200
    encoder's records are sized for (the reference keeps int32, encoder.c:808).  Either the bytes are the
    reference's or the call fails loudly -- never different bytes."""
    base = (synth_image(300, 260, 16, seed=4).astype(np.float32) / np.float32(65535))
    for scale in (8.0, 300.0, 1e6):
        img = (base * np.float32(scale)).astype(np.float32)
        try:
            want = encode_cli_loop(reflib, img)
        except HydriumError:
            want = None   # the reference gives up on its own (token alphabet beyond its tables)
        try:
            got = encode_cli_loop(product_lib, img)
        except HydriumError as e:
            assert e.code < -10 and ("16-bit range" in (e.message or "") or "alphabet" in (e.message or "")), (scale, e.message)
            continue
        assert want is not None and got == want, scale


def test_engine_jobs_api(product_lib, oracle):
    """hydb_engine_submit_frames / job_poll / job_regather / job_release (include/hydrium_b200.h): two jobs in
    flight on disjoint workspace slots, one gathering into device memory and one straight into page-locked
    host memory, pixels copied from page-locked staging by the job itself; an output area that is too small
    answers HYD_NEED_MORE_OUTPUT and the frames are gathered again."""
    from hydrium_b200.lib import load_library

    class HydbFrame(C.Structure):
        _fields_ = [("plane", C.c_void_p * 3), ("row_stride", C.c_int64), ("pixel_stride", C.c_int64),
                    ("width", C.c_uint32), ("height", C.c_uint32), ("x0", C.c_uint32), ("y0", C.c_uint32),
                    ("image_width", C.c_uint32), ("image_height", C.c_uint32), ("is_last", C.c_int32),
                    ("sample_fmt", C.c_int32), ("linear_light", C.c_int32), ("with_image_header", C.c_int32),
                    ("one_frame", C.c_int32), ("lf_part", C.c_int32), ("preset", C.c_uint32), ("preset_bits", C.c_uint32),
                    ("alpha_floor", C.c_uint32), ("clusters_per_preset", C.c_uint32)]

    lib = load_library()
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    lib.hydb_engine_submit_frames.restype = C.c_int
    lib.hydb_engine_submit_frames.argtypes = [vp, vp, u32, u32, vp, vp, C.c_size_t, vp, u64, C.POINTER(u32), C.POINTER(u32)]
    lib.hydb_engine_job_poll.restype = C.c_int
    lib.hydb_engine_job_poll.argtypes = [vp, u32, C.c_int, C.POINTER(u64)]
    lib.hydb_engine_job_regather.restype = C.c_int
    lib.hydb_engine_job_regather.argtypes = [vp, u32, vp, u64, C.POINTER(u64)]
    lib.hydb_engine_job_release.restype = C.c_int
    lib.hydb_engine_job_release.argtypes = [vp, u32]
    imgs = [synth_image(512, 256, 8, seed=21), synth_image(700, 200, 8, seed=22)]   # 2 and 3 tiles in a row
    with E.Engine(device=0, max_batch_tiles=8) as eng:
        jobs = []
        for k, img in enumerate(imgs):
            h, w, _ = img.shape
            n_in = img.nbytes
            h_stage = lib.hydb_host_alloc(n_in)
            C.memmove(h_stage, img.ctypes.data, n_in)
            d_stage = eng.device_alloc(n_in)
            ntx = (w + 255) // 256
            frames = (HydbFrame * ntx)()
            for tx in range(ntx):
                f = frames[tx]
                p = d_stage + tx * 256 * 3
                f.plane = (C.c_void_p * 3)(p, p + 1, p + 2)
                f.row_stride, f.pixel_stride = w * 3, 3
                f.width, f.height, f.x0, f.y0 = min(256, w - tx * 256), h, tx * 256, 0
                f.image_width, f.image_height = w, h
                f.is_last, f.sample_fmt, f.with_image_header = int(tx == ntx - 1), HYD_UINT8, int(tx == 0)
            cap = 1 << 20
            out_host = lib.hydb_host_alloc(cap) if k == 1 else None
            out = out_host if k == 1 else eng.device_alloc(cap)
            job, slots = u32(0), u32(0)
            eng._check(lib.hydb_engine_submit_frames(eng._h, frames, ntx, 4 * k, h_stage, d_stage, n_in, out, cap, C.byref(job), C.byref(slots)))
            assert slots.value == ntx
            jobs.append((job.value, out, out_host, h_stage, d_stage, img))
        for job, out, out_host, h_stage, d_stage, img in jobs:
            nbytes = u64(0)
            eng._check(lib.hydb_engine_job_poll(eng._h, job, 1, C.byref(nbytes)))
            got = C.string_at(out_host, nbytes.value) if out_host else eng.download(out, nbytes.value)
            assert got == oracle.encode_image(img)
            eng._check(lib.hydb_engine_job_release(eng._h, job))
        # too small an output area: HYD_NEED_MORE_OUTPUT, then gathered again into a big enough one
        job, out, out_host, h_stage, d_stage, img = jobs[0]
        frames = (HydbFrame * 2)()
        for tx in range(2):
            f = frames[tx]
            p = d_stage + tx * 256 * 3
            f.plane = (C.c_void_p * 3)(p, p + 1, p + 2)
            f.row_stride, f.pixel_stride = 512 * 3, 3
            f.width, f.height, f.x0, f.y0, f.image_width, f.image_height = 256, 256, tx * 256, 0, 512, 256
            f.is_last, f.sample_fmt, f.with_image_header = int(tx == 1), HYD_UINT8, int(tx == 0)
        j2 = u32(0)
        eng._check(lib.hydb_engine_submit_frames(eng._h, frames, 2, 0, None, None, 0, out, 4096, C.byref(j2), None))
        nbytes = u64(0)
        assert lib.hydb_engine_job_poll(eng._h, j2.value, 1, C.byref(nbytes)) == HYD_NEED_MORE_OUTPUT
        eng._check(lib.hydb_engine_job_regather(eng._h, j2.value, out, 1 << 20, C.byref(nbytes)))
        assert eng.download(out, nbytes.value) == oracle.encode_image(img)
        eng._check(lib.hydb_engine_job_release(eng._h, j2.value))
        for job, out, out_host, h_stage, d_stage, img in jobs:
            lib.hydb_host_free(h_stage)
            eng.device_free(d_stage)
            if out_host:
                lib.hydb_host_free(out_host)
            else:
                eng.device_free(out)


def test_compute_sanitizer(product_lib):
    """compute-sanitizer memcheck over one small pass through every kernel family (tools/sanitize_case.py:
    both rANS chain kernels, TMA staging, multi-group frames, one-frame mode over two LF groups, the
    asynchronous API with re-gather, float samples).  Invalid accesses or leaks of device errors fail it."""
    import os
    import shutil
    import subprocess
    import sys
    exe = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(exe):
        pytest.skip("compute-sanitizer not installed")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([exe, "--tool", "memcheck", "--error-exitcode", "86", "--launch-timeout", "120", sys.executable,
                        os.path.join(root, "tools", "sanitize_case.py")], capture_output=True, text=True, timeout=1500)
    tail = (p.stdout + p.stderr)[-3000:]
    assert p.returncode == 0, tail
    assert "ERROR SUMMARY: 0 errors" in p.stdout + p.stderr, tail


def test_sample_layouts(product_lib, oracle):
    """Planar, RGBA-interleaved, BGR-ordered, negative row stride and odd strides (libhydrium.h:202-220)."""
    rgb = synth_image(300, 270, 8, seed=12)
    want = oracle.encode_image(rgb)
    h, w, _ = rgb.shape

    def run(planes_of_tile, row_stride, pixel_stride, fmt=HYD_UINT8):
        enc = HYDEncoder(product_lib)
        enc.check(enc.set_metadata(w, h, 0, 0, 0))
        obuf = np.empty(1 << 20, np.uint8)
        enc.check(enc.provide_output_buffer(obuf))
        out = bytearray()
        for ty in range((h + 255) // 256):
            for tx in range((w + 255) // 256):
                enc.check(enc.send_tile(planes_of_tile(tx, ty), tx, ty, row_stride, pixel_stride, -1, fmt))
                enc.check(enc.flush())
                _, n = enc.release_output_buffer()
                out += obuf[:n].tobytes()
                enc.check(enc.provide_output_buffer(obuf))
        enc.destroy()
        return bytes(out)

    planar = np.ascontiguousarray(rgb.transpose(2, 0, 1))
    base = planar.ctypes.data
    assert run(lambda tx, ty: tuple(base + c * w * h + ty * 256 * w + tx * 256 for c in range(3)), w, 1) == want, "planar"
    rgba = np.concatenate([rgb, np.full((h, w, 1), 77, np.uint8)], axis=2).copy()
    b4 = rgba.ctypes.data
    assert run(lambda tx, ty: tuple(b4 + (ty * 256 * w + tx * 256) * 4 + c for c in range(3)), w * 4, 4) == want, "rgba"
    bgr = np.ascontiguousarray(rgb[:, :, ::-1])
    b3 = bgr.ctypes.data
    assert run(lambda tx, ty: tuple(b3 + (ty * 256 * w + tx * 256) * 3 + c for c in (2, 1, 0)), w * 3, 3) == want, "bgr"
    flipped = np.ascontiguousarray(rgb[::-1])
    bf = flipped.ctypes.data
    assert run(lambda tx, ty: tuple(bf + ((h - 1 - ty * 256) * w + tx * 256) * 3 + c for c in range(3)), -w * 3, 3) == want, "negative row stride"
    wide = np.zeros((h, w, 7), np.uint8)
    wide[:, :, [0, 3, 6]] = rgb
    bw = wide.ctypes.data
    assert run(lambda tx, ty: tuple(bw + (ty * 256 * w + tx * 256) * 7 + c for c in (0, 3, 6)), w * 7, 7) == want, "stride 7"
    rgba16 = np.zeros((h, w, 4), np.uint16)
    img16 = synth_image(w, h, 16, seed=12)
    rgba16[:, :, :3] = img16
    b16 = rgba16.ctypes.data
    got = run(lambda tx, ty: tuple(b16 + ((ty * 256 * w + tx * 256) * 4 + c) * 2 for c in range(3)), w * 4, 4, HYD_UINT16)
    assert got == oracle.encode_image(img16), "rgba16 (the CLI's 16-bit layout, hydrium.c:445-449)"


def test_float32_samples(product_lib, engine, oracle):
    """HYD_FLOAT32 input (reference: format.c:111-140): golden known answers, the nine-symbol API in
    packed-RGB and planar layouts, and the rejection of non-finite samples."""
    t = kat_table()
    for key in ("P_300x260_f32_srgb", "R_270x130_f32_linear"):
        img = kat_image(t[key])
        assert img.dtype == np.float32
        out = engine.encode_image(img, linear_light=t[key]["linear_light"])
        assert len(out) == t[key]["length"] and sha256(out) == t[key]["sha256"], key
        assert encode_cli_loop(product_lib, img, linear_light=t[key]["linear_light"]) == out, key
    img = kat_image(t["P_300x260_f32_srgb"])
    want = oracle.encode_image(img)
    planar = np.ascontiguousarray(img.transpose(2, 0, 1))            # three float planes
    h, w, _ = img.shape
    enc = HYDEncoder(product_lib)
    out = bytearray()
    obuf = np.empty(1 << 20, np.uint8)
    enc.check(enc.set_metadata(w, h, 0, 0, 0))
    enc.check(enc.provide_output_buffer(obuf))
    for ty in range((h + 255) // 256):
        for tx in range((w + 255) // 256):
            base = planar.ctypes.data + (ty * 256 * w + tx * 256) * 4
            enc.check(enc.send_tile(tuple(base + c * h * w * 4 for c in range(3)), tx, ty, w, 1, -1, HYD_FLOAT32))
            while True:
                ret = enc.flush()
                _, written = enc.release_output_buffer()
                out += obuf[:written].tobytes()
                enc.check(enc.provide_output_buffer(obuf))
                if ret != HYD_NEED_MORE_OUTPUT:
                    break
    enc.destroy()
    assert bytes(out) == want
    # a NaN or an infinity anywhere in a tile is refused with the reference's message
    for poison in (np.nan, np.inf, -np.inf):
        bad = img.copy()
        bad[200, 17, 2] = poison
        with pytest.raises(HydriumError) as ei:
            encode_cli_loop(product_lib, bad)
        assert ei.value.code == HYD_API_ERROR and ei.value.message == "Invalid NaN Float"
        with pytest.raises(HydriumError) as ei:
            engine.encode_image(bad)
        assert ei.value.message == "Invalid NaN Float"
    assert engine.encode_image(img) == want   # the engine is still usable afterwards
    # a negative opsin mix (here: a strongly negative red sample) makes the reference's bit-hack cube root
    # return NaN, whose conversion to int is undefined in C: refused loudly instead of guessing
    bad = img.copy()
    bad[10, 10, 0] = -3.0
    with pytest.raises(HydriumError) as ei:
        encode_cli_loop(product_lib, bad)
    assert ei.value.code == HYD_API_ERROR and "negative opsin mix" in ei.value.message
    slightly = img.copy()
    slightly[10, 10, 0] = -0.001   # the mix stays above -bias: still encodable, still the reference's bytes
    assert encode_cli_loop(product_lib, slightly) == oracle.encode_image(slightly)


def _split_frames(stream: bytes, header_len: int):
    """(for diagnostics) nothing format-aware: just report the first differing byte offset"""
    return stream[header_len:]


@pytest.mark.parametrize("case", [
    # (width, height, shift_x, shift_y, bits, linear, seed)
    (600, 520, 1, 1, 8, 0, 1),        # 2 x 2 tiles of 512^2: frames of 4, 2, 2 groups and a partial one
    (700, 300, 2, 0, 8, 0, 2),        # tiles 1024 x 256: one frame of 3 groups in a row
    (300, 700, 0, 2, 16, 1, 3),       # tiles 256 x 1024: a column of 3 groups, 16-bit linear
    (1100, 600, 3, 3, 8, 0, 4),       # one 2048^2 tile: 5 x 3 groups
    (520, 260, -1, -1, 8, 0, 5),      # one-frame mode, one LF group, 3 x 2 groups
    (256, 257, 1, 1, 8, 0, 6),        # second group one pixel row high
])
def test_multi_group_frames_match_reference(product_lib, reflib, case):
    """tile_size_shift 1..3 and one-frame mode within one LF group: the frame
    carries LFGlobal / LFGroup / HFGlobal sections shared by its groups, a TOC entry per section and
    one ANS model for all groups.  Compared with the reference library itself."""
    w, h, sx, sy, bits, lin, seed = case
    img = synth_image(w, h, bits, seed=seed)
    want = encode_cli_loop(reflib, img, linear_light=lin, shift_x=sx, shift_y=sy)
    got = encode_cli_loop(product_lib, img, linear_light=lin, shift_x=sx, shift_y=sy)
    if got != want:
        n = min(len(got), len(want))
        first = next((i for i in range(n) if got[i] != want[i]), n)
        raise AssertionError(f"{case}: {len(got)} vs {len(want)} bytes, first difference at byte {first}: "
                             f"{got[max(0, first - 4):first + 12].hex()} vs {want[max(0, first - 4):first + 12].hex()}")


@pytest.mark.parametrize("case", [
    # (width, height, bits, linear, seed, tile order or None)
    (2304, 2100, 8, 0, 0, None),                      # 2 x 2 LF groups (the golden G case)
    (4200, 300, 8, 0, 3, None),                       # three LF groups in a row, the last 104 px wide
    (300, 4200, 16, 1, 4, None),                      # a column of three, 16-bit linear
    (2100, 2060, 8, 0, 5, [(1, 1), (0, 0), (1, 0), (0, 1)]),   # sent out of raster order
    (2048 * 28 + 5, 9, 8, 0, 6, None),                # 29 LF groups: three HF clusters per preset
    (2048 * 85 + 300, 8, 8, 0, 7, None),              # 86 LF groups: two clusters per preset
    (2048 * 128 + 1, 5, 8, 0, 8, None),               # 129 LF groups: one cluster per preset
    # (256 LF groups run through the same one-cluster path; the reference itself needs many minutes there)
])
def test_one_frame_mode_over_several_lf_groups(product_lib, reflib, case):
    """The reference CLI's default mode for images beyond 2048x2048: one frame, a
    preset of nine ANS clusters per LF group, TOC permuted by send order, output only at the end."""
    w, h, bits, lin, seed, order = case
    img = synth_image(w, h, bits, seed=seed, smooth=True)

    def run(lib):
        if order is None:
            return encode_cli_loop(lib, img, linear_light=lin, shift_x=-1, shift_y=-1)
        return _encode_in_order(lib, img, lin, order)

    want, got = run(reflib), run(product_lib)
    if got != want:
        n = min(len(got), len(want))
        first = next((i for i in range(n) if got[i] != want[i]), n)
        raise AssertionError(f"{case[:5]}: {len(got)} vs {len(want)} bytes, first difference at byte {first}: "
                             f"{got[max(0, first - 4):first + 12].hex()} vs {want[max(0, first - 4):first + 12].hex()}")
    if (w, h) == (2304, 2100):
        e = kat_table()["G_2304x2100_oneframe_ref_only"]
        assert e["smooth"] is False   # the golden entry is the noisy variant: checked below
        noisy = kat_image(e)
        out = encode_cli_loop(product_lib, noisy, shift_x=-1, shift_y=-1)
        assert len(out) == e["length"] and sha256(out) == e["sha256"]


def _encode_in_order(lib, img, lin, order):
    """one-frame mode, LF groups sent in the given order, is_last set on the final call only"""
    h, w, ch = img.shape
    item = img.dtype.itemsize
    enc = HYDEncoder(lib)
    obuf = np.empty(1 << 20, np.uint8)
    out = bytearray()
    enc.check(enc.set_metadata(w, h, lin, -1, -1))
    enc.check(enc.provide_output_buffer(obuf))
    for i, (tx, ty) in enumerate(order):
        p = img.ctypes.data + (ty * 2048 * w * ch + tx * 2048 * ch) * item
        enc.check(enc.send_tile((p, p + item, p + 2 * item), tx, ty, w * ch, ch, int(i == len(order) - 1), E._fmt_of(img)))
        while True:
            ret = enc.flush()
            _, written = enc.release_output_buffer()
            out += obuf[:written].tobytes()
            enc.check(enc.provide_output_buffer(obuf))
            if ret != HYD_NEED_MORE_OUTPUT:
                break
        enc.check(ret)
    enc.destroy()
    return bytes(out)


def test_multi_group_golden_and_float(product_lib, reflib):
    e = kat_table()["H_1024_tile512_ref_only"]
    out = encode_cli_loop(product_lib, kat_image(e), shift_x=e["shift"], shift_y=e["shift"])
    assert len(out) == e["length"] and sha256(out) == e["sha256"]
    img = (synth_image(530, 300, 16, seed=8).astype(np.float32) / np.float32(65535))
    assert encode_cli_loop(product_lib, img, shift_x=1, shift_y=1) == encode_cli_loop(reflib, img, shift_x=1, shift_y=1)


def test_far_tiles_of_huge_images(engine, oracle):
    """Frame headers of tiles deep inside config-3 / config-5 sized images, and the level-10 container."""
    buf = np.zeros(128, np.uint8)
    n = engine.lib.hydb_image_header(65536, 65536, buf.ctypes.data, 128)
    assert buf[:n].tobytes() == oracle.image_header(65536, 65536) and n == 59
    d_out = engine.device_alloc(1 << 20)
    try:
        for (iw, ih, tx, ty, bits) in [(65536, 65536, 255, 255, 8), (65536, 65536, 200, 3, 8), (16384, 16384, 63, 63, 16),
                                       (16384, 16384, 9, 40, 16), (1 << 20, 300, 4095, 1, 8)]:
            tw, th = min(256, iw - tx * 256), min(256, ih - ty * 256)
            win = synth_image(tw, th, bits, x0=tx * 256, y0=ty * 256, full_width=iw, full_height=ih)
            d_img = engine.upload(win)
            lin = int(bits == 16)
            got = engine.download(d_out, engine.encode_tiles([_tile_desc(d_img, win, tx, ty, lin, (iw, ih), origin=True)], d_out, 1 << 20))
            engine.device_free(d_img)
            assert got == oracle.encode_tile(win, tx, ty, linear_light=lin, image_size=(iw, ih), window=True), (iw, tx, ty)
    finally:
        engine.device_free(d_out)


def test_device_synth_generator_matches_numpy(engine):
    for (w, h, bits, smooth, seed) in [(300, 200, 8, False, 0), (129, 77, 16, False, 5), (256, 256, 8, True, 2), (64, 64, 16, True, 9)]:
        n = w * h * 3 * (bits // 8)
        d = engine.device_alloc(n)
        engine.synth_fill(d, w, h, bits=bits, seed=seed, smooth=smooth)
        got = np.frombuffer(engine.download(d, n), np.uint8 if bits == 8 else np.uint16).reshape(h, w, 3)
        engine.device_free(d)
        assert np.array_equal(got, synth_image(w, h, bits, seed=seed, smooth=smooth)), (w, h, bits, smooth)
    d = engine.device_alloc(256 * 256 * 3)
    engine.synth_fill(d, 256, 256, bits=8, x0=256 * 200, y0=256 * 3, full_width=65536, full_height=65536)
    got = np.frombuffer(engine.download(d, 256 * 256 * 3), np.uint8).reshape(256, 256, 3)
    engine.device_free(d)
    assert np.array_equal(got, synth_image(256, 256, 8, x0=256 * 200, y0=256 * 3, full_width=65536, full_height=65536))


def test_config2_full_size(product_lib, oracle):
    """BASELINE configs[1] at full size: 4096x4096 sRGB8, 256 tiles, full compare with the oracle,
    plus size-independent properties: tile frames are position- but not order-dependent, and a
    band-sharded encode concatenates to the same stream (what the multi-GPU run relies on)."""
    w = h = 4096
    with E.Engine(device=0, max_batch_tiles=256) as eng:
        n_in = w * h * 3
        d_in = eng.device_alloc(n_in)
        cap = E.output_bound(w, h)
        d_out = eng.device_alloc(cap)
        eng.synth_fill(d_in, w, h, bits=8, seed=0)
        n = eng.encode_image_device(d_in, w, h, 3, d_out=d_out, d_out_cap=cap)
        full = eng.download(d_out, n)
        img = np.frombuffer(eng.download(d_in, n_in), np.uint8).reshape(h, w, 3)
        assert np.array_equal(img, synth_image(w, h, 8))
        assert full == oracle.encode_image(img)
        # idempotence
        assert eng.download(d_out, eng.encode_image_device(d_in, w, h, 3, d_out=d_out, d_out_cap=cap)) == full
        # sharded by tile rows == whole (ranks own contiguous bands; only rank 0 writes the header)
        parts = []
        for r in range(4):
            band = d_in + r * 4 * 256 * w * 3
            m = eng.encode_image_device(band, w, h, 3, tile_rows=(r * 4, r * 4 + 4), with_header=(r == 0), d_out=d_out, d_out_cap=cap)
            parts.append(eng.download(d_out, m))
        assert b"".join(parts) == full
        eng.device_free(d_in)
        eng.device_free(d_out)


def test_lf_stream_stress(engine, oracle):
    """Blocky images (constant 8x8 blocks) drive the LF path hard: long runs, many distinct
    residual tokens, skewed histograms that hit the depth limiter of the code-length builder."""
    rng = np.random.default_rng(17)
    fib = [1, 1]
    while len(fib) < 24:
        fib.append(fib[-1] + fib[-2])
    cases = []
    for kind in range(8):
        bw_, bh_ = int(rng.integers(1, 33)), int(rng.integers(1, 33))
        if kind == 0:
            lv = rng.integers(0, 256, (bh_, bw_, 3))
        elif kind == 1:
            lv = np.repeat(rng.integers(0, 256, (bh_, 1, 3)), bw_, axis=1)          # constant rows: long runs
        elif kind == 2:
            lv = rng.choice([0, 255], (bh_, bw_, 3))
        elif kind == 3:
            lv = (rng.geometric(0.08, (bh_, bw_, 3)) - 1).clip(0, 255)
        elif kind == 4:   # Fibonacci-distributed levels
            pool = np.concatenate([np.full(min(f, 300), (i * 11) % 256) for i, f in enumerate(fib)])
            lv = rng.choice(pool, (32, 32, 3))
        elif kind == 5:
            lv = np.zeros((32, 32, 3), np.int64)
            lv[::5, ::3] = 200
        elif kind == 6:
            lv = np.cumsum(rng.integers(-3, 4, (32, 32, 3)), axis=1).clip(0, 255) + 100
        else:
            lv = np.tile(np.arange(32)[None, :, None] * 8, (32, 1, 3))
        img = np.repeat(np.repeat(lv.astype(np.uint8 if kind != 6 else np.uint16), 8, axis=0), 8, axis=1)
        if img.dtype == np.uint16:
            img = (img.astype(np.uint32) * 200).clip(0, 65535).astype(np.uint16)
        cases.append(np.ascontiguousarray(img))
    for i, img in enumerate(cases):
        lin = int(img.dtype == np.uint16)
        assert engine.encode_image(img, linear_light=lin) == oracle.encode_image(img, linear_light=lin), i


def test_config3_slice_16bit_linear(product_lib, oracle):
    """BASELINE configs[2] in miniature: 16-bit linear RGB, packed (stride 3) and RGBA (stride 4,
    the CLI's layout), several tile rows, sharded into bands like the 8-GPU run."""
    w, h = 1024, 768
    img = synth_image(w, h, 16, seed=4)
    want = oracle.encode_image(img, linear_light=1)
    with E.Engine(device=0, max_batch_tiles=5) as eng:
        assert eng.encode_image(img, linear_light=1) == want
        rgba = np.zeros((h, w, 4), np.uint16)
        rgba[:, :, :3] = img
        assert eng.encode_image(rgba, linear_light=1) == want
        d_in = eng.upload(img)
        cap = E.output_bound(w, h)
        d_out = eng.device_alloc(cap)
        parts = []
        for r in range(3):
            band = d_in + r * 256 * w * 3 * 2
            n = eng.encode_image_device(band, w, h, 3, sample_fmt=HYD_UINT16, linear_light=1, tile_rows=(r, r + 1),
                                        with_header=(r == 0), d_out=d_out, d_out_cap=cap)
            parts.append(eng.download(d_out, n))
        assert b"".join(parts) == want
        eng.device_free(d_in)
        eng.device_free(d_out)


def test_config4_batch_of_frames(product_lib, oracle):
    """BASELINE configs[3] in miniature: independent 1920x1080 frames (8 x 5 tiles, partial right
    column and bottom row), each its own codestream, tiles of different frames sharing launches."""
    w, h, count = 1920, 1080, 5
    frames = [synth_image(w, h, 8, seed=100 + k) for k in range(count)]
    want = [oracle.encode_image(f) for f in frames]
    with E.Engine(device=0, max_batch_tiles=96) as eng:
        d_in = eng.upload(np.stack(frames))
        cap = count * E.output_bound(w, h)
        d_out = eng.device_alloc(cap)
        spans = eng.encode_image_batch(d_in, count, w, h, 3, d_out=d_out, d_out_cap=cap)
        assert len(spans) == count
        for k, (off, n) in enumerate(spans):
            assert eng.download(d_out + off, n) == want[k], k
        eng.device_free(d_in)
        eng.device_free(d_out)
    # the same through one nine-symbol encoder per frame (engine is parked and reused between them)
    import os
    os.environ["HYDRIUM_B200_BATCH"] = "40"
    try:
        for k in range(2):
            assert encode_cli_loop(product_lib, frames[k]) == want[k]
    finally:
        del os.environ["HYDRIUM_B200_BATCH"]


def test_config5_streaming_window(product_lib, oracle):
    """BASELINE configs[4] in miniature: a band of the 65536x65536 image generated on the device and
    streamed through the engine in several launches; level-10 container header on the first tile."""
    iw = ih = 65536
    ntx = 6
    with E.Engine(device=0, max_batch_tiles=4) as eng:
        band_w = ntx * 256
        d_in = eng.device_alloc(band_w * 256 * 3)
        eng.synth_fill(d_in, band_w, 256, bits=8, x0=0, y0=0, full_width=iw, full_height=ih)
        tiles = []
        for tx in range(ntx):
            t = HydbTile()
            p = d_in + tx * 256 * 3
            t.plane = (C.c_void_p * 3)(p, p + 1, p + 2)
            t.row_stride, t.pixel_stride = band_w * 3, 3
            t.x0, t.y0, t.width, t.height = tx * 256, 0, 256, 256
            t.image_width, t.image_height = iw, ih
            t.sample_fmt, t.linear_light, t.is_last = HYD_UINT8, 0, 0
            t.with_image_header = int(tx == 0)
            tiles.append(t)
        cap = 8 << 20
        d_out = eng.device_alloc(cap)
        pos = 0
        for i in range(0, ntx, 4):
            pos += eng.encode_tiles(tiles[i:i + 4], d_out, cap, pos)
        got = eng.download(d_out, pos)
        win = synth_image(band_w, 256, 8, full_width=iw, full_height=ih)
        want = oracle.image_header(iw, ih)
        for tx in range(ntx):
            want += oracle.encode_tile(np.ascontiguousarray(win[:, tx * 256:(tx + 1) * 256]), tx, 0, image_size=(iw, ih), window=True)
        assert got == want
        eng.device_free(d_in)
        eng.device_free(d_out)


def _fake_icc(size: int, seed: int, vendor: bytes = b"APPL") -> bytes:
    """A profile-shaped byte string: a plausible 128-byte header (so that the header predictor mostly
    hits), a tag table and a mix of text, small numbers and noise behind it, cut to `size` bytes."""
    rng = np.random.default_rng(seed)
    head = bytearray(128)
    head[0:4] = size.to_bytes(4, "big")
    head[4:8] = b"lcms"
    head[8:12] = bytes([4, 0x30, 0, 0])
    head[12:24] = b"mntrRGB XYZ "
    head[24:36] = bytes([7, 0xE6, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0])
    head[36:40] = b"acsp"
    head[40:44] = vendor
    head[68:80] = bytes([0, 0, 0xF6, 0xD6, 0, 1, 0, 0, 0, 0, 0xD3, 0x2D])
    head[80:84] = head[4:8]
    body = bytearray()
    body += (9).to_bytes(4, "big")
    for k, tag in enumerate([b"desc", b"cprt", b"wtpt", b"rXYZ", b"gXYZ", b"bXYZ", b"rTRC", b"gTRC", b"bTRC"]):
        body += tag + (128 + 4 + 9 * 12 + 40 * k).to_bytes(4, "big") + (40).to_bytes(4, "big")
    while len(body) + 128 < size:
        kind = int(rng.integers(0, 4))
        if kind == 0:
            body += b"Copyright 2026, no rights reserved. sRGB IEC61966-2.1, v1.0 "
        elif kind == 1:
            body += bytes(rng.integers(0, 16, 48, dtype=np.uint8))
        elif kind == 2:
            body += bytes(rng.integers(0, 256, 64, dtype=np.uint8))
        else:
            body += bytes(int(rng.integers(0, 2)) * 255 if i % 3 else 0 for i in range(40)) + bytes(24)
    return bytes(head + body)[:size]


@pytest.mark.parametrize("case", [
    # (width, height, profile size, vendor, seed)
    (64, 48, 60, b"APPL", 1),          # profile shorter than its 128-byte header: no command stream
    (64, 48, 128, b"MSFT", 2),         # exactly the header
    (64, 48, 129, b"ADBE", 3),         # one byte behind the header
    (200, 120, 3144, b"MSFT", 4),      # the size of the common sRGB profile
    (520, 260, 524, b"none", 5),       # several groups in the one frame
    (2100, 300, 70000, b"APPL", 6),    # two LF groups; a large profile (lookup tables)
])
def test_icc_tagged_image_header_matches_reference(product_lib, reflib, case):
    """hyd_set_suggested_icc_profile (libhydrium.c:242-305) + the ICC branch of the image header
    (encoder.c:122-162, 203-236): mangled profile, 41-context / 9-cluster prefix stream, written by
    k_icc_header on the device."""
    w, h, size, vendor, seed = case
    img = synth_image(w, h, 8, seed=seed)
    icc = _fake_icc(size, seed, vendor)
    want = encode_cli_loop(reflib, img, shift_x=-1, shift_y=-1, icc=icc)
    got = encode_cli_loop(product_lib, img, shift_x=-1, shift_y=-1, icc=icc)
    plain = encode_cli_loop(product_lib, img, shift_x=-1, shift_y=-1)
    assert len(got) > len(plain) and got != plain
    if got != want:
        n = min(len(got), len(want))
        first = next((i for i in range(n) if got[i] != want[i]), n)
        raise AssertionError(f"{case}: {len(got)} vs {len(want)} bytes, first difference at byte {first}: "
                             f"{got[max(0, first - 4):first + 12].hex()} vs {want[max(0, first - 4):first + 12].hex()}")


def test_icc_vendors_the_reference_crashes_on(product_lib):
    """Profiles whose platform signature is "SGI " or "SUNW" make the reference index a two-byte string
    with (unsigned)(41 - 42) and segfault (libhydrium.c:225-230).  Here they are predicted as the format
    says; the frame behind the header is the untagged image's."""
    img = synth_image(64, 48, 8, seed=2)
    plain = encode_cli_loop(product_lib, img, shift_x=-1, shift_y=-1)
    base = encode_cli_loop(product_lib, img, shift_x=-1, shift_y=-1, icc=_fake_icc(700, 5, b"APPL"))
    for vendor in (b"SGI ", b"SUNW"):
        got = encode_cli_loop(product_lib, img, shift_x=-1, shift_y=-1, icc=_fake_icc(700, 5, vendor))
        frame = len(plain) - 16
        assert got[-frame:] == plain[-frame:]
        assert abs(len(got) - len(base)) <= 4   # the vendor is predicted, like "APPL"


def test_icc_profile_cleared_and_refused_in_tile_mode(product_lib):
    img = synth_image(64, 48, 8, seed=9)
    enc = HYDEncoder(product_lib)
    assert enc.set_metadata(64, 48, 0, 0, 0) == 0
    assert enc.set_suggested_icc_profile(_fake_icc(300, 1)) == HYD_API_ERROR
    assert "one-frame mode required" in enc.error_message_get()
    enc.destroy()
    # set, then cleared: the codestream is the untagged one
    plain = encode_cli_loop(product_lib, img, shift_x=-1, shift_y=-1)
    enc = HYDEncoder(product_lib)
    enc.check(enc.set_metadata(64, 48, 0, -1, -1))
    enc.check(enc.set_suggested_icc_profile(_fake_icc(300, 1)))
    enc.check(enc.set_suggested_icc_profile(None))
    obuf = np.empty(1 << 20, np.uint8)
    enc.check(enc.provide_output_buffer(obuf))
    p = img.ctypes.data
    enc.check(enc.send_tile((p, p + 1, p + 2), 0, 0, 64 * 3, 3, -1, 0))
    assert enc.flush() != HYD_NEED_MORE_OUTPUT
    _, written = enc.release_output_buffer()
    enc.destroy()
    assert obuf[:written].tobytes() == plain


@pytest.mark.gpu
def test_recurring_chunks_replayed_as_cuda_graphs(product_lib, oracle):
    """A chunk geometry seen for the second time is captured as a CUDA graph and replayed from then on
    (engine.cu, hydb_engine_submit_frames): same bytes as the oracle every time, the replay counter moves
    from the third encode on, and the kernel counter keeps counting replayed kernels.  HYDRIUM_B200_GRAPHS=0
    is read once per process, so the switch is not exercised here."""
    img = synth_image(2304, 512, 8, seed=21)   # 18 tiles: one chunk
    want = oracle.encode_image(img)
    seen = []
    for _ in range(4):
        st = {}
        assert encode_cli_loop(product_lib, img, stats=st) == want
        seen.append(st)
    assert seen[3]["graph_launches"] > seen[1]["graph_launches"], seen
    per_encode = seen[3]["kernel_launches"] - seen[2]["kernel_launches"]
    assert per_encode == seen[1]["kernel_launches"] - seen[0]["kernel_launches"] > 0, seen
    # a different image of the same geometry through the replayed graph
    img2 = synth_image(2304, 512, 8, seed=22)
    assert encode_cli_loop(product_lib, img2) == oracle.encode_image(img2)


@pytest.mark.gpu
def test_host_side_under_address_sanitizer(product_lib, tmp_path):
    """hyd_api.c + stage_pool.c + the C harness built with -fsanitize=address,undefined and linked with the same
    CUDA objects (make asan): tile mode, larger tiles, one-frame mode over several LF groups and a ragged image
    run clean and write the same bytes as the regular build."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "hydrium_b200", "csrc"), "asan"], check=True, capture_output=True)
    plain = os.path.join(root, "hydrium_b200", "bin", "api_bench")
    asan = os.path.join(root, "hydrium_b200", "bin", "api_bench_asan")
    env = dict(os.environ, ASAN_OPTIONS="protect_shadow_gap=0:detect_leaks=0:abort_on_error=0", UBSAN_OPTIONS="halt_on_error=1")
    cases = [["--width", "2304", "--height", "1100"], ["--width", "1300", "--height", "700", "--shift", "1"],
             ["--width", "4096", "--height", "2100", "--one-frame"], ["--width", "300", "--height", "260", "--bits", "16", "--linear"]]
    for i, case in enumerate(cases):
        a, b = str(tmp_path / f"a{i}.jxl"), str(tmp_path / f"b{i}.jxl")
        p = subprocess.run([asan, "--reps", "3", "--warmup", "1", "--out", a] + case, env=env, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0 and "ERROR: AddressSanitizer" not in p.stderr and "runtime error" not in p.stderr, (case, p.stderr[-3000:])
        subprocess.run([plain, "--reps", "1", "--warmup", "1", "--out", b] + case, check=True, capture_output=True, timeout=600)
        assert open(a, "rb").read() == open(b, "rb").read(), case


@pytest.mark.gpu
def test_one_frame_head_streams_reused_across_images(product_lib, reflib):
    """The engine keeps the bits of the HF context map's stream (a function of the number of LF groups) and of
    the TOC permutation's stream (geometry + send order) from one image to the next (k_oneframe_finish).  Same
    geometry with different pixels, another send order, back again, another geometry, a first image (with image
    header) against the same cache: always the reference's bytes."""
    a = synth_image(2304, 2100, 8, seed=31, smooth=True)
    b = synth_image(2304, 2100, 8, seed=32, smooth=True)
    c = synth_image(4200, 300, 8, seed=33, smooth=True)
    raster = [(0, 0), (1, 0), (0, 1), (1, 1)]
    other = [(1, 1), (0, 0), (1, 0), (0, 1)]
    for img, order in ((a, raster), (b, raster), (a, other), (b, other), (b, raster), (c, [(0, 0), (1, 0), (2, 0)]),
                       (c, [(2, 0), (1, 0), (0, 0)]), (a, raster)):
        assert _encode_in_order(product_lib, img, 0, order) == _encode_in_order(reflib, img, 0, order), (img.shape, order)
