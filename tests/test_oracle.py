"""The oracle (oracle/hyd_oracle.c) against the committed golden vectors and, where the reference
build is present, against the unmodified reference itself: final bytes and all stage taps."""
import numpy as np
import pytest

from hydrium_b200.encoder import encode_cli_loop
from oracle.pyoracle import Stages
from util import GOLDEN, kat_image, kat_table, sha256, image_set

ORACLE_SCOPE = [k for k, v in kat_table().items() if "width" in v and "ref_only" not in k]


@pytest.mark.parametrize("name", ORACLE_SCOPE)
def test_oracle_matches_golden_kat(oracle, name):
    e = kat_table()[name]
    out = oracle.encode_image(kat_image(e), linear_light=e["linear_light"])
    assert len(out) == e["length"]
    assert sha256(out) == e["sha256"]


def test_oracle_matches_golden_files(oracle):
    t = kat_table()
    for fname, key in [("synth_40x24.jxl", "T_40x24"), ("synth_300x260.jxl", "Q_300x260")]:
        want = open(f"{GOLDEN}/{fname}", "rb").read()
        assert oracle.encode_image(kat_image(t[key]), linear_light=t[key]["linear_light"]) == want
    flat = np.full((512, 512, 3), 128, np.uint8)
    assert oracle.encode_tile(flat, 0, 0, is_last=0) == open(f"{GOLDEN}/flat128_tile.bin", "rb").read()


@pytest.mark.parametrize("name", list(kat_table().keys()))
def test_reference_reproduces_golden(reflib, name):
    """The fixtures really are what the reference emits (guards against a stale kat.json)."""
    e = kat_table()[name]
    if "width" not in e:
        pytest.skip("tile fixture")
    if e["width"] * e["height"] > 1 << 21:
        pytest.skip("large case: covered by make_golden.py")
    out = encode_cli_loop(reflib, kat_image(e), linear_light=e["linear_light"], shift_x=e["shift"], shift_y=e["shift"])
    assert sha256(out) == e["sha256"]


def test_oracle_stages_match_reference_taps(oracle, reftap):
    rng = np.random.default_rng(11)
    for name, img, lin in image_set(rng):
        h, w, _ = img.shape
        hdr = oracle.image_header(w, h)
        for ty in range((h + 255) // 256):
            for tx in range((w + 255) // 256):
                a, b = Stages(), Stages()
                ref = reftap.encode_tile(img, tx, ty, linear_light=lin, stages=a)
                got = oracle.encode_tile(img, tx, ty, linear_light=lin, stages=b)
                n = a.vbw * a.vbh * 64 * 3
                assert ref == hdr + got, (name, tx, ty)
                assert np.array_equal(a.xyb[:n].view(np.uint32), b.xyb[:n].view(np.uint32)), (name, "xyb")
                assert np.array_equal(a.dct[:n].view(np.uint32), b.dct[:n].view(np.uint32)), (name, "dct")
                assert np.array_equal(a.quant[:n], b.quant[:n]), (name, "quant")
                assert np.array_equal(a.nonzeroes, b.nonzeroes), (name, "nz")
                assert a.n_syms == b.n_syms and np.array_equal(a.hf_syms[:a.n_syms], b.hf_syms[:b.n_syms]), (name, "syms")
                assert np.array_equal(a.freqs, b.freqs), (name, "freqs")
                assert a.lf_bitlen == b.lf_bitlen and a.bits("lf") == b.bits("lf"), (name, "lf")
                assert a.pre_bitlen == b.pre_bitlen and a.bits("pre") == b.bits("pre"), (name, "pre")
                assert a.ans_bitlen == b.ans_bitlen and a.bits("ans") == b.bits("ans"), (name, "ans")


def test_oracle_whole_images_match_reference(oracle, reflib):
    rng = np.random.default_rng(5)
    for name, img, lin in image_set(rng):
        assert oracle.encode_image(img, linear_light=lin) == encode_cli_loop(reflib, img, linear_light=lin), name


def test_oracle_luts_and_cosines_match_reference(oracle, reftap):
    for fmt in (0, 1):
        for lin in (0, 1):
            a, b = oracle.luts(fmt, lin), reftap.luts(fmt, lin)
            assert np.array_equal(a[0], b[0])
            assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    cos = reftap.cosine_lut().view(np.uint32)
    assert cos[0, 0] == 0x3e318a87 and cos[3, 1] == 0xbe000000 and cos[6, 7] == 0xbd0d42a9


def test_oracle_prefix_streams_match_reference(oracle, reftap):
    rng = np.random.default_rng(7)
    cfgs = [dict(custom=(7, 1, 1), lz77_min_symbol=1 << 14, modular=1), dict(custom=None, lz77_min_symbol=0, modular=0),
            dict(custom=None, lz77_min_symbol=29, modular=1), dict(custom=(4, 1, 0), lz77_min_symbol=64, modular=0)]
    for trial in range(120):
        n = int(rng.integers(1, 3100))
        kind = trial % 6
        if kind == 0:
            vals = rng.integers(0, 4000, n)
        elif kind == 1:
            vals = np.repeat(rng.integers(0, 50, n // 7 + 1), rng.integers(1, 300, n // 7 + 1))[:n]
        elif kind == 2:
            vals = np.zeros(n, np.int64)
        elif kind == 3:
            vals = rng.integers(0, 3, n)
        elif kind == 4:
            vals = np.repeat(rng.integers(0, 2000, n // 3 + 1), rng.integers(1, 6, n // 3 + 1))[:n]
        else:
            vals = rng.geometric(0.05, n) - 1
        for cfg in cfgs:
            v = vals if cfg["lz77_min_symbol"] not in (29, 64) else np.minimum(vals, 15)
            assert oracle.prefix_stream(v, **cfg) == reftap.prefix_stream(v, **cfg), (trial, cfg)


def test_image_headers(oracle):
    # SURVEY.md Appendix A, measured on the reference
    want = {(256, 256): "ff0af807fe814c", (700, 600): "ff0aba126857804c", (1024, 1024): "ff0afa1fe87f804c",
            (1920, 1080): "ff0aba21e8ef804c", (4096, 4096): "ff0afa7fe8ff814c", (16384, 16384): "ff0afcff01feff003201"}
    for (w, h), hx in want.items():
        assert oracle.image_header(w, h).hex() == hx
    big = oracle.image_header(65536, 65536)
    assert len(big) == 59 and big[49:].hex() == "ff0afcff07feff033201" and big[4:8] == b"JXL "


def test_bench_sampled_tile_reference_driver(reflib):
    """bench.py's parity leg drives the unmodified reference with the full image's metadata and only a
    few tiles (gaps are legal, libhydrium.h:235-240) and cuts our stream at the engine's frame lengths.
    Here, on the CPU: the frames that driver returns, in any order and from a window of the image, are
    exactly the pieces of the reference's own whole-image stream."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from hydrium_b200.encoder import encode_cli_loop
    from hydrium_b200.synth import synth_image
    img = synth_image(1100, 800, 8, seed=3)            # 5 x 4 tiles, partial right / bottom tiles
    whole = encode_cli_loop(reflib, img)
    tiles = [(x, y) for y in range(4) for x in range(5)]
    frames = bench.ref_encode_tiles(reflib, img, 1100, 800, tiles)
    assert b"".join(frames) == whole
    # a subset in another order: same frames (position-, not order-dependent), image header stripped
    sub = [(4, 3), (0, 0), (2, 1)]
    got = bench.ref_encode_tiles(reflib, img, 1100, 800, sub)
    hdr = len(frames[0]) - len(bench.ref_encode_tiles(reflib, img, 1100, 800, [(1, 0), (0, 0)])[1])
    assert got[0] == frames[19] and got[2] == frames[7]
    assert got[1] == frames[0][hdr:]                   # (0, 0) sent second carries no image header
    # a window of the image whose upper-left tile is (1, 2), 16-bit linear
    img16 = synth_image(700, 900, 16, seed=5)
    all16 = bench.ref_encode_tiles(reflib, img16, 700, 900, [(x, y) for y in range(4) for x in range(3)], linear=1)
    win = np.ascontiguousarray(img16[512:, 256:])
    part = bench.ref_encode_tiles(reflib, win, 700, 900, [(1, 2), (2, 3)], linear=1, origin=(1, 2))
    assert part[0] == all16[2 * 3 + 1] and part[1] == all16[3 * 3 + 2]
