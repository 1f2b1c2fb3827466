"""The command-line front end (hydrium_b200/cli, reference: src/hydrium.c): its own PNG / PFM readers
and the reference CLI's call sequence.  The same C sources are linked twice -- against the product
library and against the unmodified reference build -- and must write the same file; the reference-
linked binary is also checked against the library driven directly with the pixels the PNG should
decode to (what libspng hands the reference CLI: RGB8 for depths <= 8, RGBA16 for 16 bit)."""
from __future__ import annotations

import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

from hydrium_b200.encoder import encode_cli_loop
from hydrium_b200.synth import synth_image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI_SRC = [os.path.join(ROOT, "hydrium_b200", "cli", f) for f in ("hydrium_cli.c", "png_reader.c")]
CLI_OURS = os.path.join(ROOT, "hydrium_b200", "bin", "hydrium")
CLI_REF = os.path.join(ROOT, "tests", "host_harness", "_build", "hydrium_ref")


# ---- a PNG writer that exercises the reader: every filter type, split IDAT, ancillary chunks ----------
def _chunk(kind: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xFFFFFFFF)


def _filter_rows(rows: np.ndarray, bpp: int) -> bytes:
    """rows: (h, nbytes) uint8 raw scanlines; filter type cycles 0..4 with the row number."""
    h, n = rows.shape
    out = bytearray()
    zero = np.zeros(n, np.int32)
    for y in range(h):
        cur = rows[y].astype(np.int32)
        up = rows[y - 1].astype(np.int32) if y else zero
        left = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]]) if n > bpp else np.zeros(n, np.int32)
        upleft = np.concatenate([np.zeros(bpp, np.int32), up[:-bpp]]) if n > bpp else np.zeros(n, np.int32)
        t = y % 5
        if t == 0:
            pred = zero
        elif t == 1:
            pred = left
        elif t == 2:
            pred = up
        elif t == 3:
            pred = (left + up) >> 1
        else:
            p = left + up - upleft
            pa, pb, pc = np.abs(p - left), np.abs(p - up), np.abs(p - upleft)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, upleft))
        out.append(t)
        out += ((cur - pred) & 0xFF).astype(np.uint8).tobytes()
    return bytes(out)


def _pack_samples(samples: np.ndarray, depth: int) -> np.ndarray:
    """samples: (h, w * channels) integers -> (h, nbytes) uint8 raw scanlines."""
    h, n = samples.shape
    if depth == 16:
        return samples.astype(">u2").view(np.uint8).reshape(h, n * 2)
    if depth == 8:
        return samples.astype(np.uint8)
    per = 8 // depth
    pad = (-n) % per
    s = np.concatenate([samples.astype(np.uint8), np.zeros((h, pad), np.uint8)], axis=1).reshape(h, -1, per)
    out = np.zeros(s.shape[:2], np.uint8)
    for k in range(per):
        out |= s[:, :, k] << (8 - depth * (k + 1))
    return out


_ADAM7 = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]


def write_png(path: str, samples: np.ndarray, color: int, depth: int, interlace: bool = False,
              palette: np.ndarray | None = None) -> None:
    """samples: (h, w, channels) integer samples as stored in the file."""
    h, w, ch = samples.shape
    bpp = max(1, ch * depth // 8)
    if interlace:
        raw = b""
        for (x0, y0, dx, dy) in _ADAM7:
            sub = samples[y0::dy, x0::dx]
            if sub.size:
                raw += _filter_rows(_pack_samples(sub.reshape(sub.shape[0], -1), depth), bpp)
    else:
        raw = _filter_rows(_pack_samples(samples.reshape(h, -1), depth), bpp)
    z = zlib.compress(raw, 6)
    cut = [0, len(z) // 3, len(z) // 3, 2 * len(z) // 3, len(z)]   # three chunks + one empty
    body = _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color, 0, 0, 1 if interlace else 0))
    body += _chunk(b"gAMA", struct.pack(">I", 45455)) + _chunk(b"tEXt", b"Comment\0test image")
    if palette is not None:
        body += _chunk(b"PLTE", palette.astype(np.uint8).tobytes())
    for a, b in zip(cut[:-1], cut[1:]):
        body += _chunk(b"IDAT", z[a:b])
    body += _chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + body)


def _png_case(name: str, w: int, h: int, seed: int):
    """-> (samples as stored, color type, depth, interlaced, palette, the array libspng would decode to)"""
    rgb8 = synth_image(w, h, 8, seed=seed)
    rgb16 = synth_image(w, h, 16, seed=seed)
    rng = np.random.default_rng(seed)
    inter = name.endswith("_i")
    base = name[:-2] if inter else name
    if base == "rgb8":
        return rgb8, 2, 8, inter, None, rgb8
    if base == "rgba8":
        a = rng.integers(0, 256, (h, w, 1), dtype=np.uint8)
        return np.concatenate([rgb8, a], 2), 6, 8, inter, None, rgb8
    if base == "grey8":
        g = rgb8[:, :, 1:2]
        return g, 0, 8, inter, None, np.repeat(g, 3, 2)
    if base == "ga8":
        g = rgb8[:, :, 1:2]
        return np.concatenate([g, 255 - g], 2), 4, 8, inter, None, np.repeat(g, 3, 2)
    if base in ("grey1", "grey2", "grey4"):
        d = int(base[-1])
        g = (rgb8[:, :, 1:2] >> (8 - d)).astype(np.uint8)
        return g, 0, d, inter, None, np.repeat(g * (255 // ((1 << d) - 1)), 3, 2).astype(np.uint8)
    if base in ("pal2", "pal4", "pal8"):
        d = int(base[-1])
        pal = rng.integers(0, 256, (1 << d, 3), dtype=np.uint8)
        idx = (rgb8[:, :, 0:1] >> (8 - d)).astype(np.uint8)
        return idx, 3, d, inter, pal, pal[idx[:, :, 0]]
    if base == "rgb16":
        return rgb16, 2, 16, inter, None, np.concatenate([rgb16, np.full((h, w, 1), 65535, np.uint16)], 2)
    if base == "rgba16":
        a = rng.integers(0, 65536, (h, w, 1), dtype=np.uint16)
        full = np.concatenate([rgb16, a], 2)
        return full, 6, 16, inter, None, full
    if base == "grey16":
        g = rgb16[:, :, 2:3]
        return g, 0, 16, inter, None, np.concatenate([g, g, g, np.full((h, w, 1), 65535, np.uint16)], 2)
    if base == "ga16":
        g = rgb16[:, :, 0:1]
        return np.concatenate([g, g ^ 0x5555], 2), 4, 16, inter, None, np.concatenate([g, g, g, g ^ 0x5555], 2)
    raise KeyError(name)


PNG_CASES = [
    # (kind, width, height, CLI options, shift)
    ("rgb8", 300, 260, ["--tile-size=0"], 0),
    ("rgb8", 300, 260, [], -1),
    ("rgba8", 70, 50, [], -1),
    ("grey8", 70, 50, ["--tile-size=0"], 0),
    ("ga8", 33, 40, [], -1),
    ("grey1", 67, 30, [], -1),
    ("grey2", 67, 30, [], -1),
    ("grey4", 67, 30, ["--one-frame"], -1),
    ("pal2", 45, 31, [], -1),
    ("pal4", 45, 31, [], -1),
    ("pal8", 280, 20, ["--tile-size=0"], 0),
    ("rgb16", 300, 260, ["--tile-size=0", "--linear"], 0),
    ("rgba16", 90, 70, ["--tile-size=1"], 1),
    ("grey16", 40, 40, [], -1),
    ("ga16", 40, 40, ["--linear"], -1),
    ("rgb8_i", 300, 260, ["--tile-size=0"], 0),
    ("rgb16_i", 41, 37, [], -1),
    ("grey2_i", 29, 23, [], -1),
    ("pal4_i", 9, 3, [], -1),
    ("rgb8", 520, 300, ["--tile-size=1"], 1),
]


@pytest.fixture(scope="module")
def cli_ref(reflib):
    """hydrium_cli.c + png_reader.c linked against the unmodified reference library."""
    refdir = os.path.join(ROOT, "oracle", "_ref")
    os.makedirs(os.path.dirname(CLI_REF), exist_ok=True)
    subprocess.run(["gcc", "-std=c99", "-O2", "-I", os.path.join(ROOT, "include"), "-o", CLI_REF, *CLI_SRC,
                    "-L", refdir, "-l:libhydrium_ref_O3.so", "-lz", f"-Wl,-rpath,{refdir}"], check=True)
    return CLI_REF


def _run(cli: str, args: list[str], src: str, dst: str, expect_ok: bool = True) -> bytes:
    r = subprocess.run([cli, *args, src, dst], capture_output=True, timeout=300)
    if expect_ok:
        assert r.returncode == 0, r.stderr.decode()
    with open(dst, "rb") as f:
        return f.read()


def _want_png(lib, decoded: np.ndarray, opts: list[str], shift: int) -> bytes:
    return encode_cli_loop(lib, decoded, linear_light=int("--linear" in opts), shift_x=shift, shift_y=shift,
                           pixel_stride=decoded.shape[2])


@pytest.mark.parametrize("case", PNG_CASES, ids=[f"{c[0]}-{c[1]}x{c[2]}{''.join(c[3])}" for c in PNG_CASES])
def test_png_reader_and_call_sequence_against_reference_library(cli_ref, reflib, tmp_path, case):
    kind, w, h, opts, shift = case
    samples, color, depth, inter, pal, decoded = _png_case(kind, w, h, seed=len(kind) + w)
    src = str(tmp_path / "in.png")
    write_png(src, samples, color, depth, inter, pal)
    got = _run(cli_ref, opts, src, str(tmp_path / "out.jxl"))
    assert got == _want_png(reflib, np.ascontiguousarray(decoded), opts, shift)


def _write_pfm(path: str, img: np.ndarray, little: bool) -> None:
    h, w, _ = img.shape
    with open(path, "wb") as f:
        f.write(f"PF\n{w} {h}\n{'-1.0' if little else '1.0'}\n".encode())
        f.write(img[::-1].astype("<f4" if little else ">f4").tobytes())


def _pfm_image(w: int, h: int, seed: int) -> np.ndarray:
    return (synth_image(w, h, 16, seed=seed).astype(np.float32) / np.float32(65535.0)).astype(np.float32)


PFM_CASES = [(300, 270, True, ["--tile-size=0"], 0), (300, 270, False, ["--tile-size=0", "--linear"], 0),
             (130, 90, True, [], -1), (600, 280, False, ["--tile-size=1"], 1)]


def _want_pfm(lib, img, opts, shift):
    from hydrium_b200.encoder import tile_grid
    _, _, ntx, nty = tile_grid(img.shape[1], img.shape[0], shift, shift)
    order = [(x, y) for y in range(nty - 1, -1, -1) for x in range(ntx)]   # bottom tile row first
    return encode_cli_loop(lib, img, linear_light=int("--linear" in opts), shift_x=shift, shift_y=shift, tiles=order,
                           is_last=lambda tx, ty: int(ty == 0 and tx == ntx - 1))


@pytest.mark.parametrize("case", PFM_CASES)
def test_pfm_reader_against_reference_library(cli_ref, reflib, tmp_path, case):
    w, h, little, opts, shift = case
    img = _pfm_image(w, h, seed=w)
    src = str(tmp_path / "in.pfm")
    _write_pfm(src, img, little)
    got = _run(cli_ref, opts, src, str(tmp_path / "out.jxl"))
    assert got == _want_pfm(reflib, img, opts, shift)


def test_cli_reads_stdin_and_writes_stdout(cli_ref, tmp_path):
    samples, color, depth, inter, pal, _ = _png_case("rgb8", 70, 50, seed=5)
    src = str(tmp_path / "in.png")
    write_png(src, samples, color, depth, inter, pal)
    want = _run(cli_ref, [], src, str(tmp_path / "out.jxl"))
    with open(src, "rb") as f:
        r = subprocess.run([cli_ref, "-", "-"], stdin=f, capture_output=True, timeout=120)
    assert r.returncode == 0, r.stderr.decode()
    assert r.stdout == want
    r = subprocess.run([cli_ref, "--", src], capture_output=True, timeout=120)   # output defaults to stdout
    assert r.returncode == 0 and r.stdout == want


def test_cli_option_errors(cli_ref, tmp_path):
    def rc(*args):
        return subprocess.run([cli_ref, *args], capture_output=True, timeout=60)
    assert rc().returncode == 1
    assert rc("--help").returncode == 0
    assert rc("--tile-size=4", "a", "b").returncode == 2
    assert rc("--one-frame", "--tile-size=1", "a", "b").returncode == 2
    assert rc("--tile-size=0", "--tag-icc-from=x.icc", "a", "b").returncode == 2
    assert rc("--bogus", "a", "b").returncode == 2
    assert rc("a", "b", "c").returncode == 2
    r = rc(str(tmp_path / "missing.png"), str(tmp_path / "o.jxl"))
    assert r.returncode == 1 and b"error opening file" in r.stderr
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"not a png at all")
    r = rc(str(bad), str(tmp_path / "o.jxl"))
    assert r.returncode == 1 and b"invalid signature" in r.stderr


# ---- the product binary: same files, same bytes as the reference-linked binary ----------------------
@pytest.fixture(scope="module", autouse=True)
def _product_cli_built():
    if not os.path.exists(CLI_OURS):
        subprocess.run(["make", "-C", os.path.join(ROOT, "hydrium_b200", "csrc")], check=True, capture_output=True)


GPU_CASES = [PNG_CASES[0], PNG_CASES[1], PNG_CASES[8], PNG_CASES[11], PNG_CASES[12], PNG_CASES[15], PNG_CASES[19],
             ("rgb8", 2100, 300, [], -1),            # one frame over two LF groups
             ("rgb16", 2060, 40, ["--linear"], -1)]  # ... 16-bit linear, the second LF group 12 px wide


@pytest.mark.gpu
@pytest.mark.parametrize("case", GPU_CASES, ids=[f"{c[0]}-{c[1]}x{c[2]}{''.join(c[3])}" for c in GPU_CASES])
def test_product_cli_writes_the_reference_cli_bytes_png(cli_ref, tmp_path, case):
    kind, w, h, opts, _ = case
    samples, color, depth, inter, pal, _ = _png_case(kind, w, h, seed=len(kind) + w)
    src = str(tmp_path / "in.png")
    write_png(src, samples, color, depth, inter, pal)
    assert _run(CLI_OURS, opts, src, str(tmp_path / "a.jxl")) == _run(cli_ref, opts, src, str(tmp_path / "b.jxl"))


@pytest.mark.gpu
@pytest.mark.parametrize("case", PFM_CASES)
def test_product_cli_writes_the_reference_cli_bytes_pfm(cli_ref, tmp_path, case):
    w, h, little, opts, _ = case
    src = str(tmp_path / "in.pfm")
    _write_pfm(src, _pfm_image(w, h, seed=w), little)
    assert _run(CLI_OURS, opts, src, str(tmp_path / "a.jxl")) == _run(cli_ref, opts, src, str(tmp_path / "b.jxl"))


@pytest.mark.gpu
def test_product_cli_icc_tagging(cli_ref, tmp_path):
    from test_gpu_parity import _fake_icc
    icc = tmp_path / "p.icc"
    icc.write_bytes(_fake_icc(3144, 11, b"MSFT"))
    samples, color, depth, inter, pal, _ = _png_case("rgb8", 300, 260, seed=3)
    src = str(tmp_path / "in.png")
    write_png(src, samples, color, depth, inter, pal)
    opts = [f"--tag-icc-from={icc}"]
    a = _run(CLI_OURS, opts, src, str(tmp_path / "a.jxl"))
    assert a == _run(cli_ref, opts, src, str(tmp_path / "b.jxl"))
    assert a != _run(CLI_OURS, [], src, str(tmp_path / "c.jxl"))
