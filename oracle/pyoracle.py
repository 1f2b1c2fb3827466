"""ctypes access to the parity oracle -- TEST INFRASTRUCTURE ONLY.

Two checkers live here:
  * `Oracle`     our plain-C restatement (oracle/hyd_oracle.c -> oracle/_build/libhyd_oracle.so)
  * `RefTap`     stage taps on the unmodified reference (oracle/ref_tap.c -> oracle/_ref/libhydrium_tap.so)
and `ref_library()` returns the unmodified reference build (oracle/_ref/libhydrium_ref*.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under hydrium_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_SYMS = 3 * 1024 * 64


def build(quiet: bool = True) -> None:
    """Compile the restatement and, when /root/reference is present, the reference builds."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _load(path: str) -> C.CDLL:
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run `make -C oracle` (needs /root/reference for _ref)")
    return C.CDLL(path)


def ref_library(opt: str = "O3") -> C.CDLL:
    """The unmodified reference library. `opt`: "Os" (project default flags) or "O3"
    (byte-identical output)."""
    name = "libhydrium_ref.so" if opt == "Os" else "libhydrium_ref_O3.so"
    return _load(os.path.join(HERE, "_ref", name))


def have_ref() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libhydrium_ref_O3.so"))


class _Meta(C.Structure):
    _fields_ = [("width", C.c_size_t), ("height", C.c_size_t), ("linear_light", C.c_int),
                ("shift_x", C.c_int), ("shift_y", C.c_int)]


class _OrcStages(C.Structure):
    _fields_ = [
        ("xyb", C.c_void_p), ("dct", C.c_void_p), ("quant", C.c_void_p), ("nonzeroes", C.c_void_p),
        ("hf_syms", C.c_void_p), ("hf_syms_cap", C.c_uint64), ("hf_syms_n", C.c_uint64),
        ("lf_bits", C.c_void_p), ("lf_bits_cap", C.c_uint64), ("lf_bitlen", C.c_uint64),
        ("freqs", C.c_void_p), ("alphabet_sizes", C.c_uint16 * 16), ("max_alphabet_size", C.c_uint32),
        ("ans_bits", C.c_void_p), ("ans_bits_cap", C.c_uint64), ("ans_bitlen", C.c_uint64),
        ("pre_bits", C.c_void_p), ("pre_bits_cap", C.c_uint64), ("pre_bitlen", C.c_uint64),
        ("vbw", C.c_uint32), ("vbh", C.c_uint32),
    ]


class _TapOut(C.Structure):
    _fields_ = [
        ("xyb", C.c_void_p), ("dct", C.c_void_p), ("quant", C.c_void_p), ("nonzeroes", C.c_void_p),
        ("hf_syms", C.c_void_p), ("hf_syms_cap", C.c_uint64), ("hf_syms_n", C.c_uint64),
        ("lf_bits", C.c_void_p), ("lf_bits_cap", C.c_uint64), ("lf_bitlen", C.c_uint64),
        ("freqs", C.c_void_p), ("alphabet_sizes", C.c_uint16 * 16), ("max_alphabet_size", C.c_uint32),
        ("ans_bits", C.c_void_p), ("ans_bits_cap", C.c_uint64), ("ans_bitlen", C.c_uint64),
        ("pre_bits", C.c_void_p), ("pre_bits_cap", C.c_uint64), ("pre_bitlen", C.c_uint64),
        ("vbw", C.c_uint32), ("vbh", C.c_uint32),
    ]


class _OrcTile(C.Structure):
    _fields_ = [
        ("image_width", C.c_uint64), ("image_height", C.c_uint64), ("linear_light", C.c_int),
        ("tile_x", C.c_uint32), ("tile_y", C.c_uint32), ("is_last", C.c_int), ("sample_fmt", C.c_int),
        ("plane", C.c_void_p * 3), ("row_stride", C.c_ssize_t), ("pixel_stride", C.c_ssize_t),
    ]


class Stages:
    """Numpy-backed stage buffers shared by both checkers (same field meaning in both)."""

    def __init__(self):
        n = 256 * 256 * 3
        self.xyb = np.zeros(n, np.float32)
        self.dct = np.zeros(n, np.float32)
        self.quant = np.zeros(n, np.int32)
        self.nonzeroes = np.zeros(1024 * 3, np.uint8)
        self.hf_syms = np.zeros((MAX_SYMS, 4), np.uint32)   # token, cluster, nbits, residue
        self.lf_bits = np.zeros(1 << 17, np.uint8)
        self.freqs = np.zeros((9, 256), np.uint32)
        self.ans_bits = np.zeros(1 << 20, np.uint8)
        self.pre_bits = np.zeros(1 << 17, np.uint8)
        self.n_syms = 0
        self.lf_bitlen = self.ans_bitlen = self.pre_bitlen = 0
        self.alphabet_sizes = [0] * 16
        self.max_alphabet_size = 0
        self.vbw = self.vbh = 0

    def _fill(self, st):
        st.xyb = self.xyb.ctypes.data
        st.dct = self.dct.ctypes.data
        st.quant = self.quant.ctypes.data
        st.nonzeroes = self.nonzeroes.ctypes.data
        st.hf_syms = self.hf_syms.ctypes.data
        st.hf_syms_cap = MAX_SYMS
        st.lf_bits = self.lf_bits.ctypes.data
        st.lf_bits_cap = self.lf_bits.nbytes
        st.freqs = self.freqs.ctypes.data
        st.ans_bits = self.ans_bits.ctypes.data
        st.ans_bits_cap = self.ans_bits.nbytes
        st.pre_bits = self.pre_bits.ctypes.data
        st.pre_bits_cap = self.pre_bits.nbytes
        return st

    def _read(self, st, tap_layout: bool):
        self.n_syms = int(st.hf_syms_n)
        self.lf_bitlen, self.ans_bitlen, self.pre_bitlen = int(st.lf_bitlen), int(st.ans_bitlen), int(st.pre_bitlen)
        self.alphabet_sizes = list(st.alphabet_sizes)
        self.max_alphabet_size = int(st.max_alphabet_size)
        self.vbw, self.vbh = int(st.vbw), int(st.vbh)
        if tap_layout:   # reference tap stores (token, cluster, nbits, residue) as well
            pass

    def bits(self, which: str) -> bytes:
        arr, n = {"lf": (self.lf_bits, self.lf_bitlen), "ans": (self.ans_bits, self.ans_bitlen),
                  "pre": (self.pre_bits, self.pre_bitlen)}[which]
        return arr[:(n + 7) // 8].tobytes()


def _sample_fmt(image: np.ndarray) -> int:
    """HYDSampleFormat of a numpy image: uint8 -> 0, uint16 -> 1, float32 -> 2 (libhydrium.h:101-107)."""
    try:
        return {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.float32): 2}[image.dtype]
    except KeyError:
        raise ValueError(f"unsupported sample type {image.dtype}") from None


def _tile_args(image: np.ndarray, tx: int, ty: int, pixel_stride=None):
    h, w, ch = image.shape
    item = image.dtype.itemsize
    p = image.ctypes.data + (ty * 256 * w * ch + tx * 256 * ch) * item
    return (p, p + item, p + 2 * item), w * ch, (ch if pixel_stride is None else pixel_stride)


class Oracle:
    def __init__(self):
        self.lib = _load(os.path.join(HERE, "_build", "libhyd_oracle.so"))
        L = self.lib
        L.orc_encode_tile.restype = C.c_int64
        L.orc_encode_tile.argtypes = [C.POINTER(_OrcTile), C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_encode_image.restype = C.c_int64
        L.orc_encode_image.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_uint64]
        L.orc_image_header.restype = C.c_int64
        L.orc_image_header.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64]
        L.orc_build_luts.restype = None
        L.orc_build_luts.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_prefix_stream.restype = C.c_int64
        L.orc_prefix_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_void_p,
                                        C.c_uint64, C.POINTER(C.c_uint64)]
        L.orc_last_error.restype = C.c_char_p

    def error(self):
        m = self.lib.orc_last_error()
        return m.decode() if m else None

    def image_header(self, width: int, height: int) -> bytes:
        buf = np.zeros(128, np.uint8)
        n = self.lib.orc_image_header(width, height, buf.ctypes.data, buf.nbytes)
        if n < 0:
            raise RuntimeError(f"oracle error {n}: {self.error()}")
        return buf[:n].tobytes()

    def encode_tile(self, image: np.ndarray, tx: int, ty: int, *, linear_light=0, is_last=-1,
                    stages: Stages | None = None, image_size=None, window=False) -> bytes:
        """Frame bytes for tile (tx, ty) of `image` (H, W, C interleaved, uint8/uint16).
        With `window=True`, `image` is just that tile's pixels and `image_size` the real image."""
        image = np.ascontiguousarray(image)
        h, w, _ = image.shape
        planes, rs, ps = _tile_args(image, 0 if window else tx, 0 if window else ty)
        t = _OrcTile()
        t.image_width, t.image_height = (w, h) if image_size is None else image_size
        t.linear_light = linear_light
        t.tile_x, t.tile_y, t.is_last = tx, ty, is_last
        t.sample_fmt = _sample_fmt(image)
        t.plane = (C.c_void_p * 3)(*planes)
        t.row_stride, t.pixel_stride = rs, ps
        st = None
        if stages is not None:
            st = stages._fill(_OrcStages())
        out = np.zeros(1 << 20, np.uint8)
        n = self.lib.orc_encode_tile(C.byref(t), out.ctypes.data, out.nbytes, C.byref(st) if st else None)
        if n < 0:
            raise RuntimeError(f"oracle error {n}: {self.error()}")
        if stages is not None:
            stages._read(st, False)
        return out[:n].tobytes()

    def encode_image(self, image: np.ndarray, *, linear_light=0) -> bytes:
        image = np.ascontiguousarray(image)
        h, w, ch = image.shape
        cap = 64 + ((w + 255) // 256) * ((h + 255) // 256) * (1 << 20)
        out = np.zeros(min(cap, max(1 << 20, w * h * ch * 3 + (1 << 16))), np.uint8)
        n = self.lib.orc_encode_image(image.ctypes.data, w, h, ch, _sample_fmt(image),
                                      linear_light, out.ctypes.data, out.nbytes)
        if n < 0:
            raise RuntimeError(f"oracle error {n}: {self.error()}")
        return out[:n].tobytes()

    def luts(self, sample_fmt: int, linear_light: int):
        inp = np.zeros(256 if sample_fmt == 0 else 65536, np.uint16)
        bias = np.zeros(65536, np.float32)
        self.lib.orc_build_luts(sample_fmt, linear_light, inp.ctypes.data, bias.ctypes.data)
        return inp, bias

    def prefix_stream(self, values, ctx=None, cluster_map=None, num_dists=1, custom=None,
                      lz77_min_symbol=0, modular=0):
        """Bits of a prefix-coded stream; returns (bytes, bitlen)."""
        v = np.ascontiguousarray(values, np.uint32)
        cx = None if ctx is None else np.ascontiguousarray(ctx, np.uint32)
        cm = None if cluster_map is None else np.ascontiguousarray(cluster_map, np.uint8)
        out = np.zeros(max(1 << 16, v.size * 8), np.uint8)
        bl = C.c_uint64(0)
        s, m, l = custom if custom else (0, 0, 0)
        n = self.lib.orc_prefix_stream(v.ctypes.data, cx.ctypes.data if cx is not None else None, v.size,
                                       cm.ctypes.data if cm is not None else None, num_dists,
                                       1 if custom else 0, s, m, l, lz77_min_symbol, modular,
                                       out.ctypes.data, out.nbytes, C.byref(bl))
        if n < 0:
            raise RuntimeError(f"oracle error {n}: {self.error()}")
        return out[:n].tobytes(), bl.value


class RefTap:
    def __init__(self):
        self.lib = _load(os.path.join(HERE, "_ref", "libhydrium_tap.so"))
        L = self.lib
        L.hyd_tap_encode_tile.restype = C.c_int
        L.hyd_tap_encode_tile.argtypes = [C.POINTER(_Meta), C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32,
                                          C.c_ssize_t, C.c_ssize_t, C.c_int, C.c_int, C.POINTER(_TapOut),
                                          C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.hyd_tap_cosine_lut.argtypes = [C.c_void_p]
        L.hyd_tap_luts.restype = C.c_int
        L.hyd_tap_luts.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.hyd_tap_prefix_stream.restype = C.c_int64
        L.hyd_tap_prefix_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_int,
                                            C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_uint64]

    def prefix_stream(self, values, ctx=None, cluster_map=None, num_dists=1, custom=None,
                      lz77_min_symbol=0, modular=0):
        """Same contract as Oracle.prefix_stream, but through the reference's own entropy.c."""
        v = np.ascontiguousarray(values, np.uint32)
        cx = None if ctx is None else np.ascontiguousarray(ctx, np.uint32)
        cm = None if cluster_map is None else np.ascontiguousarray(cluster_map, np.uint8)
        out = np.zeros(max(1 << 16, v.size * 8), np.uint8)
        s, m, l = custom if custom else (0, 0, 0)
        bits = self.lib.hyd_tap_prefix_stream(v.ctypes.data, cx.ctypes.data if cx is not None else None, v.size,
                                              cm.ctypes.data if cm is not None else None, num_dists,
                                              1 if custom else 0, s, m, l, lz77_min_symbol, modular,
                                              out.ctypes.data, out.nbytes)
        if bits < 0:
            raise RuntimeError(f"reference error {bits}")
        return out[:(bits + 7) // 8].tobytes(), bits

    def encode_tile(self, image: np.ndarray, tx: int, ty: int, *, linear_light=0, is_last=-1,
                    stages: Stages | None = None, shift=0) -> bytes:
        """Bytes the reference emits for this single send_tile on a fresh encoder
        (image header + frame)."""
        image = np.ascontiguousarray(image)
        h, w, _ = image.shape
        planes, rs, ps = _tile_args(image, tx, ty)
        md = _Meta(w, h, linear_light, shift, shift)
        st = (stages or Stages())._fill(_TapOut())
        out = np.zeros(1 << 20, np.uint8)
        n = C.c_uint64(0)
        arr = (C.c_void_p * 3)(*planes)
        ret = self.lib.hyd_tap_encode_tile(C.byref(md), arr, tx, ty, rs, ps, is_last,
                                           _sample_fmt(image), C.byref(st),
                                           out.ctypes.data, out.nbytes, C.byref(n))
        if ret < -10:
            raise RuntimeError(f"reference error {ret}")
        if stages is not None:
            stages._read(st, True)
        return out[:n.value].tobytes()

    def cosine_lut(self) -> np.ndarray:
        a = np.zeros(56, np.float32)
        self.lib.hyd_tap_cosine_lut(a.ctypes.data)
        return a.reshape(7, 8)

    def luts(self, sample_fmt: int, linear_light: int):
        inp = np.zeros(256 if sample_fmt == 0 else 65536, np.uint16)
        bias = np.zeros(65536, np.float32)
        ret = self.lib.hyd_tap_luts(sample_fmt, linear_light, inp.ctypes.data, bias.ctypes.data)
        if ret < -10:
            raise RuntimeError(f"reference error {ret}")
        return inp, bias
