"""Host-side mirror of the libhydrium encoder object over the C ABI.

`HYDEncoder` wraps the nine `hyd_*` entry points (reference: libhydrium.h:165-314) with the
same names, argument meaning and error behaviour; it works over any shared library that
exports them -- ours (`load_library()`), or, in tests / the CPU baseline only, the reference
build under oracle/_ref.  `encode_cli_loop` replays the call sequence of the reference CLI
(src/hydrium.c:275-286, 402-480), which is the compatibility target of the boundary.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .abi import (HYD_API_ERROR, HYD_ERROR_START, HYD_FLOAT32, HYD_NEED_MORE_OUTPUT, HYD_OK,
                  HYD_UINT8, HYD_UINT16, HYDImageMetadata)


class HydriumError(RuntimeError):
    def __init__(self, code: int, message: str | None):
        super().__init__(f"libhydrium error {code}: {message}")
        self.code = code
        self.message = message


_FMT_OF_DTYPE = {np.dtype(np.uint8): HYD_UINT8, np.dtype(np.uint16): HYD_UINT16,
                 np.dtype(np.float32): HYD_FLOAT32}


class HYDEncoder:
    """One encoder = one image = one thread at a time (reference: internal.h:34-83)."""

    def __init__(self, lib: C.CDLL):
        self._lib = abi.bind_hyd_api(lib)
        self._enc = self._lib.hyd_encoder_new()
        if not self._enc:
            raise MemoryError("hyd_encoder_new returned NULL")
        self._out = None  # keeps the lent output buffer alive

    # -- lifecycle -----------------------------------------------------------------
    def destroy(self) -> int:
        enc, self._enc = self._enc, None
        return self._lib.hyd_encoder_destroy(enc) if enc else HYD_OK

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.destroy()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # -- the reference API, verbatim -------------------------------------------------
    def set_metadata(self, width: int, height: int, linear_light: int = 0,
                     tile_size_shift_x: int = 0, tile_size_shift_y: int = 0) -> int:
        md = HYDImageMetadata(width, height, linear_light, tile_size_shift_x, tile_size_shift_y)
        return self._lib.hyd_set_metadata(self._enc, C.byref(md))

    def provide_output_buffer(self, buffer: np.ndarray | None, length: int | None = None) -> int:
        if buffer is None:
            return self._lib.hyd_provide_output_buffer(self._enc, None, length or 0)
        n = buffer.nbytes if length is None else length
        ret = self._lib.hyd_provide_output_buffer(self._enc, buffer.ctypes.data, n)
        if ret >= HYD_ERROR_START:
            self._out = buffer
        return ret

    def send_tile(self, planes, tile_x: int, tile_y: int, row_stride: int, pixel_stride: int,
                  is_last: int, sample_fmt: int) -> int:
        """`planes` = three raw addresses (ints) of the first R, G and B sample of the tile.
        Strides are in samples (libhydrium.h:219-220)."""
        arr = (C.c_void_p * 3)(*planes)
        return self._lib.hyd_send_tile(self._enc, arr, tile_x, tile_y, row_stride, pixel_stride,
                                       is_last, sample_fmt)

    def release_output_buffer(self) -> tuple[int, int]:
        written = C.c_size_t(0)
        ret = self._lib.hyd_release_output_buffer(self._enc, C.byref(written))
        self._out = None
        return ret, written.value

    def flush(self) -> int:
        return self._lib.hyd_flush(self._enc)

    def error_message_get(self) -> str | None:
        msg = self._lib.hyd_error_message_get(self._enc)
        return msg.decode() if msg else None

    def set_suggested_icc_profile(self, icc: bytes | None) -> int:
        if icc is None:
            return self._lib.hyd_set_suggested_icc_profile(self._enc, None, 0)
        buf = C.create_string_buffer(icc, len(icc))
        return self._lib.hyd_set_suggested_icc_profile(self._enc, C.addressof(buf), len(icc))

    # -- helpers ---------------------------------------------------------------------
    def check(self, ret: int) -> int:
        if ret < HYD_ERROR_START:
            raise HydriumError(ret, self.error_message_get())
        return ret


def tile_grid(width: int, height: int, shift_x: int, shift_y: int) -> tuple[int, int, int, int]:
    """(tile_w, tile_h, tiles_x, tiles_y) as hydrium.c:281-286 computes them
    (one-frame mode addresses 2048x2048 LF groups)."""
    sx = 3 if shift_x < 0 else shift_x
    sy = 3 if shift_y < 0 else shift_y
    tw, th = 256 << sx, 256 << sy
    return tw, th, (width + tw - 1) // tw, (height + th - 1) // th


def encode_cli_loop(lib: C.CDLL, image: np.ndarray, *, linear_light: int = 0, shift_x: int = 0,
                    shift_y: int = 0, out_buf_size: int = 1 << 20, pixel_stride: int | None = None,
                    tiles=None, is_last=-1, per_tile: list | None = None, icc: bytes | None = None,
                    batch: int | None = None, stats: dict | None = None) -> bytes:
    """Encode `image` (H, W, C>=3 interleaved; uint8/uint16/float32) exactly the way the reference
    CLI drives the library: one output buffer, and after every tile the
    flush / release / consume / provide loop (hydrium.c:402-480).

    `tiles` optionally restricts/reorders the (tile_x, tile_y) sequence (gaps are legal,
    libhydrium.h:240); `per_tile`, if a list, receives the bytes surfaced after each tile; `icc` is a
    suggested ICC profile (one-frame mode only, hydrium.c:288-296); `is_last` may be a function of
    (tile_x, tile_y), as the CLI's PFM path sets it explicitly (hydrium.c:459).
    `batch` (libhydrium_b200 only): hydb_encoder_set_batch -- 1 = every tile's bytes surface in the flush
    loop right after it, like the reference; default = the library's asynchronous pipeline.
    After the last tile the flush loop is run once more: a no-op for the reference, and for
    libhydrium_b200 the documented way to collect everything when no tile was marked last
    (tile subsets without is_last are legal).
    `stats` (libhydrium_b200 only): receives the engine's kernel-launch and graph-replay counters.
    """
    if image.ndim != 3 or image.shape[2] < 3:
        raise ValueError("image must be (H, W, C>=3)")
    image = np.ascontiguousarray(image)
    h, w, ch = image.shape
    fmt = _FMT_OF_DTYPE[image.dtype]
    item = image.dtype.itemsize
    pstride = ch if pixel_stride is None else pixel_stride
    row_stride = w * ch
    tw, th, ntx, nty = tile_grid(w, h, shift_x, shift_y)
    if tiles is None:
        tiles = [(x, y) for y in range(nty) for x in range(ntx)]
    out = bytearray()
    obuf = np.empty(out_buf_size, dtype=np.uint8)
    enc = HYDEncoder(lib)
    try:
        if batch is not None:
            lib.hydb_encoder_set_batch.restype = C.c_int
            lib.hydb_encoder_set_batch.argtypes = [C.c_void_p, C.c_uint32]
            enc.check(lib.hydb_encoder_set_batch(enc._enc, batch))
        enc.check(enc.set_metadata(w, h, linear_light, shift_x, shift_y))
        if icc is not None:
            enc.check(enc.set_suggested_icc_profile(icc))
        enc.check(enc.provide_output_buffer(obuf))
        base = image.ctypes.data
        for (tx, ty) in tiles:
            p = base + (ty * th * row_stride + tx * tw * ch) * item
            last = is_last(tx, ty) if callable(is_last) else is_last
            enc.check(enc.send_tile((p, p + item, p + 2 * item), tx, ty, row_stride, pstride,
                                    last, fmt))
            got = bytearray()
            while True:
                ret = enc.flush()
                r2, written = enc.release_output_buffer()
                enc.check(r2)
                got += obuf[:written].tobytes()
                enc.check(enc.provide_output_buffer(obuf))
                if ret != HYD_NEED_MORE_OUTPUT:
                    break
            enc.check(ret)
            out += got
            if per_tile is not None:
                per_tile.append(bytes(got))
        while True:   # final poll
            ret = enc.flush()
            r2, written = enc.release_output_buffer()
            enc.check(r2)
            out += obuf[:written].tobytes()
            if per_tile and written:
                per_tile[-1] += obuf[:written].tobytes()
            enc.check(enc.provide_output_buffer(obuf))
            if ret != HYD_NEED_MORE_OUTPUT:
                enc.check(ret)
                break
        if stats is not None and hasattr(lib, "hydb_encoder_stats"):   # libhydrium_b200 only
            k, g = C.c_uint64(0), C.c_uint64(0)
            lib.hydb_encoder_stats.restype = None
            lib.hydb_encoder_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
            lib.hydb_encoder_stats(enc._enc, C.byref(k), C.byref(g))
            stats["kernel_launches"], stats["graph_launches"] = k.value, g.value
    finally:
        enc.destroy()
    return bytes(out)
