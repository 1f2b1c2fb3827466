"""Batch encoder over device-resident images: thin Python view of the hydb_engine_* C ABI.

Device memory is whatever the caller owns (a torch CUDA tensor's data_ptr(), or hydb_device_alloc);
this module never touches torch itself.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .abi import HYD_ERROR_START, HYD_FLOAT32, HYD_OK, HYD_UINT8, HYD_UINT16
from .encoder import HydriumError
from .lib import HydbTile, load_library


def _fmt_of(image: np.ndarray) -> int:
    """HYDSampleFormat of a numpy image (reference: libhydrium.h:101-107)."""
    try:
        return {np.dtype(np.uint8): HYD_UINT8, np.dtype(np.uint16): HYD_UINT16,
                np.dtype(np.float32): HYD_FLOAT32}[image.dtype]
    except KeyError:
        raise ValueError(f"unsupported sample type {image.dtype}") from None

TAP_XYB, TAP_DCT, TAP_COEF, TAP_NZINFO, TAP_LFQ, TAP_SYMS, TAP_FREQS, TAP_LFBITS, TAP_SECT, TAP_PAYLOAD, \
    TAP_NSYMS, TAP_LFBITLEN, TAP_CLK = range(13)


class Engine:
    """One engine = one GPU + one workspace sized for `max_batch_tiles` tiles per launch."""

    def __init__(self, device: int = -1, max_batch_tiles: int = 256):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.hydb_engine_create(C.byref(h), device, max_batch_tiles)
        if rc < HYD_ERROR_START or not h:
            raise HydriumError(rc, "hydb_engine_create failed (no usable CUDA device? there is no CPU path)")
        self._h = h
        self.max_batch = max_batch_tiles

    def close(self):
        if getattr(self, "_h", None):
            self.lib.hydb_engine_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -----------------------------------------------------------------------
    def error(self) -> str:
        return (self.lib.hydb_engine_error(self._h) or b"").decode()

    def _check(self, rc: int) -> int:
        if rc != HYD_OK:
            raise HydriumError(rc, self.error())
        return rc

    @property
    def stream(self) -> int:
        return int(self.lib.hydb_engine_stream(self._h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.hydb_engine_launch_count(self._h))

    def device_alloc(self, nbytes: int) -> int:
        p = self.lib.hydb_device_alloc(nbytes)
        if not p:
            raise MemoryError(f"hydb_device_alloc({nbytes})")
        return p

    def device_free(self, ptr: int) -> None:
        self.lib.hydb_device_free(ptr)

    def upload(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr)
        p = self.device_alloc(max(arr.nbytes, 16))
        if self.lib.hydb_memcpy_h2d(p, arr.ctypes.data, arr.nbytes):
            raise RuntimeError("H2D copy failed")
        return p

    def download(self, ptr: int, nbytes: int) -> bytes:
        out = np.empty(nbytes, np.uint8)
        if nbytes and self.lib.hydb_memcpy_d2h(out.ctypes.data, ptr, nbytes):
            raise RuntimeError("D2H copy failed")
        return out.tobytes()

    CHAIN_AUTO, CHAIN_TABLE, CHAIN_COMPACT = 0, 1, 2

    def set_chain_kernel(self, mode: int) -> None:
        """Which rANS chain kernel runs (hydb_engine_set_chain_kernel): by launch size, the table kernel,
        or the compact-pieces kernel.  The bytes are the same."""
        self._check(self.lib.hydb_engine_set_chain_kernel(self._h, mode))

    # -- encoding ------------------------------------------------------------------------
    def encode_image_device(self, d_pixels: int, width: int, height: int, channels: int = 3, *,
                            row_stride: int | None = None, sample_fmt: int = HYD_UINT8, linear_light: int = 0,
                            tile_rows: tuple[int, int] | None = None, with_header: bool = True,
                            d_out: int, d_out_cap: int) -> int:
        """Encode (a band of tile rows of) an image resident in device memory into d_out.
        Returns the number of bytes written."""
        rs = width * channels if row_stride is None else row_stride
        r0, r1 = (0, (height + 255) // 256) if tile_rows is None else tile_rows
        n = C.c_uint64(0)
        self._check(self.lib.hydb_encode_image_device(self._h, d_pixels, width, height, channels, rs, sample_fmt,
                                                      linear_light, r0, r1, 1 if with_header else 0, d_out,
                                                      d_out_cap, C.byref(n)))
        return int(n.value)

    def encode_image(self, image: np.ndarray, *, linear_light: int = 0) -> bytes:
        """Convenience: numpy image (H, W, C) -> complete codestream bytes (tile mode, shift 0/0)."""
        image = np.ascontiguousarray(image)
        h, w, ch = image.shape
        fmt = _fmt_of(image)
        cap = output_bound(w, h)
        d_in = self.upload(image)
        d_out = self.device_alloc(cap)
        try:
            n = self.encode_image_device(d_in, w, h, ch, sample_fmt=fmt, linear_light=linear_light,
                                         d_out=d_out, d_out_cap=cap)
            return self.download(d_out, n)
        finally:
            self.device_free(d_in)
            self.device_free(d_out)

    def encode_image_host(self, image: np.ndarray, *, linear_light: int = 0, out: np.ndarray | None = None) -> bytes:
        """hydb_encode_image_host: host pixels in, host codestream out (H2D + D2H inside)."""
        image = np.ascontiguousarray(image)
        h, w, ch = image.shape
        fmt = _fmt_of(image)
        if out is None:
            out = np.empty(output_bound(w, h), np.uint8)
        n = C.c_uint64(0)
        self._check(self.lib.hydb_encode_image_host(self._h, image.ctypes.data, w, h, ch, fmt, linear_light,
                                                    out.ctypes.data, out.nbytes, C.byref(n)))
        return out[:n.value].tobytes()

    def encode_tiles(self, tiles: list[HydbTile], d_out: int, d_out_cap: int, pos: int = 0) -> int:
        arr = (HydbTile * len(tiles))(*tiles)
        self._check(self.lib.hydb_engine_encode_tiles(self._h, arr, len(tiles), d_out, d_out_cap, pos))
        n = C.c_uint64(0)
        self._check(self.lib.hydb_engine_finish(self._h, C.byref(n)))
        return int(n.value)

    def encode_image_batch(self, d_images: int, count: int, width: int, height: int, channels: int = 3, *,
                           sample_fmt: int = HYD_UINT8, linear_light: int = 0, d_out: int, d_out_cap: int):
        """`count` independent images of the same size, stored back to back in device memory, each
        its own codestream (BASELINE config 4: a batch of frames).  Tiles of different images share
        GPU launches.  Returns a list of (offset, length) into d_out, one complete codestream each
        (image header + frames)."""
        item = {HYD_UINT8: 1, HYD_UINT16: 2}.get(sample_fmt, 4)
        ntx, nty = (width + 255) // 256, (height + 255) // 256
        per_image = ntx * nty
        spans, pos = [], 0
        img_bytes = width * height * channels * item
        imgs_per_launch = max(1, self.max_batch // per_image)
        if per_image > self.max_batch:
            raise ValueError("image has more tiles than the engine batch; use encode_image_device per image")
        for first in range(0, count, imgs_per_launch):
            group = range(first, min(count, first + imgs_per_launch))
            tiles = []
            for k in group:
                base = d_images + k * img_bytes
                for ty in range(nty):
                    for tx in range(ntx):
                        t = HydbTile()
                        p = base + (ty * 256 * width + tx * 256) * channels * item
                        t.plane = (C.c_void_p * 3)(p, p + item, p + 2 * item)
                        t.row_stride, t.pixel_stride = width * channels, channels
                        t.x0, t.y0 = tx * 256, ty * 256
                        t.width, t.height = min(256, width - tx * 256), min(256, height - ty * 256)
                        t.image_width, t.image_height = width, height
                        t.is_last = int(tx == ntx - 1 and ty == nty - 1)
                        t.sample_fmt, t.linear_light = sample_fmt, linear_light
                        t.with_image_header = int(tx == 0 and ty == 0)
                        tiles.append(t)
            arr = (HydbTile * len(tiles))(*tiles)
            self._check(self.lib.hydb_engine_encode_tiles(self._h, arr, len(tiles), d_out, d_out_cap, pos))
            n = C.c_uint64(0)
            self._check(self.lib.hydb_engine_finish(self._h, C.byref(n)))
            lens = np.zeros(len(tiles), np.uint32)
            self._check(self.lib.hydb_engine_frame_lengths(self._h, lens.ctypes.data, len(tiles)))
            for gi, _ in enumerate(group):
                size = int(lens[gi * per_image:(gi + 1) * per_image].sum())
                spans.append((pos, size))
                pos += size
        return spans

    def synth_fill(self, d_dst: int, width: int, height: int, *, bits: int = 8, seed: int = 0, smooth: bool = False,
                   x0: int = 0, y0: int = 0, full_width: int | None = None, full_height: int | None = None) -> None:
        self._check(self.lib.hydb_synth_fill(self._h, d_dst, width, height, x0, y0,
                                             width if full_width is None else full_width,
                                             height if full_height is None else full_height, bits, seed,
                                             1 if smooth else 0))

    # -- per-kernel timing ---------------------------------------------------------------------
    def enable_timing(self, on: bool = True) -> None:
        self._check(self.lib.hydb_engine_enable_timing(self._h, 1 if on else 0))

    def stage_ms(self) -> dict:
        """Device ms accumulated since the last call (CUDA events on the launching streams)."""
        arr = (C.c_double * 7)()
        self._check(self.lib.hydb_engine_stage_ms(self._h, C.byref(arr)))
        keys = ("xyb_dct_quant", "hf_tokens", "ans_chain", "ans_pack", "gather", "lf_group", "batches")
        return dict(zip(keys, list(arr)))

    # -- stage taps (parity tests) -----------------------------------------------------------
    def enable_taps(self, on: bool = True) -> None:
        self._check(self.lib.hydb_engine_enable_taps(self._h, 1 if on else 0))

    def read_tap(self, what: int, tile: int, dtype, cap_bytes: int = 1 << 20) -> np.ndarray:
        buf = np.empty(cap_bytes, np.uint8)
        n = self.lib.hydb_engine_read_tap(self._h, what, tile, buf.ctypes.data, buf.nbytes)
        if n < 0:
            raise HydriumError(int(n), f"read_tap({what}) failed")
        return buf[:n].view(dtype).copy()


def output_bound(width: int, height: int) -> int:
    """A safe device output capacity for a whole image (generous: 3 bytes per pixel + slack)."""
    tiles = ((width + 255) // 256) * ((height + 255) // 256)
    return 128 + min(tiles * 768 * 1024, width * height * 4 + tiles * 4096)
