"""Closed-form synthetic RGB test images.

Every pixel is a pure function of (x, y, channel, seed), so the same image can be
produced here with numpy for the CPU oracle and by the CUDA generator kernel
(`hydb_synth_fill` in csrc/synth.cu) directly in HBM for the large configurations.
All arithmetic is uint32 with wrap-around.
"""
from __future__ import annotations

import numpy as np

_U32 = np.uint32


def _mix32(v: np.ndarray) -> np.ndarray:
    v = v.astype(_U32, copy=True)
    v ^= v >> _U32(16)
    v *= _U32(0x7FEB352D)
    v ^= v >> _U32(15)
    v *= _U32(0x846CA68B)
    v ^= v >> _U32(16)
    return v


def synth_image(width: int, height: int, bits: int = 8, seed: int = 0, smooth: bool = False,
                x0: int = 0, y0: int = 0, full_width: int | None = None,
                full_height: int | None = None) -> np.ndarray:
    """Return an interleaved RGB image, shape (height, width, 3), uint8 or uint16.

    `x0/y0/full_*` generate a window of a larger virtual image (used to check single
    tiles of the gigapixel configuration without materialising it).
    `smooth` divides the noise amplitude by 8 (the "smooth variant" of SURVEY.md 8d).
    """
    if bits not in (8, 16):
        raise ValueError("bits must be 8 or 16")
    fw = width if full_width is None else full_width
    fh = height if full_height is None else full_height
    maxv = 255 if bits == 8 else 65535
    with np.errstate(over="ignore"):
        x = (np.arange(width, dtype=np.uint64) + np.uint64(x0))
        y = (np.arange(height, dtype=np.uint64) + np.uint64(y0))
        bx = (x * np.uint64(maxv) // np.uint64(max(fw - 1, 1))).astype(np.int64)
        by = (y * np.uint64(maxv) // np.uint64(max(fh - 1, 1))).astype(np.int64)
        bxy = ((x[None, :] + y[:, None]) * np.uint64(maxv) // np.uint64(max(fw + fh - 2, 1))).astype(np.int64)
        base = np.empty((height, width, 3), dtype=np.int64)
        base[..., 0] = bx[None, :]
        base[..., 1] = by[:, None]
        base[..., 2] = bxy
        hx = (x.astype(_U32) * _U32(0x9E3779B1))[None, :]
        hy = _mix32(y.astype(_U32) + _U32(0x7F4A7C15))[:, None]
        out = np.empty((height, width, 3), dtype=np.uint8 if bits == 8 else np.uint16)
        for c in range(3):
            cc = _U32((c * 0x85EBCA6B) & 0xFFFFFFFF)
            h = _mix32(hx ^ hy ^ cc ^ _U32(seed & 0xFFFFFFFF))
            if bits == 8:
                n = ((h >> _U32(24)) & _U32(31)).astype(np.int64) - 16
            else:
                n = ((h >> _U32(16)) & _U32(0x1FFF)).astype(np.int64) - 4096
            if smooth:
                n = n // 8
            out[..., c] = np.clip(base[..., c] + n, 0, maxv)
    return out
