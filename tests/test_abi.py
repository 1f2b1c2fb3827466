"""The C-ABI shared library: loads, exports every symbol include/hydrium_b200.h declares, keeps the
reference's struct layout / enum values, and mirrors the reference's host-side error behaviour for
calls that need no GPU.  No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from hydrium_b200 import abi
from hydrium_b200.encoder import HYDEncoder

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(product_lib):
    hdr = open(os.path.join(ROOT, "include", "hydrium_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(hydb?_[a-z0-9_]+)\s*\(", hdr)))
    assert len([d for d in declared if d.startswith("hyd_")]) == 9
    for name in declared:
        assert hasattr(product_lib, name), name
    for name in abi.HYD_SYMBOLS:
        assert name in declared


def test_shim_header_and_struct_layout():
    assert os.path.exists(os.path.join(ROOT, "include", "libhydrium", "libhydrium.h"))
    assert C.sizeof(abi.HYDImageMetadata) == 32 and abi.HYDImageMetadata.linear_light.offset == 16
    assert (abi.HYD_OK, abi.HYD_NEED_MORE_OUTPUT, abi.HYD_ERROR_START, abi.HYD_NOMEM, abi.HYD_API_ERROR,
            abi.HYD_INTERNAL_ERROR) == (0, -2, -10, -13, -14, -15)


def _api_script(lib):
    """A call sequence exercising every host-side validation; returns [(ret, message), ...]."""
    res = []
    enc = HYDEncoder(lib)

    def rec(ret):
        res.append((ret, enc.error_message_get() if ret < abi.HYD_ERROR_START else None))

    rec(enc.set_metadata(0, 5))
    rec(enc.set_metadata((1 << 30) + 1, 5))
    rec(enc.set_metadata(1 << 30, 1 << 30))
    rec(enc.set_metadata(64, 64, 0, 4, 0))
    rec(enc.set_metadata(64, 64, 0, 0, -2))
    rec(enc.set_metadata(300, 200, 0, 0, 0))
    small = np.zeros(32, np.uint8)
    rec(enc.provide_output_buffer(small))
    rec(enc.release_output_buffer()[0])
    rec(enc.flush())
    buf = np.zeros(4096, np.uint8)
    rec(enc.provide_output_buffer(None, 4096))
    rec(enc.provide_output_buffer(buf))
    rec(enc.provide_output_buffer(buf))
    px = np.zeros((256, 256, 3), np.uint8)
    p = px.ctypes.data
    rec(enc.send_tile((p, p + 1, p + 2), 0, 0, 768, 3, -1, 7))        # invalid sample format
    rec(enc.send_tile((p, p + 1, p + 2), 2, 0, 768, 3, -1, 0))        # tile out of bounds
    rec(enc.send_tile((p, p + 1, p + 2), 0, 1, 768, 3, -1, 0))        # tile out of bounds (y)
    rec(enc.set_suggested_icc_profile(None))
    rec(enc.set_suggested_icc_profile(b"abcd"))                       # tile mode: refused
    ret, written = enc.release_output_buffer()
    rec(ret)
    res.append(("written", written))
    rec(enc.set_metadata(64, 64, 0, -1, -1))                          # one-frame mode: profiles are accepted
    rec(enc.set_suggested_icc_profile(b""))                           # non-null pointer, zero size
    rec(enc.set_suggested_icc_profile(bytes(range(200))))
    rec(enc.set_suggested_icc_profile(bytes(60)))                     # shorter than the 128-byte header
    rec(enc.set_suggested_icc_profile(None))
    enc.destroy()
    return res


def test_host_side_error_behaviour_matches_reference(product_lib, reflib):
    assert _api_script(product_lib) == _api_script(reflib)


def test_destroy_null_is_ok(product_lib):
    assert product_lib.hyd_encoder_destroy(None) == abi.HYD_OK


def test_unsupported_modes_fail_loudly(product_lib):
    """One-frame mode with more than 256 LF groups of 2048x2048 is the only geometry not built."""
    enc = HYDEncoder(product_lib)
    assert enc.set_metadata(2048 * 257, 8, 0, -1, -1) == abi.HYD_API_ERROR
    assert "one-frame mode is limited" in enc.error_message_get()
    assert enc.set_metadata(2048 * 16, 2048 * 16, 0, -1, -1) == abi.HYD_OK   # 256 LF groups
    assert enc.set_metadata(1024, 1024, 0, 1, 1) == abi.HYD_OK       # larger tiles: multi-group frames
    assert enc.set_metadata(1024, 1024, 0, 3, 0) == abi.HYD_OK
    assert enc.set_metadata(2048, 1100, 0, -1, -1) == abi.HYD_OK    # one LF group
    assert enc.set_metadata(256, 200, 0, -1, -1) == abi.HYD_OK      # one-frame that fits one group
    enc.destroy()


def test_image_header_entry_point(product_lib, oracle):
    for (w, h) in [(256, 256), (700, 600), (4096, 4096), (16384, 16384), (65536, 65536), (1 << 21, 3)]:
        buf = np.zeros(128, np.uint8)
        n = product_lib.hydb_image_header(w, h, buf.ctypes.data, 128)
        assert buf[:n].tobytes() == oracle.image_header(w, h)
