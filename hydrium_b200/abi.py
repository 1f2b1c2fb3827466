"""ctypes view of the libhydrium C ABI (reference: src/include/libhydrium/libhydrium.h:67-314).

The same binding is used for the reference library built under oracle/_ref (tests and the
CPU baseline only) and for our own libhydrium_b200.so, which exports the same nine symbols.
"""
from __future__ import annotations

import ctypes as C

HYD_OK = 0
HYD_DEFAULT = -1
HYD_NEED_MORE_OUTPUT = -2
HYD_NEED_MORE_INPUT = -3
HYD_ERROR_START = -10
HYD_NOMEM = -13
HYD_API_ERROR = -14
HYD_INTERNAL_ERROR = -15

HYD_UINT8 = 0
HYD_UINT16 = 1
HYD_FLOAT32 = 2


class HYDImageMetadata(C.Structure):
    _fields_ = [
        ("width", C.c_size_t),
        ("height", C.c_size_t),
        ("linear_light", C.c_int),
        ("tile_size_shift_x", C.c_int),
        ("tile_size_shift_y", C.c_int),
    ]


HYD_SYMBOLS = (
    "hyd_encoder_new", "hyd_encoder_destroy", "hyd_set_metadata", "hyd_provide_output_buffer",
    "hyd_send_tile", "hyd_release_output_buffer", "hyd_flush", "hyd_error_message_get",
    "hyd_set_suggested_icc_profile",
)


def bind_hyd_api(lib: C.CDLL) -> C.CDLL:
    """Attach argtypes/restypes for the nine hyd_* entry points."""
    vp = C.c_void_p
    lib.hyd_encoder_new.restype = vp
    lib.hyd_encoder_new.argtypes = []
    lib.hyd_encoder_destroy.restype = C.c_int
    lib.hyd_encoder_destroy.argtypes = [vp]
    lib.hyd_set_metadata.restype = C.c_int
    lib.hyd_set_metadata.argtypes = [vp, C.POINTER(HYDImageMetadata)]
    lib.hyd_provide_output_buffer.restype = C.c_int
    lib.hyd_provide_output_buffer.argtypes = [vp, vp, C.c_size_t]
    lib.hyd_send_tile.restype = C.c_int
    lib.hyd_send_tile.argtypes = [vp, C.POINTER(vp), C.c_uint32, C.c_uint32, C.c_ssize_t, C.c_ssize_t,
                                  C.c_int, C.c_int]
    lib.hyd_release_output_buffer.restype = C.c_int
    lib.hyd_release_output_buffer.argtypes = [vp, C.POINTER(C.c_size_t)]
    lib.hyd_flush.restype = C.c_int
    lib.hyd_flush.argtypes = [vp]
    lib.hyd_error_message_get.restype = C.c_char_p
    lib.hyd_error_message_get.argtypes = [vp]
    lib.hyd_set_suggested_icc_profile.restype = C.c_int
    lib.hyd_set_suggested_icc_profile.argtypes = [vp, vp, C.c_size_t]
    return lib
