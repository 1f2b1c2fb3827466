"""Loader for libhydrium_b200.so (built in-tree by hydrium_b200/csrc/Makefile).

There is no fallback: if the CUDA library is missing or does not load, importing code gets a
loud error.  Nothing here ever imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from . import abi

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libhydrium_b200.so")
_lib = None

HYDB_SYMBOLS = (
    "hydb_encoder_set_batch", "hydb_encoder_set_device", "hydb_engine_create", "hydb_engine_destroy",
    "hydb_engine_error", "hydb_engine_max_batch", "hydb_engine_stream", "hydb_engine_launch_count",
    "hydb_engine_graph_launch_count", "hydb_encoder_stats",
    "hydb_engine_encode_tiles", "hydb_engine_finish", "hydb_encode_image_device", "hydb_encode_image_host",
    "hydb_image_header", "hydb_host_alloc", "hydb_host_free", "hydb_device_alloc", "hydb_device_free",
    "hydb_memcpy_h2d", "hydb_memcpy_d2h", "hydb_device_count", "hydb_synth_fill", "hydb_engine_enable_taps",
    "hydb_engine_read_tap", "hydb_engine_enable_timing", "hydb_engine_stage_ms", "hydb_engine_frame_lengths",
    "hydb_engine_encode_frames", "hydb_engine_read_model", "hydb_oneframe_finish", "hydb_engine_icc_header", "hydb_ipc_export", "hydb_ipc_open", "hydb_ipc_close",
    "hydb_engine_compact_regions", "hydb_engine_store_u64", "hydb_engine_set_chain_kernel",
    "hydb_engine_submit_frames", "hydb_engine_job_poll", "hydb_engine_job_regather", "hydb_engine_job_release",
    "hydb_engine_slot_frame_lengths",
)


class HydbTile(C.Structure):
    _fields_ = [
        ("plane", C.c_void_p * 3), ("row_stride", C.c_int64), ("pixel_stride", C.c_int64),
        ("width", C.c_uint32), ("height", C.c_uint32), ("x0", C.c_uint32), ("y0", C.c_uint32),
        ("image_width", C.c_uint32), ("image_height", C.c_uint32), ("is_last", C.c_int32),
        ("sample_fmt", C.c_int32), ("linear_light", C.c_int32), ("with_image_header", C.c_int32),
    ]


def build_library(verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", os.path.join(_PKG, "csrc"), "-j8"], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C hydrium_b200/csrc` "
                           "(hydrium_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    abi.bind_hyd_api(lib)
    vp, u32, u64, i64 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int64
    lib.hydb_encoder_set_batch.restype = C.c_int
    lib.hydb_encoder_set_batch.argtypes = [vp, u32]
    lib.hydb_encoder_set_device.restype = C.c_int
    lib.hydb_encoder_set_device.argtypes = [vp, C.c_int]
    lib.hydb_engine_create.restype = C.c_int
    lib.hydb_engine_create.argtypes = [C.POINTER(vp), C.c_int, u32]
    lib.hydb_engine_destroy.restype = None
    lib.hydb_engine_destroy.argtypes = [vp]
    lib.hydb_engine_error.restype = C.c_char_p
    lib.hydb_engine_error.argtypes = [vp]
    lib.hydb_engine_max_batch.restype = u32
    lib.hydb_engine_max_batch.argtypes = [vp]
    lib.hydb_engine_stream.restype = u64
    lib.hydb_engine_stream.argtypes = [vp]
    lib.hydb_engine_launch_count.restype = u64
    lib.hydb_engine_launch_count.argtypes = [vp]
    lib.hydb_engine_graph_launch_count.restype = u64
    lib.hydb_engine_graph_launch_count.argtypes = [vp]
    lib.hydb_encoder_stats.restype = None
    lib.hydb_encoder_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    lib.hydb_engine_encode_tiles.restype = C.c_int
    lib.hydb_engine_encode_tiles.argtypes = [vp, C.POINTER(HydbTile), u32, vp, u64, u64]
    lib.hydb_engine_finish.restype = C.c_int
    lib.hydb_engine_finish.argtypes = [vp, C.POINTER(u64)]
    lib.hydb_encode_image_device.restype = C.c_int
    lib.hydb_encode_image_device.argtypes = [vp, vp, u32, u32, u32, i64, C.c_int, C.c_int, u32, u32, C.c_int,
                                             vp, u64, C.POINTER(u64)]
    lib.hydb_encode_image_host.restype = C.c_int
    lib.hydb_encode_image_host.argtypes = [vp, vp, u32, u32, u32, C.c_int, C.c_int, vp, u64, C.POINTER(u64)]
    lib.hydb_image_header.restype = i64
    lib.hydb_image_header.argtypes = [u32, u32, vp, u64]
    lib.hydb_host_alloc.restype = vp
    lib.hydb_host_alloc.argtypes = [C.c_size_t]
    lib.hydb_host_free.restype = None
    lib.hydb_host_free.argtypes = [vp]
    lib.hydb_device_alloc.restype = vp
    lib.hydb_device_alloc.argtypes = [C.c_size_t]
    lib.hydb_device_free.restype = None
    lib.hydb_device_free.argtypes = [vp]
    lib.hydb_ipc_export.restype = C.c_int
    lib.hydb_ipc_export.argtypes = [vp, vp]
    lib.hydb_ipc_open.restype = vp
    lib.hydb_ipc_open.argtypes = [vp]
    lib.hydb_ipc_close.restype = None
    lib.hydb_ipc_close.argtypes = [vp]
    lib.hydb_engine_store_u64.restype = C.c_int
    lib.hydb_engine_store_u64.argtypes = [vp, vp, u64]
    lib.hydb_engine_compact_regions.restype = C.c_int
    lib.hydb_engine_compact_regions.argtypes = [vp, vp, u32, u64, vp, u64, C.POINTER(u64)]
    lib.hydb_memcpy_h2d.restype = C.c_int
    lib.hydb_memcpy_h2d.argtypes = [vp, vp, C.c_size_t]
    lib.hydb_memcpy_d2h.restype = C.c_int
    lib.hydb_memcpy_d2h.argtypes = [vp, vp, C.c_size_t]
    lib.hydb_device_count.restype = C.c_int
    lib.hydb_device_count.argtypes = []
    lib.hydb_synth_fill.restype = C.c_int
    lib.hydb_synth_fill.argtypes = [vp, vp, u32, u32, u32, u32, u32, u32, C.c_int, u32, C.c_int]
    lib.hydb_engine_enable_taps.restype = C.c_int
    lib.hydb_engine_enable_taps.argtypes = [vp, C.c_int]
    lib.hydb_engine_read_tap.restype = i64
    lib.hydb_engine_read_tap.argtypes = [vp, C.c_int, u32, vp, u64]
    lib.hydb_engine_frame_lengths.restype = C.c_int
    lib.hydb_engine_frame_lengths.argtypes = [vp, vp, u32]
    lib.hydb_engine_enable_timing.restype = C.c_int
    lib.hydb_engine_enable_timing.argtypes = [vp, C.c_int]
    lib.hydb_engine_stage_ms.restype = C.c_int
    lib.hydb_engine_stage_ms.argtypes = [vp, C.POINTER(C.c_double * 7)]
    lib.hydb_engine_set_chain_kernel.restype = C.c_int
    lib.hydb_engine_set_chain_kernel.argtypes = [vp, C.c_int]
    _lib = lib
    return lib
