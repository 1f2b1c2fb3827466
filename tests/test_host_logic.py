"""The device's sequential entropy logic (hydrium_b200/csrc/*.cuh), compiled for the host by
tests/host_harness, against the oracle: sparse prefix coder, constant sections, LF stream,
ANS model + chain formulation + packing, headers, exact reciprocal division.
These are the same source lines the CUDA kernels execute in one thread per tile."""
import ctypes as C

import numpy as np
import pytest

from oracle.pyoracle import Stages
from util import bits_of_bytes, bits_of_words, image_set


def _prefix(hh, vals, plain, lz, mod, c0, c1):
    out = np.zeros(1 << 16, np.uint32)
    bl = C.c_uint32(0)
    v = np.ascontiguousarray(vals, np.uint32)
    err = hh.hh_prefix_stream(v.ctypes.data_as(C.c_void_p), C.c_uint32(v.size), plain, lz, mod, (C.c_int * 3)(*c0),
                              (C.c_int * 3)(*c1), out.ctypes.data_as(C.c_void_p), out.size, C.byref(bl))
    return err, bits_of_words(out, bl.value)


def test_exact_reciprocal_division(host_harness):
    assert host_harness.hh_div_check(1, 4096, 3000) == 0


def test_sparse_prefix_coder_matches_oracle(host_harness, oracle):
    rng = np.random.default_rng(1)
    for trial in range(240):
        n = int(rng.integers(1, 3100))
        kind = trial % 8
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
        elif kind == 5:
            vals = rng.geometric(0.05, n) - 1
        elif kind == 6:   # Fibonacci-like weights: forces the depth limiter of the tree builder
            fib = [1, 1]
            while len(fib) < 22:
                fib.append(fib[-1] + fib[-2])
            vals = np.concatenate([np.full(min(f, 400), i * 3) for i, f in enumerate(fib)])[:3000]
            rng.shuffle(vals)
        else:
            vals = np.full(n, 77)
        ref_bytes, ref_bits = oracle.prefix_stream(vals, custom=(7, 1, 1), lz77_min_symbol=1 << 14, modular=1)
        err, bits = _prefix(host_harness, vals, 1, 1 << 14, 1, (7, 1, 1), (7, 1, 1))
        assert err == 0 and np.array_equal(bits, bits_of_bytes(ref_bytes, ref_bits)), (trial, kind)
    for vals, lz, mod, c0, c1, kw in [
        (np.r_[np.zeros(100), np.full(64, 8), np.zeros(64)], 29, 1, (4, 1, 1), (7, 0, 0), dict(custom=None)),
        (rng.integers(0, 9, 1485), 64, 0, (4, 1, 0), (4, 1, 0), dict(custom=(4, 1, 0))),
    ]:
        ref_bytes, ref_bits = oracle.prefix_stream(vals, lz77_min_symbol=lz, modular=mod, **kw)
        err, bits = _prefix(host_harness, vals, 1, lz, mod, c0, c1)
        assert err == 0 and np.array_equal(bits, bits_of_bytes(ref_bytes, ref_bits))


def _section(fn, *args):
    out = np.zeros(256, np.uint32)
    bl = C.c_uint32()
    err = fn(*args, out.ctypes.data_as(C.c_void_p), out.size, C.byref(bl))
    assert err == 0
    return bits_of_words(out, bl.value)


def _encode_like_device(hh, st: Stages):
    """Assemble a tile payload from the oracle's quantised ints / symbols the way k_lf_group + k_ans do."""
    vbw, vbh = st.vbw, st.vbh
    A = _section(hh.hh_section_a)
    B = _section(hh.hh_section_b, vbw, vbh)
    q = st.quant[:vbw * 8 * vbh * 8 * 3].reshape(vbh * 8, vbw * 8, 3)
    lfq = np.zeros((3, 32, 32), np.int32)
    lfq[:, :vbh, :vbw] = q[::8, ::8, :].transpose(2, 0, 1)
    out = np.zeros(8192, np.uint32)
    bl = C.c_uint32()
    assert hh.hh_lf_stream(lfq.ctypes.data_as(C.c_void_p), vbw, vbh, out.ctypes.data_as(C.c_void_p), out.size, C.byref(bl)) == 0
    L = bits_of_words(out, bl.value)
    s = st.hf_syms[:st.n_syms]
    packed = (s[:, 0] | (s[:, 1] << 8) | (s[:, 2] << 12) | (s[:, 3] << 16)).astype(np.uint32)
    fr = np.zeros((9, 64), np.uint32)
    al = np.zeros(9, np.uint32)
    d = np.zeros(1024, np.uint32)
    dl = C.c_uint32()
    e = np.zeros(1 << 18, np.uint32)
    el = C.c_uint32()
    err = hh.hh_ans_encode(packed.ctypes.data_as(C.c_void_p), packed.size, fr.ctypes.data_as(C.c_void_p),
                           al.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), d.size, C.byref(dl),
                           e.ctypes.data_as(C.c_void_p), e.size, C.byref(el))
    assert err == 0
    return A, L, B, bits_of_words(d, dl.value), bits_of_words(e, el.value), fr


@pytest.mark.parametrize("compact", [0, 1], ids=["table", "compact_pieces"])
def test_device_formulation_reproduces_oracle_payload(host_harness, oracle, compact):
    """compact = 1: the rANS chain runs over the sorted-piece form of the inverse alias map
    (k_ans_chain_compact); the harness also checks the pieces against the direct table slot by slot."""
    host_harness.hh_set_compact(compact)
    rng = np.random.default_rng(3)
    for name, img, lin in image_set(rng):
        h, w, _ = img.shape
        for ty in range((h + 255) // 256):
            for tx in range((w + 255) // 256):
                st = Stages()
                frame = oracle.encode_tile(img, tx, ty, linear_light=lin, stages=st)
                A, L, B, D, E, fr = _encode_like_device(host_harness, st)
                pre = bits_of_bytes(st.bits("pre"), st.pre_bitlen)
                assert np.array_equal(np.concatenate([A, L, B, D]), pre), (name, tx, ty, "prefix")
                assert np.array_equal(E, bits_of_bytes(st.bits("ans"), st.ans_bitlen)), (name, tx, ty, "E")
                assert np.array_equal(fr, st.freqs[:, :64]), (name, tx, ty, "freqs")
                payload = np.concatenate([pre, E])
                payload = np.packbits(np.concatenate([payload, np.zeros((-len(payload)) % 8, np.uint8)]), bitorder="little").tobytes()
                tw, th = min(256, w - tx * 256), min(256, h - ty * 256)
                last = (tx + 1) * 256 >= w and (ty + 1) * 256 >= h
                hdr = np.zeros(16, np.uint32)
                nb = host_harness.hh_frame_header(int(w > tw or h > th), tx * 256, ty * 256, tw, th, int(last), len(payload),
                                                  hdr.ctypes.data_as(C.c_void_p), 16)
                assert hdr.view(np.uint8)[:nb // 8].tobytes() + payload == frame, (name, tx, ty, "frame")


def test_image_header_matches_oracle(host_harness, oracle):
    for (w, h) in [(1, 1), (256, 256), (700, 600), (4096, 4096), (16384, 16384), (1920, 1080), (1 << 20, 9)]:
        out = np.zeros(16, np.uint32)
        n = host_harness.hh_image_header(w, h, out.ctypes.data_as(C.c_void_p), 16)
        got = out.view(np.uint8)[:n // 8].tobytes()
        want = oracle.image_header(w, h)
        assert got == want[-len(got):]


def test_staging_copy_pool(tmp_path):
    """hydrium_b200/csrc/stage_pool.c (the staging copy of hyd_send_tile spread over helper threads): 400 random
    jobs (interleaved and planar, negative pitches, 0..8 helpers), helpers that went to sleep in between, and
    four threads calling at once -- every byte in place, nothing written past the destination."""
    import os
    import subprocess
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(ROOT, "tests", "host_harness", "stage_pool_test.c")
    inc = os.path.join(ROOT, "hydrium_b200", "csrc")
    exe = str(tmp_path / "stage_pool_test")
    subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-I", inc, "-o", exe, src, os.path.join(inc, "stage_pool.c"), "-lpthread"],
                   check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and res.stdout.strip() == "0", res.stdout + res.stderr


def test_staging_copy_pool_under_thread_sanitizer(tmp_path):
    """The same driver built with -fsanitize=thread: the claim protocol (descriptor published word by word, rows
    claimed through the 64-bit ticket) must be free of data races, not only produce the right bytes."""
    import os
    import subprocess
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(ROOT, "tests", "host_harness", "stage_pool_test.c")
    inc = os.path.join(ROOT, "hydrium_b200", "csrc")
    exe = str(tmp_path / "stage_pool_tsan")
    b = subprocess.run(["gcc", "-std=c99", "-O1", "-g", "-fsanitize=thread", "-I", inc, "-o", exe, src,
                        os.path.join(inc, "stage_pool.c"), "-lpthread"], capture_output=True, text=True)
    if b.returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available: " + b.stderr[-200:])
    res = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    if "FATAL: ThreadSanitizer" in res.stderr:   # e.g. unsupported address-space layout in this container
        pytest.skip(res.stderr[-200:])
    assert res.returncode == 0 and "WARNING: ThreadSanitizer" not in res.stderr and res.stdout.strip().endswith("0"), \
        res.stdout[-500:] + res.stderr[-3000:]
