"""Shared fixtures.  `-m "not gpu"` runs everything that needs no GPU (oracle vs golden vectors
and vs the reference build, the sequential device logic compiled for the host, ABI checks, the
gloo world_size-2 test); `-m gpu` are the parity tests proper, through the C ABI on a B200."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _make(path, *targets):
    subprocess.run(["make", "-C", path, *targets], check=True, stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    if not os.path.exists(os.path.join(ROOT, "oracle", "_build", "libhyd_oracle.so")):
        _make(os.path.join(ROOT, "oracle"), "oracle")
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def reflib():
    """The unmodified reference build; absent on machines without /root/reference and no prebuilt _ref."""
    from oracle import pyoracle
    if not pyoracle.have_ref():
        if os.path.isdir("/root/reference/src/libhydrium"):
            _make(os.path.join(ROOT, "oracle"), "ref")
        else:
            pytest.skip("oracle/_ref not available")
    return pyoracle.ref_library("O3")


@pytest.fixture(scope="session")
def reftap(reflib):
    from oracle import pyoracle
    return pyoracle.RefTap()


@pytest.fixture(scope="session")
def host_harness():
    """hydrium_b200/csrc/*.cuh compiled with g++ (tests/host_harness): the device's sequential
    entropy logic, runnable without a GPU.  A test tool, never a product path."""
    import ctypes
    src = os.path.join(ROOT, "tests", "host_harness", "harness.cpp")
    out_dir = os.path.join(ROOT, "tests", "host_harness", "_build")
    out = os.path.join(out_dir, "libhost_harness.so")
    inc = os.path.join(ROOT, "hydrium_b200", "csrc")
    deps = [src] + [os.path.join(inc, f) for f in os.listdir(inc) if f.endswith(".cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.run(["g++", "-x", "c++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-I", inc,
                        "-shared", "-o", out, src], check=True)
    lib = ctypes.CDLL(out)
    lib.hh_div_check.restype = ctypes.c_uint64
    return lib


@pytest.fixture(scope="session")
def product_lib():
    from hydrium_b200.lib import LIB_PATH, build_library, load_library
    if not os.path.exists(LIB_PATH):
        build_library()
    return load_library()


@pytest.fixture(scope="session")
def engine(product_lib):
    from hydrium_b200.engine import Engine
    eng = Engine(device=0, max_batch_tiles=64)
    yield eng
    eng.close()
