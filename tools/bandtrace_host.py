"""Band timeline of the host path (HYDRIUM_B200_BANDTRACE=1 python tools/bandtrace_host.py)."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydrium_b200 import engine as E
from hydrium_b200.synth import synth_image
W = H = 4096
with E.Engine(device=0, max_batch_tiles=256) as eng:
    lib = eng.lib
    n_in = W * H * 3
    cap = E.output_bound(W, H)
    h_in_p = lib.hydb_host_alloc(n_in); h_out_p = lib.hydb_host_alloc(cap)
    h_in = np.ctypeslib.as_array(C.cast(h_in_p, C.POINTER(C.c_uint8)), shape=(n_in,))
    h_in[:] = synth_image(W, H, 8).reshape(-1)
    n64 = C.c_uint64(0)
    import time
    for i in range(4):
        print("run", i, file=sys.stderr)
        t0 = time.perf_counter()
        rc = lib.hydb_encode_image_host(eng._h, h_in_p, W, H, 3, 0, 0, h_out_p, cap, C.byref(n64))
        print("rc", rc, "bytes", n64.value, "ms", 1e3 * (time.perf_counter() - t0), file=sys.stderr)
