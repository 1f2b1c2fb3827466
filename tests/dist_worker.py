"""world_size-N worker for tests/test_dist.py (gloo, CPU).  Each rank produces the codestream span
of its contiguous band of tile rows -- with the ORACLE standing in for the GPU encoder, which is
what lets this run without a GPU -- and rank 0 gathers the spans with hydrium_b200.dist.gather_spans
and checks the concatenation against the single-process encode."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hydrium_b200.dist import gather_spans, shard_range  # noqa: E402
from hydrium_b200.synth import synth_image  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    orc = Oracle()
    for (w, h) in [(700, 1100), (300, 200), (513, 769)]:
        img = synth_image(w, h, 8, seed=9)
        rows = (h + 255) // 256
        cols = (w + 255) // 256
        r0, r1 = shard_range(rows, world, rank)
        span = bytearray()
        if rank == 0:
            span += orc.image_header(w, h)
        for ty in range(r0, r1):
            for tx in range(cols):
                span += orc.encode_tile(img, tx, ty)
        local = torch.frombuffer(bytes(span), dtype=torch.uint8).clone() if span else torch.empty(0, dtype=torch.uint8)
        out, lens = gather_spans(local, dst=0)
        assert sum(lens) >= 0 and len(lens) == world
        if rank == 0:
            want = orc.encode_image(img)
            got = out.numpy().tobytes()
            assert got == want, f"{w}x{h}: gathered stream differs ({len(got)} vs {len(want)})"
        else:
            assert out is None
    # shard_range covers everything exactly once, also when units < world
    for n in (0, 1, 2, 5, 16, 17):
        cover = []
        for r in range(world):
            b, e = shard_range(n, world, r)
            cover += list(range(b, e))
        assert cover == list(range(n))
    dist.barrier()
    if rank == 0:
        print("DIST_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
