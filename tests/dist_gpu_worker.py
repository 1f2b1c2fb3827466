"""world_size-N worker for tests/test_dist.py::test_peer_memory_gather (NCCL, one GPU per rank): every
rank encodes its band of tile rows with the CUDA engine straight into its region of rank 0's buffer
(hydrium_b200.dist.PeerGather), rank 0 compacts and compares with the oracle's whole-image encode."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hydrium_b200 import engine as E  # noqa: E402
from hydrium_b200.dist import PeerGather, shard_range  # noqa: E402
from hydrium_b200.synth import synth_image  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    with E.Engine(device=rank, max_batch_tiles=64) as eng:
        for (w, h) in [(1100, 1500), (700, 300), (513, 1025)]:
            img = synth_image(w, h, 8, seed=9)
            rows = (h + 255) // 256
            r0, r1 = shard_range(rows, world, rank)
            pg = PeerGather(eng, E.output_bound(w, h))
            n = 0
            if r1 > r0:
                band = np.ascontiguousarray(img[r0 * 256:min(h, r1 * 256)])
                d_in = eng.upload(band)
                n = eng.encode_image_device(d_in, w, h, 3, tile_rows=(r0, r1), with_header=(rank == 0),
                                            d_out=pg.d_out, d_out_cap=pg.d_out_cap)
                eng.device_free(d_in)
            cap = E.output_bound(w, h)
            d_final = eng.device_alloc(cap) if rank == 0 else 0
            total = pg.finish(n, d_final, cap)
            if rank == 0:
                want = Oracle().encode_image(img)
                got = eng.download(d_final, total)
                assert got == want, f"{w}x{h}: peer-memory gather differs ({len(got)} vs {len(want)})"
                eng.device_free(d_final)
            pg.close()
    dist.barrier()
    if rank == 0:
        print("DIST_GPU_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
