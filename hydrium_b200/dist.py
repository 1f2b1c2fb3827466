"""Multi-GPU plumbing: one process per GPU, tiles sharded with no data-path collective, and a
single gather of the per-rank codestream spans onto rank 0.

Tile-mode frames depend only on (image size, tile position, is_last, the tile's pixels), so rank r
encodes a contiguous range of tile rows and its output is one contiguous span of the final stream.
The only exchange is: all-gather of the span lengths, then a gather-v of the spans (NCCL send/recv
over NVLink on GPUs; gloo on CPU tensors in the unit tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_units: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced [begin, end) of `n_units` (tile rows, images ...) for `rank`."""
    base, extra = divmod(n_units, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_spans(local: torch.Tensor, dst: int = 0, group=None) -> tuple[torch.Tensor | None, list[int]]:
    """Concatenate every rank's 1-D uint8 span, in rank order, on rank `dst`.

    Returns (stream on dst / None elsewhere, list of span lengths).  One all_gather of 8-byte
    lengths plus one grouped send/recv round: the "single NCCL gather" of the codestream.
    """
    if local.dtype != torch.uint8 or local.dim() != 1:
        raise ValueError("span must be a 1-D uint8 tensor")
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_local = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    lens_t = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(lens_t, n_local, group=group)
    lens = [int(t.item()) for t in lens_t]
    if rank == dst:
        out = torch.empty(sum(lens), dtype=torch.uint8, device=local.device)
        offs = [0]
        for n in lens:
            offs.append(offs[-1] + n)
        out[offs[rank]:offs[rank + 1]] = local
        ops = [dist.P2POp(dist.irecv, out[offs[r]:offs[r + 1]], r, group) for r in range(world)
               if r != dst and lens[r]]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return out, lens
    if local.numel():
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, local.contiguous(), dst, group)]):
            req.wait()
    return None, lens
