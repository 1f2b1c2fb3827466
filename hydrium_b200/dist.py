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


def gather_spans(local: torch.Tensor, dst: int = 0, group=None, out: torch.Tensor | None = None
                 ) -> tuple[torch.Tensor | None, list[int]]:
    """Concatenate every rank's 1-D uint8 span, in rank order, on rank `dst`.

    Returns (stream on dst / None elsewhere, list of span lengths).  One all-gather of the 8-byte
    lengths (read back with a single synchronisation) plus one grouped send/recv round: the "single
    NCCL gather" of the codestream.  `out`, if given on `dst`, is a reusable receive buffer (the
    stream is returned as a view of it when it is large enough).
    """
    if local.dtype != torch.uint8 or local.dim() != 1:
        raise ValueError("span must be a 1-D uint8 tensor")
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_local = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    lens_t = torch.empty(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(lens_t, n_local, group=group)
    lens = [int(v) for v in lens_t.tolist()]
    if rank == dst:
        total = sum(lens)
        if out is not None and out.numel() >= total and out.dtype == torch.uint8 and out.device == local.device:
            out = out[:total]
        else:
            out = torch.empty(total, dtype=torch.uint8, device=local.device)
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


class PeerGather:
    """The gather of SURVEY 8e over peer memory instead of NCCL send/recv (GPUs of one node, one process
    each): rank `dst` owns a buffer of world x region_stride bytes, every rank opens it through CUDA
    IPC and lets its encoder's compaction kernel (k_gather_frames) write its span straight into its
    region over NVLink.  What is left per step is one tiny all-reduce as the barrier and, on `dst`,
    one kernel that closes the gaps between the spans (k_compact_regions).

        pg = PeerGather(engine, region_bytes)                    # collective: call on every rank
        n = engine.encode_image_device(..., d_out=pg.d_out, d_out_cap=pg.d_out_cap)
        total = pg.finish(n, d_final, d_final_cap)               # collective; bytes on dst, 0 elsewhere

    The regions are double-buffered: `d_out` alternates between two sets from one finish() to the next, so
    a fast rank's next step never writes into a region `dst` is still compacting (`dst` cannot pass the
    barrier of step k + 1 before its own compaction of step k has run: both are on its engine stream).
    """

    HEADER = 256

    def __init__(self, engine, region_bytes: int, dst: int = 0, group=None):
        import ctypes as C
        self.eng, self.lib, self.dst, self.group = engine, engine.lib, dst, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.stride = (self.HEADER + region_bytes + 16 + 255) & ~255
        self._own = self._peer = None
        self.parity = 0
        handle = [None]
        if self.rank == dst:
            self._own = engine.device_alloc(2 * self.world * self.stride)
            buf = (C.c_uint8 * 64)()
            if self.lib.hydb_ipc_export(self._own, buf) != 0:
                raise RuntimeError("cudaIpcGetMemHandle failed")
            handle[0] = bytes(buf)
        dist.broadcast_object_list(handle, src=dst, group=group)
        if self.rank == dst:
            self.base = self._own
        else:
            buf = (C.c_uint8 * 64).from_buffer_copy(handle[0])
            self._peer = self.lib.hydb_ipc_open(buf)
            if not self._peer:
                raise RuntimeError("cudaIpcOpenMemHandle failed (no peer access between the GPUs?)")
            self.base = self._peer
        self.d_out_cap = self.stride - self.HEADER - 16
        self._flag = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
        self._ext = torch.cuda.ExternalStream(engine.stream, device=self._flag.device)

    @property
    def set_base(self) -> int:
        """first byte of the region set this step writes into"""
        return self.base + self.parity * self.world * self.stride

    @property
    def d_out(self) -> int:
        """where this rank's encode call of the current step writes (its region of the current set)"""
        return self.set_base + self.rank * self.stride + self.HEADER

    def finish(self, n_local: int, d_final: int = 0, d_final_cap: int = 0) -> int:
        """After this rank's encode call has returned (its span is in the region): publish the length,
        wait for everyone, and on `dst` compact the spans into d_final.  Returns the stream's bytes on
        `dst`, 0 elsewhere."""
        import ctypes as C
        # the length goes out on the engine's stream and the barrier is enqueued behind it on the same
        # stream, so no rank passes the barrier before every span and every length has landed
        set_base = self.set_base
        self.eng._check(self.lib.hydb_engine_store_u64(self.eng._h, set_base + self.rank * self.stride, n_local))
        with torch.cuda.stream(self._ext):
            dist.all_reduce(self._flag, group=self.group)
        self.parity ^= 1
        if self.rank != self.dst:
            return 0
        total = C.c_uint64(0)
        self.eng._check(self.lib.hydb_engine_compact_regions(self.eng._h, set_base, self.world, self.stride, d_final,
                                                             d_final_cap, C.byref(total)))
        return int(total.value)

    def close(self) -> None:
        """Collective: peers unmap the buffer before its owner frees it."""
        if self._peer:
            self.lib.hydb_ipc_close(self._peer)
            self._peer = None
        dist.barrier(group=self.group)
        if self._own:
            self.eng.device_free(self._own)
            self._own = None
