"""Framebuffer partition across GPUs and the host-side final gather.

The hot path shards with no data-path collective: in keyed-RNG mode every pixel
has its own random stream, so rank r of N renders the rows y with y % N == r (round-robin rows
balance Cornell's spatially non-uniform cost) into its own buffer; afterwards rank 0 gathers
the disjoint rows on the host.  This module is that host logic; bench.py and the world_size-2
gloo test use it."""
from __future__ import annotations

import numpy as np


def rows_of_rank(height: int, rank: int, world: int) -> range:
    return range(rank, height, world)


def own_rows(frame: np.ndarray, rank: int, world: int) -> np.ndarray:
    """The rows of a full-frame array this rank owns, contiguous."""
    return np.ascontiguousarray(frame[rank::world])


def scatter_rows(frame: np.ndarray, rows: np.ndarray, rank: int, world: int) -> None:
    frame[rank::world] = rows


def gather_rows(local_frame: np.ndarray, rank: int, world: int, group=None, dst: int = 0):
    """Host-side gather of round-robin rows to `dst` over a CPU process group (gloo).
    local_frame: this rank's full-frame structured/plain array with its own rows filled.
    Returns the assembled frame on dst, None elsewhere."""
    if world == 1:
        return local_frame
    import torch
    import torch.distributed as dist

    mine = own_rows(local_frame, rank, world)
    item = mine.dtype.itemsize
    width = local_frame.shape[1]
    max_rows = len(rows_of_rank(local_frame.shape[0], 0, world))
    payload = np.zeros((max_rows, width), dtype=mine.dtype)  # ranks may own one row fewer
    payload[: mine.shape[0]] = mine
    send = torch.from_numpy(payload.view(np.uint8).reshape(-1))
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst, group=group)
    if rank != dst:
        return None
    frame = np.zeros_like(local_frame)
    for r, chunk in enumerate(recv):
        rows = chunk.numpy().view(mine.dtype).reshape(max_rows, width)
        n = len(rows_of_rank(local_frame.shape[0], r, world))
        scatter_rows(frame, rows[:n], r, world)
    assert item == frame.dtype.itemsize
    return frame


class SharedFrame:
    """The final gather without a copy: ONE framebuffer in host shared memory (/dev/shm) mapped by
    every rank of the box.  Each rank's row-partitioned render writes its own rows y = rank (mod N)
    straight into it (ptb200_render's D2H is a strided copy of exactly those rows), so after a
    barrier rank 0 holds the assembled frame — the host-side analogue of `output += pass`
    (src/dod/Scene.cpp:242) for disjoint rows.  No collective touches pixel data."""

    def __init__(self, height: int, width: int, dtype, tag: str, rank: int, barrier=None):
        import os
        self.path = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp", f"ptb200_frame_{tag}.bin")
        self.rank, self.barrier = rank, barrier or (lambda: None)
        nbytes = height * width * np.dtype(dtype).itemsize
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(nbytes)
        self.barrier()
        self.frame = np.memmap(self.path, dtype=dtype, mode="r+", shape=(height, width))

    def gathered(self):
        """Call on every rank after its render into `self.frame`; returns the frame on rank 0."""
        self.barrier()
        return self.frame if self.rank == 0 else None

    def close(self):
        import os
        self.barrier()
        del self.frame
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass
