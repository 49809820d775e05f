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
