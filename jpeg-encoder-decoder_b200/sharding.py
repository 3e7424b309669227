"""Batch-of-frames partitioning over GPUs.

Frames are independent encodes, so N ranks split a batch into contiguous ranges and never exchange pixel or
bitstream data; the only multi-rank step is bookkeeping: every rank publishes the sizes of the streams it produced and
everybody (or rank 0) derives the global (rank, offset, size) table that locates frame f's JFIF stream inside rank r's
compacted output buffer.  `torch.distributed` carries that table (NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range (first, count) of rank `rank`; the first n_frames % world ranks take one frame more."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_frames, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def owner_of(frame: int, n_frames: int, world: int) -> int:
    base, extra = divmod(n_frames, world)
    split = extra * (base + 1)
    return frame // (base + 1) if frame < split else extra + (frame - split) // max(base, 1)


def local_offsets(sizes: np.ndarray) -> np.ndarray:
    """Byte offset of every local stream inside the rank's compacted output (exclusive prefix of the sizes)."""
    sizes = np.asarray(sizes, dtype=np.int64)
    return np.concatenate([[0], np.cumsum(sizes)[:-1]]) if sizes.size else sizes


def gather_tables(sizes: np.ndarray, n_frames: int, rank: int, world: int, device=None):
    """All ranks call this with the sizes of their own streams.  Returns an (n_frames, 3) int64 array of
    (owner rank, byte offset inside the owner's compacted output, size) indexed by global frame number."""
    import torch
    import torch.distributed as dist
    sizes = np.asarray(sizes, dtype=np.int64)
    first, count = shard_range(n_frames, rank, world)
    if sizes.size != count:
        raise ValueError(f"rank {rank} owns {count} frames but reports {sizes.size} sizes")
    most = -(-n_frames // world)
    mine = torch.zeros(most, dtype=torch.int64, device=device)
    mine[:count] = torch.from_numpy(sizes).to(mine.device)
    if world > 1:
        parts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
    else:
        parts = [mine]
    table = np.zeros((n_frames, 3), dtype=np.int64)
    for r in range(world):
        f0, c = shard_range(n_frames, r, world)
        s = parts[r][:c].cpu().numpy()
        table[f0:f0 + c, 0] = r
        table[f0:f0 + c, 1] = local_offsets(s)
        table[f0:f0 + c, 2] = s
    return table
