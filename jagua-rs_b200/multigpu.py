"""Multi-GPU plumbing of the query path: one process per GPU, the flattened layout is
replicated (rank 0 flattens, everyone else receives the blob), the query batch is sharded into contiguous
slices, and there is no collective on the query path. torch.distributed is only the transport
(NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of an n-query batch owned by `rank` (sizes differ by at most one)."""
    assert 0 <= rank < world and n >= 0
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_segments(seg_offsets, rank: int, world: int):
    """The part of a multi-item batch (segment k = queries seg_offsets[k]..seg_offsets[k+1]) that falls into
    `rank`'s slice: (lo, hi, segment indices, offsets relative to lo) - ready for detect_multi on the slice."""
    offs = np.asarray(seg_offsets, np.int64)
    lo, hi = shard_bounds(int(offs[-1]), rank, world)
    ids, rel = [], [0]
    for k in range(len(offs) - 1):
        a, b = max(int(offs[k]), lo), min(int(offs[k + 1]), hi)
        if b > a:
            ids.append(k)
            rel.append(rel[-1] + (b - a))
    return lo, hi, ids, np.asarray(rel, np.int64)


def broadcast_blob(blob: Optional[np.ndarray], dist, src: int = 0, device: Optional[str] = None):
    """Replicates a layout blob (uint8 array, only needed on `src`) to every rank. Returns a torch uint8 tensor
    on `device` (None = CPU) holding the blob on every rank."""
    import torch

    dev = torch.device(device) if device else torch.device("cpu")
    size = torch.zeros(1, dtype=torch.int64, device=dev)
    if dist.get_rank() == src:
        assert blob is not None
        size[0] = int(blob.nbytes)
    dist.broadcast(size, src)
    buf = torch.empty(int(size.item()), dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        buf.copy_(torch.from_numpy(np.ascontiguousarray(blob).view(np.uint8)))
    dist.broadcast(buf, src)
    return buf
