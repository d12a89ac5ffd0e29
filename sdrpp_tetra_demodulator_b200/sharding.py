"""Multi-GPU layout of the path: independent channels, sharded by channel block, one process per GPU.

Nothing is exchanged while demodulating -- every channel is a private recurrence.  The only
collective is the epilogue BASELINE.json names: gather the decoded symbol streams.  Each rank ships its
dibits packed 4-per-byte ([C_local][stride/4] bytes, equal-sized rows) plus its int32 symbol counts; rank 0
receives rank-major, which is channel-major because shards are contiguous channel blocks.  The backend is
whatever the process group was built with: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Tuple


def channel_range(rank: int, world: int, n_channels: int) -> Tuple[int, int]:
    """[first, last) of the contiguous channel block owned by `rank`; blocks differ by at most one channel."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_channels, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def gather_decoded(packed, counts, dst: int = 0, group=None):
    """Gather equal-shaped per-rank tensors `packed` [C_local][W] uint8 and `counts` [C_local] int32 to `dst`.

    Returns (packed_all [world*C_local][W], counts_all [world*C_local]) on dst, (None, None) elsewhere.
    Shards must be equal-sized (pad the last one); that keeps it ONE collective of fixed size per tensor.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return packed, counts
    if rank == dst:
        p_all = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device=packed.device)
        c_all = torch.empty((world,) + tuple(counts.shape), dtype=counts.dtype, device=counts.device)
        dist.gather(packed, list(p_all.unbind(0)), dst=dst, group=group)
        dist.gather(counts, list(c_all.unbind(0)), dst=dst, group=group)
        return p_all.reshape(-1, packed.shape[-1]), c_all.reshape(-1)
    dist.gather(packed, None, dst=dst, group=group)
    dist.gather(counts, None, dst=dst, group=group)
    return None, None


def unpack_dibits(packed_row, count: int):
    """Inverse of tdm_pack_dibits for one channel (numpy)."""
    import numpy as np
    p = np.asarray(packed_row, np.uint8)
    d = np.stack([(p >> 6) & 3, (p >> 4) & 3, (p >> 2) & 3, p & 3], axis=1).reshape(-1)
    return d[:count].astype(np.uint8)
