"""Multi-GPU layout of the path: independent channels, sharded by channel block, one process per GPU.

Nothing is exchanged while demodulating -- every channel is a private recurrence.  The only
collective is the epilogue BASELINE.json names: gather the decoded symbol streams.  Each rank ships its
dibits packed 4-per-byte ([C_local][stride/4] bytes, equal-sized rows) plus its int32 symbol counts; rank 0
receives rank-major, which is channel-major because shards are contiguous channel blocks.  The backend is
whatever the process group was built with: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

from . import capi


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


class Communicator:
    """The C ABI's NCCL communicator (tdm_comm_*, include/tdm_b200.h), bootstrapped through an existing
    torch.distributed process group: rank 0 makes the 128-byte id, it is broadcast as a CPU/GPU tensor, every rank
    calls tdm_comm_create.  The gather itself (gather_packed) is then library code only: no torch collective."""

    def __init__(self, device: int, group=None):
        import torch
        import torch.distributed as dist
        self._lib = capi.lib()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = int(device)
        ident = (C.c_uint8 * 128)()
        if self.rank == 0:
            capi.check(self._lib.tdm_comm_unique_id(ident), "tdm_comm_unique_id")
        backend = dist.get_backend(group)
        t = torch.tensor(list(ident), dtype=torch.uint8, device=torch.device("cuda", device) if backend == "nccl" else "cpu")
        dist.broadcast(t, src=0, group=group)
        ident = (C.c_uint8 * 128)(*t.cpu().tolist())
        h = C.c_void_p()
        capi.check(self._lib.tdm_comm_create(ident, self.rank, self.world, self.device, C.byref(h)), "tdm_comm_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.tdm_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def gather_packed(self, packed, counts, dst: int = 0, out=None, stream=None):
        """packed [C_local][W] uint8 + counts [C_local] int32 (CUDA tensors) of every rank -> (packed_all
        [world*C_local][W], counts_all [world*C_local]) on dst, (None, None) elsewhere.  Asynchronous on `stream`
        (a torch.cuda.Stream; default: the current one)."""
        import torch
        st = stream if stream is not None else torch.cuda.current_stream(packed.device)
        p_all = c_all = None
        if self.rank == dst:
            if out is not None:
                p_all, c_all = out
            else:
                p_all = torch.empty((self.world * packed.shape[0], packed.shape[1]), dtype=torch.uint8, device=packed.device)
                c_all = torch.empty(self.world * counts.shape[0], dtype=torch.int32, device=counts.device)
        ptr = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        assert packed.is_contiguous() and counts.is_contiguous()
        capi.check(self._lib.tdm_gather_packed(self._h, dst, packed.shape[0], ptr(packed), packed.stride(0), ptr(counts), ptr(p_all), ptr(c_all),
                                               C.c_void_p(st.cuda_stream)), "tdm_gather_packed")
        return p_all, c_all
