"""N>1 path on CPU: two processes over gloo.  Each rank 'demodulates' its channel block (here with the oracle,
tests are allowed to), packs the dibits 4-per-byte exactly like tdm_pack_dibits, and the shards are gathered
with the same sharding.gather_decoded() the GPU bench uses over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _pack(d, counts, width):
    out = np.zeros((d.shape[0], width), np.uint8)
    for c in range(d.shape[0]):
        n = int(counts[c])
        v = np.zeros(width * 4, np.uint8)
        v[:n] = d[c, :n]
        v = v.reshape(-1, 4)
        out[c] = v[:, 0] << 6 | v[:, 1] << 4 | v[:, 2] << 2 | v[:, 3]
    return out


def _worker(rank, world, port, n_channels, n_samples, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from sdrpp_tetra_demodulator_b200.sharding import channel_range, gather_decoded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, last = channel_range(rank, world, n_channels)
    iq = O.generate(last - first, n_samples, first_channel=first)
    b = O.OracleB(last - first)
    counts, _, dibits, _ = b.process(iq, want_syms=False)
    width = (dibits.shape[1] + 3) // 4
    packed = torch.from_numpy(_pack(dibits, counts, width))
    p_all, c_all = gather_decoded(packed, torch.from_numpy(counts.astype(np.int32)), dst=0)
    if rank == 0:
        q.put((p_all.numpy(), c_all.numpy()))
    else:
        assert p_all is None and c_all is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process():
    import torch.multiprocessing as mp
    from oracle import oracle as O
    from sdrpp_tetra_demodulator_b200.sharding import unpack_dibits
    n_channels, n_samples, world = 6, 6000, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_channels, n_samples, q)) for r in range(world)]
    for p in procs:
        p.start()
    packed_all, counts_all = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = O.OracleB(n_channels)
    counts, _, dibits, _ = ref.process(O.generate(n_channels, n_samples), want_syms=False)
    assert np.array_equal(counts_all, counts)
    for c in range(n_channels):
        assert np.array_equal(unpack_dibits(packed_all[c], int(counts[c])), dibits[c, :counts[c]])
