"""The multi-GPU epilogue through the C ABI (tdm_comm_*, tdm_gather_packed, tdm_unpack_dibits) on real GPUs.

With one visible GPU the communicator has one rank (NCCL refuses two ranks on one device): the whole API surface runs,
the data path is the local copy.  With two or more GPUs two ranks are spawned and rank 0 checks that what arrived over
NCCL, unpacked on the device, is every rank's dibit stream and the bit stream BitUnpacker would emit
(src/dsp/bit_unpacker.cpp:4-10) -- the NETSYMS payload (src/main.cpp:385-389)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_ch, n_s, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    try:
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        import sdrpp_tetra_demodulator_b200 as pkg
        from sdrpp_tetra_demodulator_b200.sharding import Communicator, channel_range
        first, last = channel_range(rank, world, n_ch * world)
        assert last - first == n_ch
        iq, _ = pkg.synth_capture(n_ch, n_s, device=rank, first_channel=first)
        comm = Communicator(rank)
        with pkg.Demodulator(n_ch, 1024, device=rank) as dm:
            dm.use_torch_stream()
            r = dm.process(iq, dibits=True, bits=True, packed=True)
            g_packed, g_counts = comm.gather_packed(r.packed, r.counts, dst=0)
            torch.cuda.synchronize()
            # every rank ships its plain streams to rank 0 through torch as the independent route to compare against
            S = r.dibits.shape[1]
            dall = [torch.empty_like(r.dibits) for _ in range(world)] if rank == 0 else None
            ball = [torch.empty_like(r.bits) for _ in range(world)] if rank == 0 else None
            dist.gather(r.dibits, dall, dst=0)
            dist.gather(r.bits, ball, dst=0)
            ok, msg = True, ""
            if rank == 0:
                ud, ub = dm.unpack_dibits(g_packed, g_counts, max_symbols=S, dibits=True, bits=True)
                torch.cuda.synchronize()
                cnt = g_counts.cpu().numpy()
                ud, ub = ud.cpu().numpy(), ub.cpu().numpy()
                for rr in range(world):
                    dd, bb = dall[rr].cpu().numpy(), ball[rr].cpu().numpy()
                    for c in range(n_ch):
                        n = int(cnt[rr * n_ch + c])
                        if n < n_s // 2 - 4 or not np.array_equal(ud[rr * n_ch + c, :n], dd[c, :n]) or not np.array_equal(ub[rr * n_ch + c, :2 * n], bb[c, :2 * n]):
                            ok, msg = False, f"rank {rr} channel {c}: gathered stream differs (n={n})"
            q.put((rank, ok, msg))
        comm.close()
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:          # surface the failure in the parent
        q.put((rank, False, f"{type(e).__name__}: {e}"))


def test_gather_packed_over_nccl(pkg):
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    world = 2 if torch.cuda.device_count() >= 2 else 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 40, 30000, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, msg in res:
        assert ok, f"rank {rank}: {msg}"


def test_comm_argument_errors(pkg):
    L = pkg.capi.lib()
    ident = (C.c_uint8 * 128)()
    h = C.c_void_p()
    assert L.tdm_comm_create(ident, 3, 2, 0, C.byref(h)) == pkg.capi.TDM_ERR_ARG          # rank >= world
    assert L.tdm_comm_create(None, 0, 1, 0, C.byref(h)) == pkg.capi.TDM_ERR_ARG
    assert L.tdm_gather_packed(None, 0, 1, None, 0, None, None, None, None) == pkg.capi.TDM_ERR_ARG
    assert L.tdm_comm_destroy(None) == pkg.capi.TDM_OK
