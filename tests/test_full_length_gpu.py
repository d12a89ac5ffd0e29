"""GPU parity at BASELINE.json's full sizes (run with -m gpu on the B200 box; through the C ABI).

The oracle is too slow for every channel of a bench-size capture, so the capture is generated on the device at full
size, demodulated at full size, and SAMPLED channels are copied back and compared over their FULL length:
bit-exact against the canonical-order checker (dibits and loop state), and against the reference's own code
(oracle/_ref) from the reference's lock point on.
  * configs[2]: 256 channels x 4e6 samples, plain batch call
  * configs[3] per-GPU shard shapes: 512 and 4096 channels x 4e6 samples (the 4096-channel one is the bench workload)
  * configs[1] in kind: 1 channel x 1e8 samples through tdm_process_long against the SEQUENTIAL oracle
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _lock_index(O, dibits, channel, n):
    tx = O.tx_dibits(channel, n + 64)
    best = None
    for lag in range(10, 30):
        e = np.flatnonzero(dibits[lag:n] != tx[:n - lag])
        last = int(e[-1]) + lag + 1 if len(e) else lag
        best = last if best is None or last < best else best
    return best


@pytest.mark.parametrize("n_channels,n_sampled", [(256, 16), (512, 8), (4096, 8)])
def test_sampled_channels_full_length(O, pkg, torch_cuda, n_channels, n_sampled):
    torch = torch_cuda
    N = 4_000_000
    free, _ = torch.cuda.mem_get_info()
    if free < n_channels * N * 8 * 1.15:
        pytest.skip("not enough device memory for this shape")
    cores = os.cpu_count() or 1
    iq, _ = pkg.synth_capture(n_channels, N)
    rng = np.random.default_rng(n_channels)
    idx = np.sort(rng.choice(n_channels, size=n_sampled, replace=False))
    with pkg.Demodulator(n_channels, 1024) as dm:
        dm.use_torch_stream()
        r = dm.process(iq, dibits=True, packed=True)
        torch.cuda.synchronize()
        tidx = torch.from_numpy(idx).to(iq.device)
        rows = iq[tidx].cpu().numpy()
        got = r.dibits[tidx].cpu().numpy()
        pk = r.packed[tidx].cpu().numpy()
        cnt = r.counts[tidx].cpu().numpy()
        st = dm.get_state()
        allc = r.counts.cpu().numpy()
        assert allc.min() >= N // 2 - 4 and allc.max() <= N // 2 + 4
    del iq, r
    torch.cuda.empty_cache()
    ob = O.OracleB(n_sampled)
    cb, _, db, _ = ob.process(rows, want_syms=False, nthreads=cores)
    assert np.array_equal(cb, cnt)
    for k in range(n_sampled):
        n = int(cb[k])
        assert np.array_equal(got[k, :n], db[k, :n]), f"channel {idx[k]}: dibits differ from the canonical-order checker"
        d4 = np.stack([(pk[k] >> 6) & 3, (pk[k] >> 4) & 3, (pk[k] >> 2) & 3, pk[k] & 3], axis=1).reshape(-1)[:n]
        assert np.array_equal(d4, db[k, :n]), f"channel {idx[k]}: packed output differs"
    for f in O.EXACT_STATE_FIELDS:
        a, b = st[f][idx], ob.states[f]
        assert np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8)), f
    if O.have_ref():
        oa = O.OracleA(n_sampled)
        ca, da = oa.process_multi(rows, cores)
        oa.close()
        for k in range(n_sampled):
            n = min(int(ca[k]), int(cnt[k]))
            assert abs(int(ca[k]) - int(cnt[k])) <= 1
            lock = _lock_index(O, da[k], int(idx[k]), n)
            assert lock < n // 2, f"the reference itself did not lock on channel {idx[k]}"
            assert np.array_equal(got[k, lock:n], da[k, lock:n]), f"channel {idx[k]}: differs from the reference after its lock point {lock}"
            assert np.count_nonzero(got[k, :lock] != da[k, :lock]) <= max(8, lock // 500)


def test_long_capture_1e8_against_the_sequential_oracle(O, pkg, torch_cuda):
    """BASELINE.json configs[1] in kind (1 channel, one long capture; 1e8 samples keeps the sequential CPU oracle at
    ~25 s): tdm_process_long's dibits equal the SEQUENTIAL chain's from its lock point to the end, and exactly from
    symbol 0 up to the first join (segment 0 is the sequential chain)."""
    torch = torch_cuda
    N = 100_000_000
    iq, _ = pkg.synth_capture(1, N)
    with pkg.Demodulator(1024, 1024) as dm:
        dm.use_torch_stream()
        d, info = dm.process_long(iq[0])
        torch.cuda.synchronize()
        got = d.cpu().numpy()
    row = iq.cpu().numpy()
    del iq
    torch.cuda.empty_cache()
    ob = O.OracleB(1)
    cb, _, db, _ = ob.process(row, want_syms=False)
    n = min(int(cb[0]), len(got))
    assert abs(int(cb[0]) - len(got)) <= 2 * max(1, info["n_forced"] + 1)
    lock = _lock_index(O, db[0], 0, n)
    assert lock < 200_000
    first_join = info["segment_samples"] // 2 - 64
    assert np.array_equal(got[:first_join], db[0, :first_join])            # segment 0 IS the sequential chain
    if info["n_forced"] == 0:
        assert np.array_equal(got[lock:n], db[0, lock:n]), "stitched stream differs from the sequential chain after lock"
    assert info["n_segments"] >= 256
