"""GPU tests of tdm_process_long (BASELINE.json configs[1]): one long capture of one channel
demodulated as overlapping time segments in parallel.  Contract (include/tdm_b200.h): the decoded dibits equal the
sequential chain's from the point where the sequential chain has locked."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _lock_index(dibits_seq, tx_lagged):
    """first symbol from which the sequential chain's dibits equal the transmitted ones to the end"""
    bad = np.flatnonzero(dibits_seq != tx_lagged)
    return 0 if len(bad) == 0 else int(bad.max()) + 1


def _sequential(O, iq):
    ob = O.OracleB(1)
    cb, _, db, _ = ob.process(iq[None])
    return db[0, :cb[0]]


@pytest.mark.parametrize("snr_db,channel", [(30.0, 0), (30.0, 3), (20.0, 1)])
def test_segmented_equals_sequential_after_lock(O, pkg, torch_cuda, snr_db, channel):
    torch = torch_cuda
    N = 1_200_000
    iq = O.generate(1, N, O.default_sg_params(snr_db=snr_db), first_channel=channel)[0]
    seq = _sequential(O, iq)                                     # the canonical-order oracle, one channel, sequentially
    tx = O.tx_dibits(channel, N // 2)
    lag = min(range(10, 30), key=lambda L: int((seq[L + 50000:L + 60000] != tx[50000:60000]).sum()))
    lock = _lock_index(seq[lag:], tx[:len(seq) - lag]) + lag
    assert lock < 100_000, lock
    with pkg.Demodulator(16, 1024) as dm:
        got, info = dm.process_long(torch.from_numpy(iq).cuda(), warmup=32768)
        torch.cuda.synchronize()
        got = got.cpu().numpy()
    assert info["n_segments"] == 16 and info["warmup"] == 32768, info
    assert abs(len(got) - len(seq)) <= 1, (len(got), len(seq), info)
    n = min(len(got), len(seq))
    assert np.array_equal(got[lock:n], seq[lock:n]), (info, np.flatnonzero(got[lock:n] != seq[lock:n])[:10] + lock)
    # segment 0 IS the sequential chain: identical from the very first symbol
    first = info["segment_samples"] // 2
    assert np.array_equal(got[:first], seq[:first])


def test_rerun_path_when_the_warm_up_is_too_short(O, pkg, torch_cuda):
    """a 1024-sample warm-up cannot converge the loops: joins fail, those segments are redone as the sequential
    continuation of their predecessor (one per failed run and pass), successors re-join against the new stream,
    and the result still equals the sequential chain after lock -- no symbol lost or doubled at any boundary"""
    torch = torch_cuda
    N = 1_200_000
    iq = O.generate(1, N)[0]
    seq = _sequential(O, iq)
    tx = O.tx_dibits(0, N // 2)
    lag = min(range(10, 30), key=lambda L: int((seq[L + 50000:L + 60000] != tx[50000:60000]).sum()))
    lock = _lock_index(seq[lag:], tx[:len(seq) - lag]) + lag
    with pkg.Demodulator(8, 1024) as dm:
        got, info = dm.process_long(torch.from_numpy(iq).cuda(), warmup=1024)
        torch.cuda.synchronize()
        got = got.cpu().numpy()
    assert info["n_segments"] == 8 and info["n_rerun"] > 0, info
    assert abs(len(got) - len(seq)) <= 1, (len(got), len(seq), info)
    n = min(len(got), len(seq))
    lock = max(lock, info["segment_samples"] // 2)      # segment 1 onwards starts from states that had a whole segment to settle
    assert np.array_equal(got[lock:n], seq[lock:n]), (info, np.flatnonzero(got[lock:n] != seq[lock:n])[:10] + lock)


def test_short_capture_is_the_plain_sequential_call(O, pkg, torch_cuda):
    """too short to be worth splitting: one segment, bit-identical to tdm_process from the first symbol, and the
    loop state carried into the next call exactly"""
    torch = torch_cuda
    N = 90_000
    iq = O.generate(1, N)[0]
    seq = _sequential(O, iq)
    with pkg.Demodulator(8, 1024) as dm:
        a, ia = dm.process_long(torch.from_numpy(iq[:50_000]).cuda().contiguous())
        b, ib = dm.process_long(iq[50_000:])                    # host buffers for the second call
        torch.cuda.synchronize()
        got = np.concatenate([a.cpu().numpy(), b])
    assert ia["n_segments"] == 1 and ib["n_segments"] == 1 and ia["warmup"] == 0
    assert np.array_equal(got, seq)


def test_consecutive_calls_continue_one_stream(O, pkg, torch_cuda):
    torch = torch_cuda
    N = 1_000_000
    iq = O.generate(1, N, first_channel=2)[0]
    seq = _sequential(O, iq)
    with pkg.Demodulator(8, 1024) as dm:
        parts = []
        for lo, hi in [(0, 400_000), (400_000, 1_000_000)]:
            d, info = dm.process_long(torch.from_numpy(iq[lo:hi]).cuda().contiguous(), warmup=20_000)
            parts.append(d.cpu().numpy())
            assert info["n_segments"] >= 4
        got = np.concatenate(parts)
    assert abs(len(got) - len(seq)) <= 1
    n = min(len(got), len(seq))
    assert np.array_equal(got[100_000:n], seq[100_000:n])


def test_stretch_without_signal(O, pkg, torch_cuda):
    """a dead stretch several segments long: a run that ends inside it is not locked (DQPSKSymbolExtractor::sync down),
    so there is nothing for its successor to agree with and the join is made at the nominal place.  Before the gap the
    stream equals the sequential one exactly; after it, once both have locked again, up to the few symbols the gap
    may have gained or lost"""
    torch = torch_cuda
    N = 2_400_000
    iq = O.generate(1, N)[0].copy()
    rng = np.random.default_rng(1)
    iq[500_000:1_500_000] = (1e-4 * rng.standard_normal((1_000_000, 2))).astype(np.float32)
    seq = _sequential(O, iq)
    with pkg.Demodulator(16, 1024) as dm:
        got, info = dm.process_long(torch.from_numpy(iq).cuda(), warmup=40_000)
        torch.cuda.synchronize()
        got = got.cpu().numpy()
    assert info["n_segments"] >= 12 and info["n_forced"] >= 3, info
    # without signal the timing loop free-runs (omega wanders inside its +-2 % limits), so how many garbage symbols a
    # run emits across the gap is its own business: the two streams may differ by a fraction of a percent of the gap
    assert abs(len(got) - len(seq)) <= 5000, (len(got), len(seq), info)
    assert np.array_equal(got[100_000:240_000], seq[100_000:240_000])
    # after the gap, once both have locked again: identical.  Both streams end at the capture's last sample, so they
    # are compared aligned at the END (give or take the symbol that can fall either side of it)
    tail = 150_000
    ref = seq[len(seq) - tail - 4:len(seq) - 4]
    best = min(int((got[len(got) - tail - 4 - d:len(got) - 4 - d] != ref).sum()) for d in range(-3, 4))
    assert best == 0, best


def test_batch_of_long_captures(O, pkg, torch_cuda):
    """tdm_process_long_batch (BASELINE.json configs[2] in small): every channel cut into rows // C segments; each
    channel's stream equals its own sequential chain after lock; a second call continues every channel; host buffers
    give the same as device buffers"""
    torch = torch_cuda
    C_, N = 3, 900_000
    iq = O.generate(C_, N, first_channel=4)
    seqs = [_sequential(O, iq[c]) for c in range(C_)]
    with pkg.Demodulator(24, 1024) as dm, pkg.Demodulator(24, 1024) as dm_host:
        dib, cnt, info = dm.process_long_batch(torch.from_numpy(iq).cuda(), warmup=20_000)
        torch.cuda.synchronize()
        assert info["n_segments"] == 8 and info["warmup"] == 20_000, info
        dib, cnt = dib.cpu().numpy(), cnt.cpu().numpy()
        hd, hc, hinfo = dm_host.process_long_batch(iq, warmup=20_000)
        assert np.array_equal(hc, cnt) and all(np.array_equal(hd[c, :cnt[c]], dib[c, :cnt[c]]) for c in range(C_))
        assert info["n_dibits"] == int(cnt.sum())
        for c in range(C_):
            got, seq = dib[c, :cnt[c]], seqs[c]
            assert abs(len(got) - len(seq)) <= 1, (c, len(got), len(seq), info)
            n = min(len(got), len(seq))
            assert np.array_equal(got[100_000:n], seq[100_000:n]), (c, info)
            assert np.array_equal(got[:info["segment_samples"] // 2], seq[:info["segment_samples"] // 2])      # first segments ARE sequential
        # second call: every channel continues its own stream
        more = O.generate(C_, N, first_channel=4, n0=N)               # the generator is a function of the sample index
        full = [_sequential(O, np.concatenate([iq[c], more[c]])) for c in range(C_)]
        dib2, cnt2, info2 = dm.process_long_batch(torch.from_numpy(np.ascontiguousarray(more)).cuda(), warmup=20_000)
        torch.cuda.synchronize()
        dib2, cnt2 = dib2.cpu().numpy(), cnt2.cpu().numpy()
        for c in range(C_):
            got = np.concatenate([dib[c, :cnt[c]], dib2[c, :cnt2[c]]])
            assert abs(len(got) - len(full[c])) <= 1
            n = min(len(got), len(full[c]))
            assert np.array_equal(got[100_000:n], full[c][100_000:n]), c


def test_argument_checks(pkg, torch_cuda):
    torch = torch_cuda
    with pkg.Demodulator(4, 1024) as dm:
        iq = torch.zeros((5000, 2), dtype=torch.float32, device="cuda")
        with pytest.raises(pkg.TdmError):
            dm.process_long(iq, warmup=10)                      # warm-up below the minimum
        d, info = dm.process_long(iq[:0].contiguous())
        assert info["n_dibits"] == 0
