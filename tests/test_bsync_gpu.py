"""GPU parity tests of the burst-sync stage (include/tdm_burst_b200.h) through the C ABI: bit-exact against the
reference's own phy/tetra_burst.c + phy/tetra_burst_sync.c compiled unmodified (oracle/_ref, when present), the
restatement (oracle/oracle_bsync.c) and the committed golden fixture."""
import os

import numpy as np
import pytest

from oracle import oracle_bsync as B

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "bsync_c6.npz")
BURST_FIELDS = ["bitnum", "train_seq", "tn", "fn", "mn", "call_index", "bits"]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    B.build()
    return torch


def _rows(streams):
    n = np.array([len(s) for s in streams], dtype=np.int32)
    m = np.zeros((len(streams), max(int(n.max()), 1)), dtype=np.uint8)
    for c, s in enumerate(streams):
        m[c, :len(s)] = s
    return m, n


def _assert_bursts(pkg, nb, bursts, nb_ref, bursts_ref, what=""):
    nb = nb.cpu().numpy() if hasattr(nb, "cpu") else nb
    bv = pkg.bursts_view(bursts)
    assert np.array_equal(nb, nb_ref), (what, nb, nb_ref)
    for c in range(len(nb)):
        for i in range(min(int(nb[c]), bv.shape[1])):
            for f in BURST_FIELDS:
                assert np.array_equal(bv[c, i][f], bursts_ref[c, i][f]), (what, f, c, i, bv[c, i][f], bursts_ref[c, i][f])


def _assert_state(st, port, fields=B.STATE_COMPARE):
    for f in fields:
        assert np.array_equal(st[f], port.states[f]), (f, st[f], port.states[f])


def _mixed_streams(n_channels, n_slots=24):
    out = []
    for c in range(n_channels):
        kind = c % 6
        if kind == 5:
            out.append(np.random.default_rng(900 + c).integers(0, 2, 510 * n_slots).astype(np.uint8))     # noise
        else:
            out.append(B.downlink_stream(300 + c, n_slots, ber=[0, 1e-3, 0, 8e-3, 0][kind],
                                         glitch_at=[(), (), (5,), (9, 15), ()][kind]))
    return out


@pytest.mark.parametrize("call_bits", [1, 7, 100, 432, 509, 510])
@pytest.mark.parametrize("device_buffers", [True, False])
def test_state_machine_bit_exact(pkg, torch_cuda, call_bits, device_buffers):
    """ragged per-channel lengths, two feeds with carried state, device and host buffers; vs restatement and reference"""
    torch = torch_cuda
    C_ = 13
    streams = _mixed_streams(C_)
    port = B.PortBsync(C_)
    ref = B.RefBsync(C_) if B.have_ref() else None
    with pkg.BurstSync(C_, 8000) as bs:
        for part in range(2):
            cut = [len(s) // 2 + 3 * c for c, s in enumerate(streams)]
            rows, n = _rows([s[:k] if part == 0 else s[k:] for s, k in zip(streams, cut)])
            nb_p, bu_p = port.feed(rows, n, call_bits, 40, detect_ts=True)
            if device_buffers:
                nb, bu = bs.feed(torch.from_numpy(rows).cuda(), torch.from_numpy(n).cuda(), call_bits=call_bits, max_bursts=40, detect_ts=True)
                torch.cuda.synchronize()
            else:
                nb, bu = bs.feed(rows, n, call_bits=call_bits, max_bursts=40, detect_ts=True)
            _assert_bursts(pkg, nb, bu, nb_p, bu_p, "vs restatement")
            st = bs.get_state()
            _assert_state(st, port)
            _assert_state(st, port, B.TS_COMPARE)
            if ref is not None:
                nb_r, bu_r, _ = ref.feed(rows, n, call_bits, 40)
                _assert_bursts(pkg, nb, bu, nb_r, bu_r, "vs reference")
                for c in range(C_):
                    rs = ref.state(c)
                    for f in rs:
                        assert np.array_equal(st[c][f], rs[f]), (f, c)
        assert int(st["n_bursts"].sum()) > 100 or call_bits > 432
    if ref is not None:
        ref.close()


def test_dibit_input_equals_bit_input(pkg, torch_cuda):
    """TDM_BSYNC_IN_DIBITS: DQPSKSymbolExtractor's output fed directly; odd carried lengths exercise the re-alignment"""
    torch = torch_cuda
    C_ = 8
    streams = [s[:len(s) // 2 * 2] for s in _mixed_streams(C_, 16)]
    with pkg.BurstSync(C_, 9000) as a, pkg.BurstSync(C_, 9000) as b:
        pos = [0] * C_
        for step, take in enumerate([1001, 77, 2 * 510, 3001, 10 ** 6]):
            parts = [s[p:p + take // 2 * 2] for s, p in zip(streams, pos)]
            pos = [p + len(x) for p, x in zip(pos, parts)]
            rows, n = _rows(parts)
            drows, dn = _rows([B.bits_to_dibits(x) for x in parts])
            nb1, bu1 = a.feed(torch.from_numpy(rows).cuda(), torch.from_numpy(n).cuda(), call_bits=333, max_bursts=20, detect_ts=True)
            nb2, bu2 = b.feed(torch.from_numpy(drows).cuda(), torch.from_numpy(dn).cuda(), dibits=True, call_bits=333, max_bursts=20, detect_ts=True)
            torch.cuda.synchronize()
            assert torch.equal(nb1, nb2)
            v1, v2 = pkg.bursts_raw(bu1), pkg.bursts_raw(bu2)
            for c in range(C_):
                k = int(nb1[c])
                assert v1[c, :k].tobytes() == v2[c, :k].tobytes()
            s1, s2 = a.get_state(), b.get_state()
            for f in B.STATE_COMPARE + B.TS_COMPARE:
                assert np.array_equal(s1[f], s2[f]), (f, step)


def test_golden_fixture(pkg, torch_cuda):
    """the bursts the REFERENCE delivered for the committed bit streams"""
    torch = torch_cuda
    g = np.load(GOLDEN)
    n_bits = g["n_bits"]
    bits = np.ascontiguousarray(np.unpackbits(g["bits"], axis=1)[:, :int(n_bits.max())])
    with pkg.BurstSync(bits.shape[0], bits.shape[1]) as bs:
        nb, bu = bs.feed(torch.from_numpy(bits).cuda(), torch.from_numpy(n_bits).cuda(), call_bits=int(g["call_bits"]),
                         max_bursts=g["bursts"].shape[1])
        torch.cuda.synchronize()
        _assert_bursts(pkg, nb, bu, g["n_bursts"], g["bursts"], "golden")
        st = bs.get_state()
        for f in ["state", "bits_in_buf", "bitbuf_start_bitnum", "next_frame_start_bitnum", "tn", "fn", "mn"]:
            assert np.array_equal(st[f], g["final_" + f]), f


def test_find_train_seq_bit_exact(pkg, torch_cuda):
    """tdm_find_train_seq incl. the look-ahead quirk (first 21 positions), every mask, ragged ends, host + device"""
    torch = torch_cuda
    rng = np.random.default_rng(11)
    port = B.PortBsync(1)
    ref = B.RefBsync(1) if B.have_ref() else None
    for trial in range(12):
        C_, L = 64, int(rng.integers(1, 900))
        bufs = rng.integers(0, 2, (C_, L + 5)).astype(np.uint8)
        for c in range(C_):
            for _ in range(int(rng.integers(0, 3))):
                s = B.SEQ[str(rng.choice(list("npqxy")))]
                p = int(rng.integers(0, 40 if rng.random() < 0.4 else L))
                m = min(len(s), L - p)
                if m > 0:
                    bufs[c, p:p + m] = s[:m]
        end, mask = int(rng.integers(0, L + 1)), int(rng.integers(1, 32))
        if trial % 2:
            typ, off = pkg.find_train_seq(torch.from_numpy(bufs).cuda(), end, mask)
            torch.cuda.synchronize()
            typ, off = typ.cpu().numpy(), off.cpu().numpy().astype(np.uint32)
        else:
            typ, off = pkg.find_train_seq(bufs, end, mask)
        for c in range(C_):
            want = port.find_train_seq(bufs[c], end, mask)
            assert typ[c] == want[0] and (want[0] < 0 or off[c] == want[1]), (trial, c, typ[c], off[c], want)
            if ref is not None:
                assert ref.find_train_seq(bufs[c], end, mask)[0] == typ[c]
    if ref is not None:
        ref.close()


def test_edge_cases(pkg, torch_cuda):
    """empty input, one bit at a time, all-zero / all-one streams, more bursts than records, checkpoint / resume"""
    torch = torch_cuda
    C_ = 4
    streams = [B.downlink_stream(40 + c, 20, lead_bits=100 * c) for c in range(C_)]
    rows, n = _rows(streams)
    port = B.PortBsync(C_)
    nb_p, bu_p = port.feed(rows, n, 510, 64)
    with pkg.BurstSync(C_, rows.shape[1]) as bs:
        nb, _ = bs.feed(torch.from_numpy(rows).cuda(), 0, call_bits=510, max_bursts=0)            # nothing to do
        torch.cuda.synchronize()
        assert int(nb.sum()) == 0 and int(bs.get_state()["n_bits"].sum()) == 0
        nb, bu = bs.feed(torch.from_numpy(rows).cuda(), torch.from_numpy(n).cuda(), call_bits=510, max_bursts=3)   # records overflow
        torch.cuda.synchronize()
        assert np.array_equal(nb.cpu().numpy(), nb_p) and int(nb.min()) > 3
        v = pkg.bursts_view(bu)
        for c in range(C_):
            for f in BURST_FIELDS:
                assert np.array_equal(v[c, :3][f], bu_p[c, :3][f])
        # checkpoint in the middle, resume in a fresh handle
        bs.reset()
        half = rows.shape[1] // 2
        bs.feed(torch.from_numpy(np.ascontiguousarray(rows[:, :half])).cuda(), None, call_bits=510, max_bursts=0)
        saved = bs.get_state()
    with pkg.BurstSync(C_, rows.shape[1]) as bs2:
        bs2.set_state(saved)
        rest = np.ascontiguousarray(rows[:, half:])
        nrest = np.maximum(n - half, 0).astype(np.int32)
        bs2.feed(torch.from_numpy(rest).cuda(), torch.from_numpy(nrest).cuda(), call_bits=510, max_bursts=0)
        port2 = B.PortBsync(C_)
        port2.feed(np.ascontiguousarray(rows[:, :half]), half, 510, 0)
        port2.feed(rest, nrest, 510, 0)
        _assert_state(bs2.get_state(), port2)
    for fill in (0, 1):
        const = np.full((2, 20000), fill, dtype=np.uint8)
        p3 = B.PortBsync(2)
        p3.feed(const, 20000, 1, 0, detect_ts=True)
        with pkg.BurstSync(2, 20000) as bs3:
            bs3.feed(const, None, call_bits=1, max_bursts=0, detect_ts=True)
            _assert_state(bs3.get_state(), p3)
            _assert_state(bs3.get_state(), p3, B.TS_COMPARE)
    with pytest.raises(pkg.TdmError):
        with pkg.BurstSync(2, 100) as bs4:
            bs4.feed(np.zeros((2, 100), np.uint8), None, call_bits=511)


def test_ts_detector_chunked(pkg, torch_cuda):
    """src/main.cpp:385-414 detector: the state after any chunking equals the restatement's, incl. sequences that
    straddle feeds and the 2048-bit expiry"""
    torch = torch_cuda
    rng = np.random.default_rng(21)
    C_ = 6
    bits = rng.integers(0, 2, (C_, 12000)).astype(np.uint8)
    names = list(B.SEQ)
    for c in range(C_):
        for pos in [int(rng.integers(0, 11900)) for _ in range(c)]:
            s = B.SEQ[names[int(rng.integers(0, 8))]]
            m = min(len(s), 12000 - pos)
            bits[c, pos:pos + m] = s[:m]
    for chunk in (12000, 1, 31, 45, 1000, 2048):
        port = B.PortBsync(C_)
        with pkg.BurstSync(C_, 12000) as bs:
            steps = range(0, 12000, chunk) if chunk > 1 else range(0, 300)
            for p0 in steps:
                part = np.ascontiguousarray(bits[:, p0:p0 + chunk])
                port.feed(part, part.shape[1], 510, 0, detect_ts=True)
                bs.feed(torch.from_numpy(part).cuda(), None, call_bits=510, max_bursts=0, detect_ts=True)
                if chunk >= 1000 or p0 % 7 == 0:
                    _assert_state(bs.get_state(), port, B.TS_COMPARE)
            _assert_state(bs.get_state(), port, B.TS_COMPARE + B.STATE_COMPARE)


def _modulate(pkg, dibits, amp=0.7, df_hz=120.0, snr_db=30.0, seed=0):
    """pi/4-DQPSK, TETRA map of src/decoder/src/phy/tetra_burst.c:99-104, RRC-shaped with the demodulator's own taps, 2 sps"""
    d = pkg.design_from_config(pkg.default_config())
    rrc = np.array(d.rrc[:], dtype=np.float64)
    dphi = np.array([1, 3, -1, -3])[dibits] * (np.pi / 4)        # dibit value (b1<<1|b2): 00->+pi/4, 01->+3pi/4, 10->-pi/4, 11->-3pi/4
    ph = np.cumsum(dphi)
    up = np.zeros(2 * len(dibits), dtype=np.complex128)
    up[::2] = np.exp(1j * ph)
    x = np.convolve(up, 2.0 * rrc)[:len(up)]
    n = np.arange(len(x))
    x = amp * x * np.exp(2j * np.pi * df_hz / 36000.0 * n)
    rng = np.random.default_rng(seed)
    sigma = amp * 10 ** (-snr_db / 20) / np.sqrt(2) * np.sqrt(2)
    x = x + sigma * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x))) / np.sqrt(2)
    return np.stack([x.real, x.imag], axis=1).astype(np.float32)


def test_demodulator_into_burst_sync(pkg, torch_cuda):
    """the two stages chained on the device as the plugin chains them (src/main.cpp:84-95): protocol-valid downlink
    bit streams -> IQ -> tdm_process (dibits stay in HBM) -> tdm_bsync_in(TDM_BSYNC_IN_DIBITS) -> bursts whose
    payload equals what was transmitted"""
    torch = torch_cuda
    C_, n_slots = 4, 60
    tx = [B.downlink_stream(70 + c, n_slots, lead_bits=200 * c + 2 * c) for c in range(C_)]
    tx = [t[:len(t) // 2 * 2] for t in tx]
    n_sym = min(len(t) for t in tx) // 2
    iq = np.stack([_modulate(pkg, B.bits_to_dibits(t)[:n_sym], amp=0.3 + 0.4 * c, df_hz=-200 + 130 * c, seed=c) for c, t in enumerate(tx)])
    N = iq.shape[1]
    with pkg.Demodulator(C_, N) as dm, pkg.BurstSync(C_, dm.max_symbols(N)) as bs:
        dm.use_torch_stream()
        bs.use_torch_stream()
        r = dm.process(torch.from_numpy(iq).cuda(), dibits=True)
        nb, bu = bs.feed(r.dibits, r.counts, dibits=True, call_bits=432, max_bursts=n_slots)
        torch.cuda.synchronize()
        v = pkg.bursts_view(bu)
        nb = nb.cpu().numpy()
        assert int(nb.min()) >= n_slots // 2, nb                      # locked for most of the capture
        for c in range(C_):
            txs = np.array2string(tx[c], separator="", threshold=10 ** 9, max_line_width=10 ** 9)[1:-1]
            good = 0
            for i in range(int(nb[c])):
                payload = np.array2string(v[c, i]["bits"][:510], separator="", threshold=10 ** 9, max_line_width=10 ** 9)[1:-1]
                good += payload in txs
            assert good >= int(nb[c]) - 1, (c, good, nb[c])          # every delivered burst is a transmitted slot, bit for bit
        # and the same bits through the restatement give the same bursts
        dib = r.dibits.cpu().numpy()
        cnt = r.counts.cpu().numpy()
        port = B.PortBsync(C_)
        rows, n = _rows([np.stack([(dib[c, :cnt[c]] >> 1) & 1, dib[c, :cnt[c]] & 1], axis=1).reshape(-1) for c in range(C_)])
        nb_p, bu_p = port.feed(rows, n, 432, n_slots)
        _assert_bursts(pkg, nb, bu, nb_p, bu_p, "chained")
