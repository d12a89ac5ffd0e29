"""Front-end channeliser (include/tdm_chan_b200.h) on the GPU.

PARITY UNPINNED BY THE REFERENCE (it has no channeliser: src/main.cpp:75 asks SDR++ for a VFO).  Checked against the
float64 defining sum (oracle/oracle_chan.py) with tolerance 2e-5 of the output's RMS (fp32 accumulation over T M
taps and an M-point fp32 FFT), for any chunking of the stream, and end to end: narrowband TETRA captures placed on the
25 kHz raster of one wideband capture come back out of channeliser -> demodulator as the transmitted dibits."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def test_matches_the_defining_sum(pkg, torch_cuda):
    torch = torch_cuda
    from oracle import oracle_chan as OC
    cfg = pkg.chan_default_config(2)             # M = 72, D = 50, T = 16
    M, D = cfg.n_channels, cfg.decimation
    h = pkg.chan_design(cfg)
    rng = np.random.default_rng(3)
    N = D * 400
    x = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
    x += 30 * np.exp(2j * np.pi * (5.2 / M) * np.arange(N)).astype(np.complex64)          # a strong carrier near channel 5
    wide = torch.from_numpy(np.stack([x.real, x.imag], axis=1).copy()).cuda()
    with pkg.Channelizer(cfg) as ch:
        y = ch.process(wide)
        torch.cuda.synchronize()
        got = y.cpu().numpy()
        got = got[..., 0] + 1j * got[..., 1]
        # the same stream in ragged calls (multiples of D) must give the same samples
        ch.reset()
        parts, pos = [], 0
        for k in (1, 7, 64, 3, 325):
            parts.append(ch.process(wide[pos:pos + k * D].contiguous()).cpu().numpy())
            pos += k * D
        torch.cuda.synchronize()
        chunked = np.concatenate(parts, axis=1)
        assert np.array_equal(chunked, y.cpu().numpy()), "chunked calls differ from the single call"
    chans = [0, 1, 5, 6, 36, 71]
    inst = [0, 1, 2, 17, 18, 150, 399]
    want = OC.channelize_direct(x.astype(np.complex128), h.astype(np.float64), M, D, chans, inst)
    rms = np.sqrt(np.mean(np.abs(got) ** 2))
    err = np.abs(got[np.ix_(chans, inst)] - want).max()
    assert err < 2e-5 * max(rms, 1.0), (err, rms)
    # the strong carrier sits 0.2 of a channel off channel 5's centre: channel 40 (far away) must not see it
    assert np.sqrt(np.mean(np.abs(got[40, 50:]) ** 2)) < 3.0


def test_wideband_to_dibits_end_to_end(O, pkg, torch_cuda):
    """6 TETRA carriers on a 144-channel raster (3.6 MS/s): wideband capture -> tdm_chan_process -> tdm_process (device buffers,
    same stream, no host round trip) -> decoded dibits == transmitted dibits after the chain's lag, for every carrier."""
    torch = torch_cuda
    from oracle import oracle_chan as OC
    cfg = pkg.chan_default_config(4)             # M = 144, D = 100
    M, D = cfg.n_channels, cfg.decimation
    n = 60000
    carriers = [3, 4, 40, 71, 100, 143]          # 3 and 4 are neighbours on the raster: each sees the other's skirt inside +-18 kHz
    sp = O.default_sg_params(snr_db=60.0, max_freq_off_hz=200.0, min_amp=0.5, max_amp=1.0)
    nb = O.generate(len(carriers), n, sp)
    narrow = nb[..., 0] + 1j * nb[..., 1]
    wide = OC.place_on_raster(narrow, carriers, M, D)
    rng = np.random.default_rng(1)
    wide += 1e-3 * (rng.standard_normal(len(wide)) + 1j * rng.standard_normal(len(wide)))
    w = torch.from_numpy(np.stack([wide.real, wide.imag], axis=1).astype(np.float32)).cuda()
    with pkg.Channelizer(cfg) as ch, pkg.Demodulator(M, n) as dm:
        dm.use_torch_stream()
        y = ch.process(w)                          # [M][n][2] in HBM
        r = dm.process(y, dibits=True)             # every channel of the raster, occupied or not
        torch.cuda.synchronize()
        counts, dib = r.counts.cpu().numpy(), r.dibits.cpu().numpy()
        sync = dm.metrics()["sync"]
    for k, c in enumerate(carriers):
        tx = O.tx_dibits(k, n)
        cnt = int(counts[c])
        # the reference's loops can take more than 10^4 symbols to settle at an unlucky timing phase (the golden fixtures'
        # lock indices show the same; reproduced with the float64 channeliser + the CPU checker for the carrier on channel
        # 40, last error at symbol 13702): the property is stated on the last third of the capture
        first = 20000
        errs = {lag: np.flatnonzero(dib[c, lag + first:cnt - 8] != tx[first:cnt - 8 - lag]) for lag in range(10, 60)}
        lag = min(errs, key=lambda k: len(errs[k]))
        assert len(errs[lag]) == 0, f"carrier on channel {c}: {len(errs[lag])} dibit errors after symbol {first} (lag {lag}), at {errs[lag][:8] + first}"
        assert sync[c] == 1
    assert sync[[20, 50, 120]].sum() == 0          # empty channels do not report lock


def test_more_instants_than_a_grid_dimension(pkg, torch_cuda):
    """1.1 million output instants in one call (16 per block, a grid's y dimension ends at 65 535): the same samples as the
    stream cut in two calls."""
    torch = torch_cuda
    cfg = pkg.chan_default_config(1)             # M = 36, D = 25
    D, n_out = cfg.decimation, 1_100_000
    g = torch.Generator(device="cuda").manual_seed(5)
    wide = torch.randn((n_out * D, 2), generator=g, device="cuda", dtype=torch.float32)
    with pkg.Channelizer(cfg) as ch:
        whole = ch.process(wide)
        ch.reset()
        cut = 600_000 * D
        a = ch.process(wide[:cut].contiguous())
        b = ch.process(wide[cut:].contiguous())
        torch.cuda.synchronize()
        assert torch.equal(whole[:, :600_000], a) and torch.equal(whole[:, 600_000:], b)
        assert float(whole[:, -1000:].abs().max()) > 0


@pytest.mark.parametrize("M,D,T", [(40, 27, 16), (100, 73, 8), (96, 96, 12), (257, 64, 16)])
def test_other_rasters_match_the_defining_sum(pkg, torch_cuda, M, D, T):
    """Rasters other than TETRA's 36 : 25 -- a short branch period that is not 36 (40 : 27 -> 40), long periods that take the
    general path of the branch-sum kernel (100 : 73, 257 : 64), critical sampling (96 : 96), a slab that is not full
    (M not a multiple of 32) -- against the float64 defining sum, and chunked == whole."""
    torch = torch_cuda
    from oracle import oracle_chan as OC
    cfg = pkg.chan_default_config(1)
    cfg.n_channels, cfg.decimation, cfg.taps_per_branch = M, D, T
    cfg.passband, cfg.stopband = 0.4 * D / M * 2, 0.6 * D / M * 2
    h = pkg.chan_design(cfg)
    rng = np.random.default_rng(M + D)
    n_inst = 700
    N = D * n_inst
    x = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
    wide = torch.from_numpy(np.stack([x.real, x.imag], axis=1).copy()).cuda()
    with pkg.Channelizer(cfg) as ch:
        y = ch.process(wide)
        ch.reset()
        parts, pos = [], 0
        for k in (3, 260, 1, 436):
            parts.append(ch.process(wide[pos:pos + k * D].contiguous()).cpu().numpy())
            pos += k * D
        torch.cuda.synchronize()
        got = y.cpu().numpy()
        assert np.array_equal(np.concatenate(parts, axis=1), got), "chunked calls differ from the single call"
    got = got[..., 0] + 1j * got[..., 1]
    chans = sorted({0, 1, M // 3, M // 2, M - 1})
    inst = [0, 1, 2, 35, 36, 37, 255, 256, 257, 500, 699]
    want = OC.channelize_direct(x.astype(np.complex128), h.astype(np.float64), M, D, chans, inst)
    rms = np.sqrt(np.mean(np.abs(got) ** 2))
    err = np.abs(got[np.ix_(chans, inst)] - want).max()
    assert err < 2e-5 * max(rms, 1.0), (err, rms)


def test_instant_major_output_and_chain(O, pkg, torch_cuda):
    """tdm_chan_process_instant_major leaves [instant][channel] -- the batched DFT's own order, no transposing pass -- and the
    demodulator reads it in place (tdm_io.sample_stride): the same samples as the channel-major call, and the same
    dibits out of the chain."""
    torch = torch_cuda
    cfg = pkg.chan_default_config(2)             # 72 channels
    M, D = cfg.n_channels, cfg.decimation
    g = torch.Generator(device="cuda").manual_seed(11)
    wide = torch.randn((D * 3000, 2), generator=g, device="cuda", dtype=torch.float32)
    with pkg.Channelizer(cfg) as ch:
        a = ch.process(wide)
        ch.reset()
        b1 = ch.process(wide[:D * 1000].contiguous(), instant_major=True)
        b2 = ch.process(wide[D * 1000:].contiguous(), instant_major=True)
        torch.cuda.synchronize()
        b = torch.cat([b1, b2], dim=0)
        assert b.shape == (3000, M, 2)
        assert torch.equal(b.permute(1, 0, 2), a)
    with pkg.Demodulator(M, 3000) as d1, pkg.Demodulator(M, 3000) as d2:
        r1 = d1.process(a, dibits=True)
        r2 = d2.process(b, dibits=True, instant_major=True)
        torch.cuda.synchronize()
        assert torch.equal(r1.counts, r2.counts) and torch.equal(r1.dibits, r2.dibits)


def test_cs16_input_is_the_converted_floats(pkg, torch_cuda):
    """TDM_CHAN_IN_CS16: interleaved int16 in, taken as s / 32768 -- bit-identical to feeding those floats, in both output
    layouts, with the formats mixed from call to call (the carried history is floats)."""
    torch = torch_cuda
    cfg = pkg.chan_default_config(2)
    D = cfg.decimation
    g = torch.Generator(device="cuda").manual_seed(21)
    raw = torch.randint(-32768, 32768, (D * 900, 2), generator=g, device="cuda", dtype=torch.int32).to(torch.int16)
    as_float = raw.to(torch.float32) / 32768.0
    with pkg.Channelizer(cfg) as ch:
        want = ch.process(as_float)
        ch.reset()
        a = ch.process(raw[:D * 300].contiguous())                                  # int16
        b = ch.process(as_float[D * 300:D * 500].contiguous())                      # float32 in between
        c = ch.process(raw[D * 500:].contiguous(), instant_major=True)              # int16, the DFT's own layout
        torch.cuda.synchronize()
        got = torch.cat([a, b, c.permute(1, 0, 2)], dim=1)
        assert torch.equal(got, want)
