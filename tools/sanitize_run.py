#!/usr/bin/env python
"""Small pass through every kernel family, meant to be run under compute-sanitizer (SURVEY.md 5 "race detection /
sanitizers"):   compute-sanitizer --tool memcheck|racecheck|synccheck --error-exitcode 1 python tools/sanitize_run.py
Sizes are tiny on purpose (the tools slow kernels down by 10-100x); results are still checked against the oracles."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import sdrpp_tetra_demodulator_b200 as pkg
from oracle import oracle as O, oracle_bsync as B

C_, N = 5, 3000
iq = O.generate(C_, N)
ob = O.OracleB(C_)
cb, sb, db, bb = ob.process(iq, want_bits=True)
for variant in (2, 4, 8):                      # tpc8, ws4 (one CTA per SM), ws4 (two CTAs per SM build)
    with pkg.Demodulator(C_, N) as dm:
        dm.set_kernel_variant(variant)
        r = dm.process(torch.from_numpy(iq).cuda(), symbols=True, dibits=True, bits=True, packed=True)
        torch.cuda.synchronize()
        assert np.array_equal(r.counts.cpu().numpy(), cb)
        assert all(np.array_equal(r.dibits[c, :cb[c]].cpu().numpy(), db[c, :cb[c]]) for c in range(C_)), variant
with pkg.Demodulator(C_, 1000) as dm:          # host path with time slices + pack
    r = dm.process(iq[:, :1000].copy(), dibits=True)
    r2 = dm.process(torch.from_numpy(iq[:, 1000:]).cuda().contiguous(), dibits=True)
    pk = dm.pack_dibits(r2.dibits, r2.counts)
    dm.unpack_dibits(pk, r2.counts, dibits=True, bits=True)
    torch.cuda.synchronize()
g, _ = pkg.synth_capture(2, 2000, want_tx=True)
streams = [B.downlink_stream(5 + c, 6, ber=1e-3) for c in range(3)]
n = np.array([len(s) for s in streams], dtype=np.int32)
rows = np.zeros((3, int(n.max())), dtype=np.uint8)
for c, s in enumerate(streams):
    rows[c, :len(s)] = s
port = B.PortBsync(3)
nb_ref, _ = port.feed(rows, n, 432, 8, detect_ts=True)
with pkg.BurstSync(3, rows.shape[1]) as bs:
    nb, bu = bs.feed(torch.from_numpy(rows).cuda(), torch.from_numpy(n).cuda(), call_bits=432, max_bursts=8, detect_ts=True)
    nb2, _ = bs.feed(rows[:, :700].copy(), None, call_bits=100, max_bursts=8)
    torch.cuda.synchronize()
    assert np.array_equal(nb.cpu().numpy(), nb_ref)
pkg.find_train_seq(torch.from_numpy(rows[:, :600].copy()).cuda(), 600, 0x1f)
long_iq = O.generate(1, 60_000)[0]
with pkg.Demodulator(4, 1024) as dm:
    d, info = dm.process_long(torch.from_numpy(long_iq).cuda(), warmup=4096)
    d2, c2, info2 = dm.process_long_batch(torch.from_numpy(np.stack([long_iq, long_iq])).cuda(), warmup=4096)
    torch.cuda.synchronize()
    assert info["n_segments"] == 4 and info2["n_segments"] == 2, (info, info2)
# channeliser: the periodic (taps in registers) and the general path of the branch-sum kernel, partial slab, two calls
for (M, D, T) in ((72, 50, 16), (100, 73, 8)):
    cfg = pkg.chan_default_config(1)
    cfg.n_channels, cfg.decimation, cfg.taps_per_branch = M, D, T
    wide = torch.randn((D * 300, 2), device="cuda")
    with pkg.Channelizer(cfg) as ch:
        a = ch.process(wide[:D * 120].contiguous())
        b = ch.process(wide[D * 120:].contiguous())
        ch.reset()
        whole = ch.process(wide)
        torch.cuda.synchronize()
        assert torch.equal(torch.cat([a, b], dim=1), whole)
        ch.reset()
        raw = (wide * 4096).to(torch.int16)                       # CS16 input, the DFT's own output layout
        im = ch.process(raw, instant_major=True)
        torch.cuda.synchronize()
# instant-major input to the demodulator (its own kernel instantiation), full and partial warps
for variant, cc in ((4, 40), (8, 64), (2, 5)):
    x = O.generate(cc, 2000)
    ob2 = O.OracleB(cc)
    c2, _, d2_, _ = ob2.process(x)
    with pkg.Demodulator(cc, 2000) as dm:
        dm.set_kernel_variant(variant)
        r = dm.process(torch.from_numpy(x).cuda().permute(1, 0, 2).contiguous(), dibits=True, instant_major=True)
        torch.cuda.synchronize()
        assert all(np.array_equal(r.dibits[c, :c2[c]].cpu().numpy(), d2_[c, :c2[c]]) for c in range(cc)), variant
torch.cuda.synchronize()
print("sanitize_run ok")
