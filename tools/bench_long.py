#!/usr/bin/env python
"""bench_long.py -- BASELINE.json configs[1]: "1 channel, 1e9 complex samples on a single B200, fused kernel vs CPU
reference", through tdm_process_long (time-segment parallelism).

    python tools/bench_long.py [--samples 1000000000] [--segments 4096] [--warmup 65536] [--steps 3] [--warmup-steps 1]
    python tools/bench_long.py --channels 256 --samples 4000000      # BASELINE.json configs[2] through tdm_process_long_batch

One step = the whole capture (8 GB of IQ resident in HBM) through tdm_process_long, loop state carried from the
previous step.  Beside it: the same chain walked sequentially on the GPU (one channel = one lane of one warp: the
recurrence's latency, measured on a bounded sample) and the reference's CPU chain on one host core (the reference
cannot split a channel either).  Correctness at this size: the decoded dibits equal the TRANSMITTED dibits after the
chain's fixed lag over the last three quarters of the capture (zero errors demanded at 30 dB).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(args):

    import numpy as np
    import torch
    import sdrpp_tetra_demodulator_b200 as pkg

    dev = torch.device("cuda", 0)
    N, C_ = args.samples, args.channels
    iq, tx = pkg.synth_capture(C_, N, device=0, want_tx=True)
    torch.cuda.synchronize()
    dm = pkg.Demodulator(args.segments, 1024)
    dm.use_torch_stream()
    out = torch.empty((C_, N // 2 + 64), dtype=torch.uint8, device=dev)

    def run():
        if C_ == 1:
            d, info = dm.process_long(iq[0], warmup=args.warmup, out=out[0])
            return out, torch.tensor([info["n_dibits"]], device=dev), info
        return dm.process_long_batch(iq, warmup=args.warmup, out=out)

    infos = []
    for _ in range(args.warmup_steps):
        dm.reset_all()
        run()
    torch.cuda.synchronize()
    l0 = dm.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        infos.append(run()[2])
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = dm.launch_count() - l0

    # correctness: one clean pass from reset, against the transmitted dibits
    dm.reset_all()
    d, cnt, info = run()
    torch.cuda.synchronize()
    cnt = cnt.cpu().numpy()
    assert int(abs(cnt - N // 2).max()) <= 4, (cnt.min(), cnt.max(), N // 2)
    n = int(cnt.min())
    skip, m = n // 4, n - 64
    best = torch.full((C_,), 1 << 40, dtype=torch.int64, device=dev)
    for lag in range(12, 26):
        best = torch.minimum(best, (d[:, lag + skip:lag + m] != tx[:, skip:m]).sum(dim=1))
    errs = int(best.sum())
    bad_channels = int((best > 0).sum())
    if C_ == 1:
        assert errs == 0, f"{errs} dibit errors against the transmitted stream in the last three quarters"
    else:
        assert bad_channels <= C_ // 100, f"{bad_channels} of {C_} channels with errors in the last three quarters"

    # the same capture through the plain batch call: every channel one recurrence (bounded sample)
    ns = min(N, 4_000_000)
    with pkg.Demodulator(C_, ns) as one:
        one.use_torch_stream()
        r = one.process(iq[:, :ns], dibits=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        one.process(iq[:, :ns], dibits=True, out=r)
        b.record()
        torch.cuda.synchronize()
        seq_msps = C_ * ns / (a.elapsed_time(b) * 1e-3) / 1e6

    cpu = None
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        nc = 20_000_000
        base = O.generate(1, 100_000)
        cap = np.ascontiguousarray(np.tile(base, (1, nc // base.shape[1], 1)))
        if O.have_ref():
            ra = O.OracleA(1)
            t0 = time.perf_counter()
            for k in range(0, nc, 1_000_000):
                ra.process_multi(cap[:, k:k + 1_000_000], 1)
            dt = time.perf_counter() - t0
            ra.close()
            kind, what = "reference", "reference src/dsp/*.cpp (oracle/_ref, g++ -O3 -ffp-contract=off, scalar VOLK stand-in)"
        else:
            ob = O.OracleB(1)
            t0 = time.perf_counter()
            ob.process(cap, want_syms=False, nthreads=1)
            dt = time.perf_counter() - t0
            kind, what = "port", "oracle_b.c canonical-order port"
        cpu = {"value": round(nc / dt / 1e6, 3), "unit": "Msamples/s", "cores": 1, "kind": kind,
               "sample": f"1 channel x {nc} samples in 1e6-sample calls, one thread (a channel is one recurrence: the reference "
                         f"cannot use a second core for it); {what}; {dt:.1f} s"}

    value = C_ * N / (ms * 1e-3) / 1e6
    return {
        "metric": f"complex IQ Msamples/s through demod chain ({C_} channel{'s' if C_ > 1 else ''}, time-segmented)", "value": round(value, 1), "unit": "Msamples/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup_steps, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{C_} channel(s) x {N} samples (pi/4-DQPSK, 2 sps, SNR 30 dB), {info['n_segments']} segments per channel of "
                               f"{info['segment_samples']} + {info['warmup']} warm-up samples", "l2": "inputs larger than L2",
                   "segments_redone": [i["n_rerun"] for i in infos], "segments_joined_without_agreement": [i["n_forced"] for i in infos],
                   "dibit_errors_vs_transmitted": errs, "channels_with_errors": bad_channels},
        "gpu_launches": int(launches),
        "plain_batch_call_msps": round(seq_msps, 2),
        "overhead_vs_batch_kernel": f"{info['warmup']}/{info['segment_samples']} warm-up samples redone per segment",
        "cpu_baseline": cpu,
    }



def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=1_000_000_000)
    ap.add_argument("--channels", type=int, default=1)
    ap.add_argument("--segments", type=int, default=4096)
    ap.add_argument("--warmup", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup-steps", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args(argv)


def main():
    print(json.dumps(measure(parse_args())))


if __name__ == "__main__":
    main()
