#!/usr/bin/env python
"""bench_streaming.py -- BASELINE.json configs[4]: "streaming: 64 channels, continuous 32 k-sample chunks with
AGC/FLL/timing/Costas state carried across launches" (SURVEY.md 8d cfg 5: sustained Msamples/s AND per-launch
latency; the chunked output must equal the single-shot output).

    python tools/bench_streaming.py [--channels 64] [--chunk 32768] [--chunks 200]

Two legs: device-resident chunks (kernel latency per launch) and host chunks through tdm_process(TDM_MEM_HOST)
(what a streaming SDR host sees: H2D + kernel + D2H per chunk).  The chunked dibit stream is compared with one
single-shot launch over the same capture (position-weighted checksum per channel).
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

    C_, K, M = args.channels, args.chunk, args.chunks
    N = K * M
    dev = torch.device("cuda", 0)
    iq, _ = pkg.synth_capture(C_, N, device=0)
    torch.cuda.synchronize()

    # single shot (reference for the chunk-invariance check)
    with pkg.Demodulator(C_, N) as one:
        one.use_torch_stream()
        r1 = one.process(iq, dibits=True)
        torch.cuda.synchronize()
        n1 = r1.counts.clone()
        w = torch.arange(1, r1.dibits.shape[1] + 1, device=dev, dtype=torch.int64)
        valid = torch.arange(r1.dibits.shape[1], device=dev)[None, :] < n1[:, None]
        sum1 = (r1.dibits.to(torch.int64) * w * valid).sum(dim=1)
        del r1

    dm = pkg.Demodulator(C_, K)
    dm.use_torch_stream()
    S = dm.max_symbols(K)
    out = pkg.DemodResult(torch.empty(C_, dtype=torch.int32, device=dev), None, torch.empty((C_, S), dtype=torch.uint8, device=dev), None)
    chunks = [iq[:, k * K:(k + 1) * K] for k in range(M)]          # strided views: rows stay where they are in HBM

    # ---- leg 1: device-resident chunks, checksum accumulated on the device
    acc = torch.zeros(C_, dtype=torch.int64, device=dev)
    pos = torch.zeros(C_, dtype=torch.int64, device=dev)
    ar = torch.arange(1, S + 1, device=dev, dtype=torch.int64)
    ai = torch.arange(S, device=dev)
    for k in range(M):
        dm.process(chunks[k], dibits=True, out=out)
        v = ai[None, :] < out.counts[:, None]
        acc += (out.dibits.to(torch.int64) * (pos[:, None] + ar[None, :]) * v).sum(dim=1)
        pos += out.counts.to(torch.int64)
    torch.cuda.synchronize()
    same = bool(torch.equal(acc, sum1) and torch.equal(pos, n1.to(torch.int64)))
    assert same, "chunked output differs from the single-shot output"

    dm.reset_all()
    lat = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(M):
        dm.process(chunks[k], dibits=True, out=out)
    ev1.record()
    torch.cuda.synchronize()
    dev_ms = ev0.elapsed_time(ev1)
    for k in range(min(M, 50)):                                     # per-launch latency: one launch at a time, synchronised
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dm.process(chunks[k], dibits=True, out=out)
        b.record()
        torch.cuda.synchronize()
        lat.append(a.elapsed_time(b))
    lat.sort()

    # ---- leg 2: host chunks through the C ABI
    h = torch.empty((C_, K, 2), dtype=torch.float32, pin_memory=True)
    hd = np.zeros((C_, S), dtype=np.uint8)
    hc = np.zeros(C_, dtype=np.int32)
    import ctypes as Ct
    from sdrpp_tetra_demodulator_b200 import capi
    L = capi.lib()
    vp = lambda a: a.ctypes.data_as(Ct.c_void_p)
    dm.reset_all()
    hn = h.numpy()
    host_lat = []
    m2 = min(M, 100)
    t_host = 0.0
    for k in range(m2):
        h.copy_(chunks[k])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        capi.check(L.tdm_process(dm._h, vp(hn), K, K, None, vp(hd), None, S, vp(hc), capi.TDM_OUT_DIBITS, capi.TDM_MEM_HOST), "tdm_process")
        dt = time.perf_counter() - t0
        t_host += dt
        host_lat.append(dt * 1e3)
    host_lat.sort()

    return {
        "metric": "complex IQ Msamples/s through demod chain (streaming)", "unit": "Msamples/s", "n_gpus": 1,
        "config": {"workload": f"{C_} channels, {M} chunks of {K} samples, loop state carried across launches", "data": "synthetic"},
        "value": round(C_ * N / (dev_ms * 1e-3) / 1e6, 1),
        "ms_per_launch_sustained": round(dev_ms / M, 4),
        "launch_latency_ms": {"median": round(lat[len(lat) // 2], 4), "p95": round(lat[int(len(lat) * 0.95) - 1], 4), "min": round(lat[0], 4)},
        "e2e": {"value": round(C_ * K * m2 / t_host / 1e6, 1), "unit": "Msamples/s",
                "latency_ms": {"median": round(host_lat[len(host_lat) // 2], 4), "p95": round(host_lat[int(len(host_lat) * 0.95) - 1], 4)},
                "h2d_bytes_per_step": C_ * K * 8, "d2h_bytes_per_step": int(hd.nbytes + hc.nbytes),
                "workload": "pinned host chunks via tdm_process(TDM_MEM_HOST), synchronous per chunk"},
        "chunked_equals_single_shot": same,
        "realtime_factor": round((C_ * N / (dev_ms * 1e-3)) / (C_ * 36000.0), 1),
    }



def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", type=int, default=64)
    ap.add_argument("--chunk", type=int, default=32768)
    ap.add_argument("--chunks", type=int, default=200)
    return ap.parse_args(argv)


def main():
    print(json.dumps(measure(parse_args())))


if __name__ == "__main__":
    main()
