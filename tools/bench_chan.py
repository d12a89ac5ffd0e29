#!/usr/bin/env python
"""bench_chan.py -- front-end channeliser (include/tdm_chan_b200.h): one wideband capture ->
M channels at 36 kS/s in HBM, then straight into the demodulator on the same stream.

    python tools/bench_chan.py [--g 128] [--instants 16384] [--steps 5] [--warmup 2]

g = 128: M = 4608 channels on the 25 kHz raster, fs_wide = 115.2 MS/s, D = 3200.  One step = `instants` output samples
per channel (instants * D wideband samples).  roofline: HBM -- algorithmic bytes per wideband sample = 8 (read once)
+ 3 * 8 * M / D (branch sums written + read, channel samples written)."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(args):
    import torch
    import sdrpp_tetra_demodulator_b200 as pkg
    dev = torch.device("cuda", 0)
    cfg = pkg.chan_default_config(args.g)
    M, D = cfg.n_channels, cfg.decimation
    n_out = args.instants
    n_wide = n_out * D
    gen = torch.Generator(device=dev).manual_seed(1)
    wide = torch.randn((n_wide, 2), generator=gen, device=dev, dtype=torch.float32)
    out = torch.empty((M, n_out, 2), dtype=torch.float32, device=dev)         # channel-major (tdm_chan_process, with the transposing pass)
    out_im = torch.empty((n_out, M, 2), dtype=torch.float32, device=dev)      # instant-major (tdm_chan_process_instant_major)
    with pkg.Channelizer(cfg) as ch, pkg.Demodulator(M, 1024) as dm:
        dm.use_torch_stream()
        S = dm.max_symbols(n_out)
        res = pkg.DemodResult(torch.empty(M, dtype=torch.int32, device=dev), None, torch.empty((M, S), dtype=torch.uint8, device=dev), None)
        for _ in range(args.warmup):
            ch.process(wide, out=out)
            ch.process(wide, out=out_im, instant_major=True)
            dm.process(out_im, dibits=True, out=res, instant_major=True)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        pm = dm_ms = ems = cm_ms = 0.0
        for _ in range(args.steps):
            e[0].record()
            ch.process(wide, out=out_im, instant_major=True)
            e[1].record()
            dm.process(out_im, dibits=True, out=res, instant_major=True)
            e[2].record()
            ch.process(wide, out=out)
            e[3].record()
            torch.cuda.synchronize()
            ems += e[0].elapsed_time(e[1])
            dm_ms += e[1].elapsed_time(e[2])
            cm_ms += e[2].elapsed_time(e[3])
            ch.process(wide, out=out_im, instant_major=True)
            torch.cuda.synchronize()
            a, b = ch.last_kernel_ms()
            pm += a
        cm_ms /= args.steps
        ems /= args.steps; dm_ms /= args.steps; pm /= args.steps
        # ---- end to end from HOST memory: one pinned wideband buffer in, dibits + counts out, in `parts` sub-chunks so the copy of
        # chunk k+1 crosses PCIe while chunk k is channelised and demodulated (two device buffers, copy stream + compute stream)
        import time
        parts = 8
        ci = n_out // parts
        wide_h = torch.empty((n_wide, 2), dtype=torch.float32).pin_memory()
        wide_h.copy_(wide)
        Sc = dm.max_symbols(ci)
        dib_h = torch.empty((parts, M, Sc), dtype=torch.uint8).pin_memory()
        cnt_h = torch.empty((parts, M), dtype=torch.int32).pin_memory()
        wbuf = [torch.empty((ci * D, 2), dtype=torch.float32, device=dev) for _ in range(2)]
        obuf = torch.empty((ci, M, 2), dtype=torch.float32, device=dev)
        rbuf = [pkg.DemodResult(torch.empty(M, dtype=torch.int32, device=dev), None, torch.empty((M, Sc), dtype=torch.uint8, device=dev), None) for _ in range(2)]
        copy_s, back_s, comp_s = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev), torch.cuda.current_stream(dev)
        landed = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        drained = [torch.cuda.Event() for _ in range(2)]

        def e2e_step(src_h, stage):
            for k in range(parts):
                b = k & 1
                with torch.cuda.stream(copy_s):
                    if k >= 2:
                        copy_s.wait_event(freed[b])
                    stage[b].copy_(src_h[k * ci * D:(k + 1) * ci * D], non_blocking=True)
                    landed[b].record(copy_s)
                comp_s.wait_event(landed[b])
                ch.process(stage[b], out=obuf, instant_major=True)
                freed[b].record(comp_s)
                if k >= 2:
                    comp_s.wait_event(drained[b])
                dm.process(obuf, dibits=True, out=rbuf[b], instant_major=True)
                ready = torch.cuda.Event()
                ready.record(comp_s)
                with torch.cuda.stream(back_s):               # results go back on a stream of their own: the next H2D must not queue behind them
                    back_s.wait_event(ready)
                    dib_h[k].copy_(rbuf[b].dibits, non_blocking=True)
                    cnt_h[k].copy_(rbuf[b].counts, non_blocking=True)
                    drained[b].record(back_s)
            torch.cuda.synchronize()

        def time_e2e(src_h, stage):
            ch.reset(); dm.reset_all()
            e2e_step(src_h, stage)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                e2e_step(src_h, stage)
            return (time.perf_counter() - t0) / args.steps * 1e3

        e2e_ms = time_e2e(wide_h, wbuf)
        # the same capture as int16 pairs (TDM_CHAN_IN_CS16): what SDR hardware delivers; half the bytes across PCIe
        wide16_h = torch.empty((n_wide, 2), dtype=torch.int16).pin_memory()
        wide16_h.copy_((wide.clamp(-3.9, 3.9) * 8192.0).to(torch.int16))
        wbuf16 = [torch.empty((ci * D, 2), dtype=torch.int16, device=dev) for _ in range(2)]
        e2e16_ms = time_e2e(wide16_h, wbuf16)
        h2d_bytes = wide_h.numel() * 4
        d2h_bytes = dib_h.numel() + cnt_h.numel() * 4
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak, src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    bytes_per_wide = 8.0 + 3 * 8.0 * M / D
    ach = bytes_per_wide * n_wide / (ems * 1e-3) / 1e9
    return {
        "metric": "wideband complex Msamples/s through the channeliser", "value": round(n_wide / (ems * 1e-3) / 1e6, 1), "unit": "Msamples/s",
        "channel_msps": round(M * n_out / (ems * 1e-3) / 1e6, 1), "ms_per_step": round(ems, 3), "polyphase_kernel_ms": round(pm, 3),
        "dft_ms": round(ems - pm, 3), "layout": "instant-major [sample][channel] (tdm_chan_process_instant_major; read in place by tdm_process_io with sample_stride)",
        "channel_major_ms_per_step": round(cm_ms, 3), "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "dtype": "f32", "data": "synthetic (white noise)",
        "config": {"workload": f"{M} channels on a 25 kHz raster from one {0.9 * args.g:.1f} MS/s capture (D = {D}, {cfg.taps_per_branch} taps per branch), "
                               f"{n_out} output samples per channel per step"},
        "then_demodulated_ms": round(dm_ms, 3),
        "e2e": {"value": round(M * n_out / (e2e_ms * 1e-3) / 1e6, 1), "unit": "channel Msamples/s", "ms_per_step": round(e2e_ms, 3),
                "wideband_msps": round(n_wide / (e2e_ms * 1e-3) / 1e6, 1), "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                "h2d_gbs": round(h2d_bytes / (e2e_ms * 1e-3) / 1e9, 2),
                "workload": f"pinned host wideband capture -> H2D -> channeliser -> demodulator -> dibits + counts D2H, {parts} sub-chunks pipelined over three streams; "
                            f"{8.0 * D / M:.2f} B cross PCIe per channel sample instead of 8 (one wideband stream in instead of {M} channelised float streams)",
                "cs16": {"value": round(M * n_out / (e2e16_ms * 1e-3) / 1e6, 1), "unit": "channel Msamples/s", "ms_per_step": round(e2e16_ms, 3),
                         "h2d_bytes_per_step": int(h2d_bytes // 2), "h2d_gbs": round(h2d_bytes / 2 / (e2e16_ms * 1e-3) / 1e9, 2),
                         "workload": "the same with the capture as interleaved int16 (TDM_CHAN_IN_CS16, what SDR hardware delivers): "
                                     f"{4.0 * D / M:.2f} B per channel sample across PCIe"}},
        "chain_channel_msps": round(M * n_out / ((ems + dm_ms) * 1e-3) / 1e6, 1),
        "roofline": {"bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": None,
                     "peak_source": src, "algorithmic_bytes_per_wideband_sample": round(bytes_per_wide, 2),
                     "kernel": "chan_residue_kernel + cuFFT C2C (batched)"},
        "parity": "unpinned by the reference (no channeliser there); fp64 defining sum within 2e-5 and wideband -> dibits end to end in tests/test_chan_gpu.py",
    }


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--g", type=int, default=128)
    ap.add_argument("--instants", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    return ap.parse_args(argv)


def main():
    print(json.dumps(measure(parse_args())))


if __name__ == "__main__":
    main()
